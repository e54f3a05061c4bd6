"""Streaming rate of the boolean-filter kernel (csrc/bool_filters.cu) on device arrays."""
import ctypes as C, sys
sys.path.insert(0, ".")
import numpy as np, torch
from sakura_b200.capi import SakuraLib
lib = SakuraLib(); L = lib.lib; vp = C.c_void_p
n = 1 << 30
f = torch.randn(n, dtype=torch.float32, device="cuda"); m = torch.ones(n, dtype=torch.bool, device="cuda")
out = torch.empty(n, dtype=torch.bool, device="cuda")
lo = np.array([-1.0, 0.5], dtype=np.float32); hi = np.array([-0.25, 2.0], dtype=np.float32)
def timed(fn, bytes_per_elem, name):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    for _ in range(3): fn()
    e0.record()
    for _ in range(10): fn()
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 10
    print(f"{name:40s} {ms:7.3f} ms  {n * bytes_per_elem / ms / 1e6:7.1f} GB/s")
timed(lambda: L.sakura_SetTrueIfInRangesFloatDevice(False, n, vp(f.data_ptr()), 2, vp(lo.ctypes.data), vp(hi.ctypes.data), None, vp(out.data_ptr()), None), 5, "ranges exclusive float, 2 conditions")
timed(lambda: L.sakura_SetTrueIfInRangesFloatDevice(False, n, vp(f.data_ptr()), 2, vp(lo.ctypes.data), vp(hi.ctypes.data), vp(m.data_ptr()), vp(m.data_ptr()), None), 6, "... ANDed into an existing mask")
timed(lambda: L.sakura_SetFalseIfNanOrInfFloatDevice(n, vp(f.data_ptr()), None, vp(out.data_ptr()), None), 5, "finite flags")
timed(lambda: L.sakura_InvertBoolDevice(n, vp(m.data_ptr()), None, vp(out.data_ptr()), None), 2, "invert bool")
