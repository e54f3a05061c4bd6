"""Development timing of the spline (C3) and sinusoid (C4) shapes: device-resident rows through the
batched device entry points vs the CPU oracle on a sample."""
import sys, time
sys.path.insert(0, ".")
import numpy as np, torch
from sakura_b200.capi import SakuraLib
from sakura_b200 import synth
from oracle.refbind import Oracle
import bench
lib = SakuraLib(); ora = Oracle()
R, n = int(sys.argv[1]) if len(sys.argv) > 1 else 4096, 8192
dev = torch.device("cuda", 0)
data, mask = bench.generate_on_device(torch, R, n, seed=3, device=dev)
res = torch.empty(R, n, dtype=torch.float32, device=dev); fm = torch.empty(R, n, dtype=torch.bool, device=dev)
rms = torch.empty(R, dtype=torch.float32, device=dev); st = torch.empty(R, dtype=torch.int32, device=dev); ls = torch.empty(R, dtype=torch.int32, device=dev)
def timeit(fn, reps=3):
    fn(); torch.cuda.synchronize()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps
# spline 20 pieces
s, ctx = lib.create_cubicspline_context(20, n)
co = torch.empty(R, 20, 4, dtype=torch.float64, device=dev); bd = torch.empty(R, 21, dtype=torch.int64, device=dev)
f = lambda: lib.lib.sakura_LSQFitCubicSplineFloatBatchDevice(ctx, 20, R, n, data.data_ptr(), n, mask.data_ptr(), n, 3.0, 3, co.data_ptr(), None, res.data_ptr(), fm.data_ptr(), rms.data_ptr(), bd.data_ptr(), st.data_ptr(), ls.data_ptr(), None)
assert f() == 0
ms = timeit(f); print(f"C3 spline20 {R}x{n}: {ms:.2f} ms -> {R/ms*1e3:.3e} spectra/s, ok rows {(st==0).sum().item()}")
hd, hm = data[:256].cpu().numpy(), mask[:256].cpu().numpy()
t = time.perf_counter(); o = ora.lsqfit_cubicspline_batch(20, hd, hm, 3.0, 3, nthreads=16); dt = time.perf_counter() - t
print(f"   CPU oracle 256 rows: {256/dt:.3e} spectra/s; mask mismatches vs gpu {(fm[:256].cpu().numpy() != o['final_mask']).sum()}")
lib.destroy_context(ctx)
# sinusoid nwave 0..20
s, ctx = lib.create_sinusoid_context(20, n)
nw = np.arange(21, dtype=np.uint64); co = torch.empty(R, 41, dtype=torch.float64, device=dev)
f = lambda: lib.lib.sakura_LSQFitSinusoidFloatBatchDevice(ctx, 21, nw.ctypes.data, R, n, data.data_ptr(), n, mask.data_ptr(), n, 3.0, 3, 41, co.data_ptr(), None, res.data_ptr(), fm.data_ptr(), rms.data_ptr(), st.data_ptr(), ls.data_ptr(), None)
assert f() == 0
ms = timeit(f); print(f"C4 sinusoid41 {R}x{n}: {ms:.2f} ms -> {R/ms*1e3:.3e} spectra/s, ok rows {(st==0).sum().item()}")
t = time.perf_counter(); o = ora.lsqfit_sinusoid_batch(list(range(21)), hd, hm, 3.0, 3, nthreads=16); dt = time.perf_counter() - t
print(f"   CPU oracle 256 rows: {256/dt:.3e} spectra/s; mask mismatches vs gpu {(fm[:256].cpu().numpy() != o['final_mask']).sum()}")
lib.destroy_context(ctx)
