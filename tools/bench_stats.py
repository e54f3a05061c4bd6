"""Device-resident throughput of the batched statistics kernel (5 B per element: HBM-bound)."""
import sys
sys.path.insert(0, ".")
import torch
from sakura_b200.capi import SakuraLib
from sakura_b200.device_api import DeviceFits
lib = SakuraLib(); dev = DeviceFits(lib)
for R, n in ((1 << 20, 4096), (1 << 18, 8192), (4096, 1 << 20), (8, 1 << 27)):
    d = torch.randn(R, n, device="cuda"); m = torch.rand(R, n, device="cuda") > 0.1
    dev.statistics(d, m); torch.cuda.synchronize()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5): dev.statistics(d, m)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 5
    print(f"{R} x {n}: {ms:.3f} ms, {R*n*5/ms/1e6:.0f} GB/s, {R/ms*1e3:.3e} rows/s")
    del d, m
