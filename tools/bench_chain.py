"""The e2eReduce chain on device arrays (BASELINE configs[4] shape per GPU: 524288 x 2048):
Tsys interpolation -> flags to mask + edge flags -> calibration fused into the poly-5 fit (NaN/Inf flags
merged) -> clipping of the residual -> statistics.  bench/e2eReduce/src/main.cc:80-240."""
import sys
sys.path.insert(0, ".")
import numpy as np, torch
from sakura_b200.capi import SakuraLib
from sakura_b200.device_api import DeviceFits
import bench
lib = SakuraLib(); dev = DeviceFits(lib)
R = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 19
n, nfit, edge = 2048, 5, 32
device = torch.device("cuda", 0)
sky, mask0 = bench.generate_on_device(torch, R, n, seed=5, device=device)
off = 300.0 + 20.0 * torch.randn(R, n, device=device)
on = off * (1.0 + sky / 250.0)
del sky
flags = (~mask0).to(torch.uint8) * 4          # MS-style flag bytes: non-zero = bad
del mask0
t_base = torch.linspace(0.0, 100.0, 6, dtype=torch.float64, device=device)
tsys_base = (150.0 + 10.0 * torch.randn(6, n, device=device)).float()
tsys_valid = torch.rand(6, n, device=device) < 0.98
t_rows = torch.sort(torch.rand(R, dtype=torch.float64, device=device) * 110.0 - 5.0).values
chan = torch.arange(n, dtype=torch.int32, device=device).repeat(R, 1)
st, ctx = lib.create_polynomial_context(5, n); assert st == 0
tsys = torch.empty(R, n, dtype=torch.float32, device=device); tsys_m = torch.empty(R, n, dtype=torch.bool, device=device)
mask = torch.empty(R, n, dtype=torch.bool, device=device); clipm = torch.empty(R, n, dtype=torch.bool, device=device)
res = dev.polynomial(ctx, 5, on, flags == 0, 3.0, nfit); res.pop("call_status")

def s_interp(): dev.interpolate_y(1, t_base, tsys_base, tsys_valid, t_rows, out=tsys, out_mask=tsys_m)
def s_mask():
    dev.uint8_to_bool(flags, invert=True, out=mask)
    dev.in_ranges(chan, [edge - 1], [n - edge], inclusive=False, and_with=mask, out=mask)
def s_fit(): dev.calibrate_and_polynomial(ctx, 5, tsys, on, off, mask, 3.0, nfit, flag_nan_or_inf=True, out=res)
def s_clip(): dev.in_ranges(res["residual"], [-5.0], [5.0], inclusive=False, and_with=res["final_mask"], out=clipm)
def s_stats(): return dev.statistics(res["residual"], clipm)
steps = [("interpolate Tsys (6 -> rows, linear)", s_interp, 5 * n), ("flags -> mask, edge flags", s_mask, (1 + 1) + (4 + 1 + 1)),
         ("calibrate + fit (fused, NaN/Inf merged)", s_fit, None), ("clip residual, merge", s_clip, 6 * n), ("statistics", s_stats, 5 * n)]
def timeit(fn, reps=5):
    fn(); fn(); torch.cuda.synchronize()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps
total = 0.0
for name, fn, bpr in steps:
    ms = timeit(fn); total += ms
    bw = "" if bpr is None else f"  {R * (bpr if bpr > 100 else bpr * n) / ms / 1e6:7.0f} GB/s"
    print(f"{name:42s} {ms:8.3f} ms{bw}   [{lib.last_kernel_name()}]")
def chain():
    for _, fn, _ in steps: fn()
ms = timeit(chain)
ok = int((res["status"] == 0).sum())
print(f"chain: {ms:.2f} ms for {R} x {n} -> {R / ms * 1e3:.3e} spectra/s per GPU (sum of steps {total:.2f} ms), rows ok {ok}")
