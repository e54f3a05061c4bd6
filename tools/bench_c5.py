"""Calibration fused into the fit vs calibrate-then-fit, device-resident."""
import sys
sys.path.insert(0, ".")
import torch
from sakura_b200.capi import SakuraLib
from sakura_b200.device_api import DeviceFits
import bench
lib = SakuraLib(); dev = DeviceFits(lib)
R, n = (int(sys.argv[1]) if len(sys.argv) > 1 else 262144), (int(sys.argv[2]) if len(sys.argv) > 2 else 4096)
device = torch.device("cuda", 0)
data, mask = bench.generate_on_device(torch, R, n, seed=5, device=device)
off = 300.0 + 20.0 * torch.randn(R, n, device=device)
on = off * (1.0 + data / 250.0)
tsys = 80.0 + 170.0 * torch.rand(R, n, device=device)
del data
st, ctx = lib.create_polynomial_context(5, n)
out = {k: None for k in ()}
cal = torch.empty_like(on)
def seq():
    dev.calibrate(tsys, on, off, out=cal)
    return dev.polynomial(ctx, 5, cal, mask, 3.0, 5, out=res_seq)
def fused():
    return dev.calibrate_and_polynomial(ctx, 5, tsys, on, off, mask, 3.0, 5, out=res_fused)
res_seq = dev.polynomial(ctx, 5, on, mask, 3.0, 5); res_seq.pop("call_status")
res_fused = dev.polynomial(ctx, 5, on, mask, 3.0, 5); res_fused.pop("call_status")
def timeit(fn, reps=5):
    fn(); fn(); torch.cuda.synchronize()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps
t_cal = timeit(lambda: dev.calibrate(tsys, on, off, out=cal))
t_fit = timeit(lambda: dev.polynomial(ctx, 5, cal, mask, 3.0, 5, out=res_seq))
t_seq = timeit(seq); t_fused = timeit(fused)
same = torch.equal(res_seq["final_mask"], res_fused["final_mask"]) and torch.equal(res_seq["residual"].nan_to_num(), res_fused["residual"].nan_to_num())
print(f"rows {R}: fit alone {t_fit:.2f} ms, calibrate alone {t_cal:.2f} ms ({R*n*16/t_cal/1e6:.0f} GB/s), "
      f"calibrate+fit {t_seq:.2f} ms -> {R/t_seq*1e3:.3e} spectra/s, fused {t_fused:.2f} ms -> {R/t_fused*1e3:.3e} spectra/s, identical {same}")
