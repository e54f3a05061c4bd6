import sys
sys.path.insert(0, ".")
import torch
from sakura_b200.capi import SakuraLib
from sakura_b200.device_api import DeviceFits
import bench
lib = SakuraLib(); dev = DeviceFits(lib)
R, n = 1 << 19, int(sys.argv[1]) if len(sys.argv) > 1 else 2048
if n == 4096: R = 1 << 18
device = torch.device("cuda", 0)
data, mask = bench.generate_on_device(torch, R, n, seed=5, device=device)
off = 300.0 + 20.0 * torch.randn(R, n, device=device)
on = off * (1.0 + data / 250.0)
tsys = 80.0 + 170.0 * torch.rand(R, n, device=device)
st, ctx = lib.create_polynomial_context(5, n)
res = dev.polynomial(ctx, 5, data, mask, 3.0, 5); res.pop("call_status")
def timeit(fn, reps=5):
    fn(); fn(); torch.cuda.synchronize()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps
t0 = timeit(lambda: dev.polynomial(ctx, 5, data, mask, 3.0, 5, out=res)); k0 = lib.last_kernel_name()
t1 = timeit(lambda: dev.calibrate_and_polynomial(ctx, 5, tsys, on, off, mask, 3.0, 5, flag_nan_or_inf=True, out=res)); k1 = lib.last_kernel_name()
print(f"n {n} rows {R}: fit {t0:.2f} ms [{k0}]  calibrate+fit {t1:.2f} ms [{k1}] -> {R/t1*1e3:.3e} spectra/s")
