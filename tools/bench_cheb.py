"""Chebyshev order-5 fits (same entry point as the polynomial, LSQFitType_kChebyshev)."""
import sys
sys.path.insert(0, ".")
import torch
from sakura_b200.capi import SakuraLib, LSQFIT_TYPE_CHEBYSHEV
from sakura_b200.device_api import DeviceFits
import bench
lib = SakuraLib(); dev = DeviceFits(lib)
R, n = 131072, 4096
data, mask = bench.generate_on_device(torch, R, n, seed=5, device=torch.device("cuda", 0))
st, ctx = lib.create_polynomial_context(5, n, lsqfit_type=LSQFIT_TYPE_CHEBYSHEV)
res = dev.polynomial(ctx, 5, data, mask, 3.0, 5); res.pop("call_status")
torch.cuda.synchronize()
e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(3): dev.polynomial(ctx, 5, data, mask, 3.0, 5, out=res)
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 3
print(f"chebyshev-5 {R} x {n}: {ms:.2f} ms, {R/ms*1e3:.3e} spectra/s, kernel {lib.last_kernel_name()}")
