"""Chebyshev fits (same entry point as the polynomial, LSQFitType_kChebyshev): order 5 goes through
the polynomial kernel, higher orders through the tensor-core kernel of the sinusoid fit."""
import sys
sys.path.insert(0, ".")
import torch
from sakura_b200.capi import SakuraLib, LSQFIT_TYPE_CHEBYSHEV
from sakura_b200.device_api import DeviceFits
import bench
lib = SakuraLib(); dev = DeviceFits(lib)
order = int(sys.argv[1]) if len(sys.argv) > 1 else 5
R, n = (131072, 4096) if order <= 5 else (16384, 8192)
data, mask = bench.generate_on_device(torch, R, n, seed=5, device=torch.device("cuda", 0))
st, ctx = lib.create_polynomial_context(order, n, lsqfit_type=LSQFIT_TYPE_CHEBYSHEV)
res = dev.polynomial(ctx, order, data, mask, 3.0, 3); res.pop("call_status")
torch.cuda.synchronize()
e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(3): dev.polynomial(ctx, order, data, mask, 3.0, 3, out=res)
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 3
print(f"chebyshev-{order} {R} x {n}: {ms:.2f} ms, {R/ms*1e3:.3e} spectra/s, kernel {lib.last_kernel_name()}")
