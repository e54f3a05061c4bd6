import sys; sys.path.insert(0, ".")
import torch, bench
from sakura_b200.capi import SakuraLib
from sakura_b200.device_api import DeviceFits
lib = SakuraLib(); dev = DeviceFits(lib)
d = torch.device("cuda", 0)
for n, R in ((2048, 524288), (4096, 262144)):
    data, mask = bench.generate_on_device(torch, R, n, seed=1, device=d)
    st, ctx = lib.create_polynomial_context(5, n)
    for nfit in (5, 3):
        out = dev.polynomial(ctx, 5, data, mask, 3.0, nfit)
        torch.cuda.synchronize()
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(5): dev.polynomial(ctx, 5, data, mask, 3.0, nfit, out=out)
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 5
        print(n, R, nfit, lib.last_kernel_name(), f"{ms:.2f} ms {R/ms*1e3:.3e} spectra/s  frac {R*(n*10+56)/ms*1e3/1e9/6534.1:.3f}")
    lib.destroy_context(ctx)
