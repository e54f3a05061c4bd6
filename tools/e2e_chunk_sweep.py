"""e2e (pinned host buffers through sakura_LSQFitPolynomialFloatBatch) against the chunk size of the
host pipeline, next to the copy-only time of the same bytes."""
import sys, time
sys.path.insert(0, ".")
import torch
from sakura_b200.capi import SakuraLib
import bench
lib = SakuraLib(); R, n, K = 262144, 4096, 6
dev = torch.device("cuda", 0)
data, mask = bench.generate_on_device(torch, R, n, seed=1, device=dev)
h = {k: torch.empty(R, n, dtype=t, pin_memory=True) for k, t in (("data", torch.float32), ("mask", torch.bool), ("resid", torch.float32), ("fmask", torch.bool))}
h["data"].copy_(data); h["mask"].copy_(mask)
co = torch.empty(R, K, dtype=torch.float64, pin_memory=True); rms = torch.empty(R, dtype=torch.float32, pin_memory=True)
st = torch.empty(R, dtype=torch.int32, pin_memory=True); ls = torch.empty(R, dtype=torch.int32, pin_memory=True)
s, ctx = lib.create_polynomial_context(5, n)
def call():
    rc = lib.lib.sakura_LSQFitPolynomialFloatBatch(ctx, 5, R, n, h["data"].data_ptr(), n, h["mask"].data_ptr(), n, 3.0, 5, K, co.data_ptr(), None,
                                                   h["resid"].data_ptr(), h["fmask"].data_ptr(), rms.data_ptr(), st.data_ptr(), ls.data_ptr())
    assert rc == 0, lib.last_error()
for mib in (24, 48, 96, 144, 192, 288):
    lib.set_host_pipeline_tuning(mib, -1)
    call(); torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(3): call()
    dt = (time.perf_counter() - t0) / 3
    print(f"chunk {mib:4d} MiB: {dt*1e3:7.2f} ms  {R/dt:.3e} spectra/s  {R*n*5/dt/1e9:.1f} GB/s per direction")
