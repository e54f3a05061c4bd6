"""e2e (pinned host buffers through sakura_LSQFitPolynomialFloatBatch) with the masks over PCIe as bits
(default) against bytes (SAKURA_B200_PACK_MASKS=0), same process, alternating."""
import os, sys, time
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
ref = None
for rep in range(2):
    for pack in ("1", "0"):
        os.environ["SAKURA_B200_PACK_MASKS"] = pack
        h["fmask"].zero_()
        call(); torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(3): call()
        dt = (time.perf_counter() - t0) / 3
        fm = h["fmask"].clone()
        same = True if ref is None else bool(torch.equal(fm, ref))
        ref = fm if ref is None else ref
        print(f"pack {pack}: {dt*1e3:7.2f} ms  {R/dt:.3e} spectra/s  final_mask equal to first run: {same}, valid {int(fm.sum())}")
for mib in (24, 48, 144):
    os.environ["SAKURA_B200_PACK_MASKS"] = "1"
    lib.set_host_pipeline_tuning(mib, -1)
    call(); torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(3): call()
    dt = (time.perf_counter() - t0) / 3
    print(f"pack 1, chunk {mib} MiB: {dt*1e3:7.2f} ms  {R/dt:.3e} spectra/s")
