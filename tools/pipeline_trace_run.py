import sys, time
sys.path.insert(0, ".")
import torch
from sakura_b200.capi import SakuraLib
import bench
lib = SakuraLib(); R, n, K = 131072, 4096, 6
dev = torch.device("cuda", 0)
data, mask = bench.generate_on_device(torch, R, n, seed=1, device=dev)
h = {k: torch.empty(R, n, dtype=t, pin_memory=True) for k, t in (("data", torch.float32), ("mask", torch.bool), ("resid", torch.float32), ("fmask", torch.bool))}
h["data"].copy_(data); h["mask"].copy_(mask)
co = torch.empty(R, K, dtype=torch.float64, pin_memory=True); rms = torch.empty(R, dtype=torch.float32, pin_memory=True)
st = torch.empty(R, dtype=torch.int32, pin_memory=True); ls = torch.empty(R, dtype=torch.int32, pin_memory=True)
s, ctx = lib.create_polynomial_context(5, n)
for i in range(2):
    rc = lib.lib.sakura_LSQFitPolynomialFloatBatch(ctx, 5, R, n, h["data"].data_ptr(), n, h["mask"].data_ptr(), n, 3.0, 5, K, co.data_ptr(), None,
                                                   h["resid"].data_ptr(), h["fmask"].data_ptr(), rms.data_ptr(), st.data_ptr(), ls.data_ptr())
