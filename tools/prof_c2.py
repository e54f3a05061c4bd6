"""ncu target: the polynomial (C2 shape) kernel on 131072 rows x 4096 channels, order 5, 5 fits."""
import sys
sys.path.insert(0, ".")
import numpy as np, torch
from sakura_b200.capi import SakuraLib
import bench
lib = SakuraLib(); R, n = 131072, 4096
nfit = int(sys.argv[1]) if len(sys.argv) > 1 else 5
dev = torch.device("cuda", 0)
data, mask = bench.generate_on_device(torch, R, n, seed=20240611, device=dev)
res = torch.empty(R, n, dtype=torch.float32, device=dev); fm = torch.empty(R, n, dtype=torch.bool, device=dev)
rms = torch.empty(R, dtype=torch.float32, device=dev); st = torch.empty(R, dtype=torch.int32, device=dev); ls = torch.empty(R, dtype=torch.int32, device=dev)
co = torch.empty(R, 6, dtype=torch.float64, device=dev)
s, ctx = lib.create_polynomial_context(5, n)
ev0 = torch.cuda.Event(enable_timing=True); ev1 = torch.cuda.Event(enable_timing=True)
for it in range(3):
    ev0.record()
    rc = lib.lib.sakura_LSQFitPolynomialFloatBatchDevice(ctx, 5, R, n, data.data_ptr(), n, mask.data_ptr(), n, 3.0, nfit, 6, co.data_ptr(), None, res.data_ptr(), fm.data_ptr(), rms.data_ptr(), st.data_ptr(), ls.data_ptr(), None)
    ev1.record()
    assert rc == 0, rc
torch.cuda.synchronize()
ms = ev0.elapsed_time(ev1)
print("ok", int((st == 0).sum()), "ms", ms, "spectra/s", R / ms * 1e3, lib.last_kernel_name(), "handed over", lib.last_handover_count(ctx))
