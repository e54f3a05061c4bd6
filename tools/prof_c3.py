"""ncu target: a few launches of the cubic-spline (C3) kernel on 4736 rows of 8192 channels."""
import sys
sys.path.insert(0, ".")
import numpy as np, torch
from sakura_b200.capi import SakuraLib
import bench
lib = SakuraLib(); R, n, npiece = 4736, 8192, 20
dev = torch.device("cuda", 0)
data, mask = bench.generate_on_device(torch, R, n, seed=3, device=dev)
res = torch.empty(R, n, dtype=torch.float32, device=dev); fm = torch.empty(R, n, dtype=torch.bool, device=dev)
rms = torch.empty(R, dtype=torch.float32, device=dev); st = torch.empty(R, dtype=torch.int32, device=dev); ls = torch.empty(R, dtype=torch.int32, device=dev)
s, ctx = lib.create_cubicspline_context(npiece, n)
co = torch.empty(R, npiece, 4, dtype=torch.float64, device=dev); bd = torch.empty(R, npiece + 1, dtype=torch.int64, device=dev)
ev0 = torch.cuda.Event(enable_timing=True); ev1 = torch.cuda.Event(enable_timing=True)
for it in range(3):
    ev0.record()
    rc = lib.lib.sakura_LSQFitCubicSplineFloatBatchDevice(ctx, npiece, R, n, data.data_ptr(), n, mask.data_ptr(), n, 3.0, 3, co.data_ptr(), None, res.data_ptr(), fm.data_ptr(), rms.data_ptr(), bd.data_ptr(), st.data_ptr(), ls.data_ptr(), None)
    ev1.record()
    assert rc == 0, rc
torch.cuda.synchronize()
print("ok", int((st == 0).sum()), "ms", ev0.elapsed_time(ev1), lib.last_kernel_name())
