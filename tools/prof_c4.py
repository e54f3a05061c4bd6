"""ncu target: a few launches of the sinusoid (C4) kernel on 9472 rows (296 tiles = 2 per SM)."""
import sys
sys.path.insert(0, ".")
import numpy as np, torch
from sakura_b200.capi import SakuraLib
import bench
lib = SakuraLib(); R, n = 9472, 8192  # 296 tiles: two resident CTAs per SM
dev = torch.device("cuda", 0)
data, mask = bench.generate_on_device(torch, R, n, seed=3, device=dev)
res = torch.empty(R, n, dtype=torch.float32, device=dev); fm = torch.empty(R, n, dtype=torch.bool, device=dev)
rms = torch.empty(R, dtype=torch.float32, device=dev); st = torch.empty(R, dtype=torch.int32, device=dev); ls = torch.empty(R, dtype=torch.int32, device=dev)
s, ctx = lib.create_sinusoid_context(20, n)
nw = np.arange(21, dtype=np.uint64); co = torch.empty(R, 41, dtype=torch.float64, device=dev)
for _ in range(3):
    rc = lib.lib.sakura_LSQFitSinusoidFloatBatchDevice(ctx, 21, nw.ctypes.data, R, n, data.data_ptr(), n, mask.data_ptr(), n, 3.0, 3, 41, co.data_ptr(), None, res.data_ptr(), fm.data_ptr(), rms.data_ptr(), st.data_ptr(), ls.data_ptr(), None)
    assert rc == 0
torch.cuda.synchronize()
print("ok", int((st == 0).sum()))
