"""Development check: GPU and CPU-reference spline residuals against a float64 SVD least-squares
solution on the same final mask (the normal equations have cond ~ 2e11 at 20 pieces, so the two
implementations are compared with an independent, better conditioned solver)."""
import sys
sys.path.insert(0, "."); sys.path.insert(0, "tests")
import numpy as np
from sakura_b200 import synth
from sakura_b200.capi import SakuraLib
from oracle.refbind import Oracle
from parity import spline_truth_residual

lib = SakuraLib(); ora = Oracle()
n, npiece = 8192, 20
data, mask = synth.make_spectra(48, n, seed=61, ripple=True)
st, ctx = lib.create_cubicspline_context(npiece, n)
g = lib.lsqfit_cubicspline_batch(ctx, npiece, data, mask, 3.0, 3, want_best_fit=True)
o = ora.lsqfit_cubicspline_batch(npiece, data, mask, 3.0, 3, want_best_fit=True)
print("kernel", lib.last_kernel_name(), "mask equal", np.array_equal(g["final_mask"], o["final_mask"]))
eg = eo = ego = 0.0
for r in range(data.shape[0]):
    if o["status"][r] != 0: continue
    t = spline_truth_residual(data[r], o["final_mask"][r], o["boundary"][r], npiece)
    eg = max(eg, np.abs(g["residual"][r].astype(np.float64) - t).max())
    eo = max(eo, np.abs(o["residual"][r].astype(np.float64) - t).max())
    ego = max(ego, np.abs(g["residual"][r].astype(np.float64) - o["residual"][r]).max())
print(f"max |gpu - truth| {eg:.3e}  max |cpu ref - truth| {eo:.3e}  max |gpu - cpu ref| {ego:.3e}")
