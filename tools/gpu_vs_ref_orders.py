"""GPU vs reference at polynomial orders 5..7 next to the reference-vs-reference spread (tools/ref_vs_ref_spread.py)."""
import json, os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle.refbind import Oracle
from sakura_b200 import synth
from sakura_b200.capi import SakuraLib
lib = SakuraLib(); ref = Oracle("reference")
res = []
for kernel in ("warp", "channel"):
    os.environ["SAKURA_B200_POLY_KERNEL"] = kernel
    for order in (5, 6, 7):
        for n in (2048, 4096):
            data, mask = synth.make_spectra(96, n, seed=10 + order)
            synth.add_edge_case_rows(data, mask, order + 1)
            st, ctx = lib.create_polynomial_context(order, n)
            g = lib.lsqfit_polynomial_batch(ctx, order, data, mask, 3.0, 4, want_best_fit=True)
            lib.destroy_context(ctx)
            o = ref.lsqfit_polynomial_batch(order, data, mask, 3.0, 4, want_best_fit=True)
            good = (g["status"] == 0) & (o["status"] == 0)
            sel = good & np.all(g["final_mask"] == o["final_mask"], axis=1)
            dr = np.abs(g["residual"][sel].astype(np.float64) - o["residual"][sel])
            fm = o["final_mask"][sel]
            scale = float(n - 1) ** np.arange(order + 1)
            ca, cb = g["coeff"][sel] * scale, o["coeff"][sel] * scale
            pop = fm.sum(axis=1) >= 8 * (order + 1)
            rel = np.abs(ca - cb)[pop] / np.abs(cb[pop]).max(axis=1, keepdims=True)
            res.append({"kernel": lib.last_kernel_name(), "order": order, "n": n, "rows_equal_mask": int(sel.sum()), "rows": int(good.sum()),
                        "residual_max_abs_diff": float(np.nanmax(dr)), "residual_max_abs_diff_valid_channels": float(np.where(fm, dr, 0.0).max()),
                        "coeff_max_diff_rel_to_row_norm": float(rel.max())})
            print(res[-1], flush=True)
