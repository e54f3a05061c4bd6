"""How far do two legitimate builds of the REFERENCE differ from each other?  (CPU only.)

oracle/_ref/libsakura_ref.so       = the reference's Release flags (-mavx2 -mfma, contraction on)
oracle/_ref/libsakura_ref_nofma.so = the same translation units built as its SIMD_ARCH=AVX option (no FMA)

For the polynomial fit at orders 5, 6, 7 (monomial normal equations, cond(G) ~ 1.5e7 / 4e8 / 1e10) this
prints the largest difference of residual, best-fit model and coefficients between the two builds on the
synthetic rows of the GPU parity tests.  tests/test_gpu_polynomial.py::test_orders_fast_path widens its
bands for orders 6 and 7; this measurement is the justification."""
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle.refbind import Oracle  # noqa: E402
from sakura_b200 import synth  # noqa: E402


def spread(order, n, rows, seed, nfit=4, clip=3.0):
    a, b = Oracle("reference"), Oracle("reference_nofma")
    data, mask = synth.make_spectra(rows, n, seed=seed)
    synth.add_edge_case_rows(data, mask, order + 1)
    ra = a.lsqfit_polynomial_batch(order, data, mask, clip, nfit, want_best_fit=True)
    rb = b.lsqfit_polynomial_batch(order, data, mask, clip, nfit, want_best_fit=True)
    good = (ra["status"] == 0) & (rb["status"] == 0)
    same_mask = np.all(ra["final_mask"] == rb["final_mask"], axis=1) & good
    out = {"order": order, "n": n, "rows": int(good.sum()), "rows_with_equal_final_mask": int(same_mask.sum()),
           "status_equal": bool(np.array_equal(ra["status"], rb["status"]))}
    sel = same_mask
    dr = np.abs(ra["residual"][sel].astype(np.float64) - rb["residual"][sel])
    model = np.abs(data[sel].astype(np.float64) - ra["residual"][sel])
    out["residual_max_abs_diff"] = float(np.nanmax(dr))
    out["residual_max_diff_in_model_ulps"] = float(np.nanmax(dr / (1.2e-7 * np.maximum(model, 1e-30))))
    fm = ra["final_mask"][sel]
    out["residual_max_abs_diff_valid_channels"] = float(np.where(fm, dr, 0.0).max())
    scale = float(n - 1) ** np.arange(order + 1)
    ca, cb = ra["coeff"][sel] * scale, rb["coeff"][sel] * scale
    populated = fm.sum(axis=1) >= 8 * (order + 1)
    rel = np.abs(ca - cb)[populated] / np.abs(ca[populated]).max(axis=1, keepdims=True)
    out["coeff_max_diff_rel_to_row_norm"] = float(rel.max())
    with np.errstate(invalid="ignore", divide="ignore"):
        out["rms_max_rel_diff"] = float(np.nanmax(np.abs(ra["rms"][sel] - rb["rms"][sel]) / ra["rms"][sel]))
    return out


def main():
    res = [spread(order, n, 96, 10 + order) for order in (5, 6, 7) for n in (2048, 4096)]
    print(json.dumps(res, indent=1))


if __name__ == "__main__":
    main()
