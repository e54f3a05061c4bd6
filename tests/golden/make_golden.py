"""Generates tests/golden/*.npz from the REFERENCE ITSELF (oracle/_ref/libsakura_ref.so, i.e.
the reference's own object code built from /root/reference by oracle/Makefile).

    python tests/golden/make_golden.py

Run in the build container (where /root/reference exists); the .npz files are committed so
that the oracle port and the GPU path stay pinned to the reference on boxes without it.
Inputs are stored next to the outputs, so the fixtures do not depend on any RNG.
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

from oracle.refbind import Oracle  # noqa: E402
from sakura_b200 import synth  # noqa: E402


def pack(prefix, d):
    return {f"{prefix}{k}": v for k, v in d.items() if v is not None}


def main():
    ref = Oracle("reference")
    out = {}
    # polynomial order 5, 3 fits, 3 sigma (config C1 shape, reduced)
    data, mask = synth.make_spectra(48, 512, seed=101)
    synth.add_edge_case_rows(data, mask, 6)
    r = ref.lsqfit_polynomial_batch(5, data, mask, 3.0, 3, want_best_fit=True)
    np.savez_compressed(os.path.join(HERE, "poly5.npz"), data=data, mask=mask, order=5, nfit=3,
                        clip=3.0, **pack("out_", r))
    # Chebyshev order 4, 5 fits
    data, mask = synth.make_spectra(16, 256, seed=102)
    r = ref.lsqfit_polynomial_batch(4, data, mask, 2.5, 5, lsqfit_type=1, want_best_fit=True)
    np.savez_compressed(os.path.join(HERE, "cheb4.npz"), data=data, mask=mask, order=4, nfit=5,
                        clip=2.5, **pack("out_", r))
    # cubic spline, 6 pieces
    data, mask = synth.make_spectra(16, 512, seed=103, ripple=True)
    synth.add_edge_case_rows(data, mask, 9)
    r = ref.lsqfit_cubicspline_batch(6, data, mask, 3.0, 3, want_best_fit=True)
    np.savez_compressed(os.path.join(HERE, "spline6.npz"), data=data, mask=mask, npiece=6, nfit=3,
                        clip=3.0, **pack("out_", r))
    # sinusoid, wave numbers 0..5
    data, mask = synth.make_spectra(16, 512, seed=104, ripple=True)
    synth.add_edge_case_rows(data, mask, 11)
    r = ref.lsqfit_sinusoid_batch(list(range(6)), data, mask, 3.0, 3, want_best_fit=True)
    np.savez_compressed(os.path.join(HERE, "sinusoid5.npz"), data=data, mask=mask, nwave=np.arange(6),
                        nfit=3, clip=3.0, **pack("out_", r))
    # statistics
    rng = np.random.Generator(np.random.Philox(105))
    rows = []
    for n in (1, 7, 8, 31, 32, 33, 255, 256, 1000, 4096):
        d = (rng.standard_normal(n) * 50).astype(np.float32)
        m = rng.random(n) > 0.25
        acc, _ = ref.statistics_batch(d, m, accurate=True)
        fast, _ = ref.statistics_batch(d, m, accurate=False)
        rows.append((n, d, m, acc[0], fast[0]))
    np.savez_compressed(os.path.join(HERE, "statistics.npz"),
                        sizes=np.array([r[0] for r in rows]),
                        **{f"data_{r[0]}": r[1] for r in rows}, **{f"mask_{r[0]}": r[2] for r in rows},
                        **{f"acc_{r[0]}": np.array([r[3]]) for r in rows},
                        **{f"fast_{r[0]}": np.array([r[4]]) for r in rows})
    print("golden fixtures written to", HERE)


if __name__ == "__main__":
    main()
