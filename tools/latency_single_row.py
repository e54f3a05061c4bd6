"""Latency of the reference-signature single-row entry points (sakura_LSQFit*Float, host pointers) on
the GPU library, next to the reference's own CPU code per row on one host core.

    python tools/latency_single_row.py [--calls 300]

Two host paths are timed: the single-arena path (one pinned arena, one upload, one download, one
synchronisation per call: the default for small batches) and the staged path it replaced (separate
copies out of pageable memory through stream 0, two synchronisations)."""
import argparse
import ctypes as C
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from sakura_b200 import synth  # noqa: E402
from sakura_b200.capi import SakuraLib, aligned_copy  # noqa: E402


def time_calls(fn, calls):
    for _ in range(10):
        fn()
    t = np.empty(calls)
    for i in range(calls):
        t0 = time.perf_counter()
        fn()
        t[i] = time.perf_counter() - t0
    return {"median_us": float(np.median(t) * 1e6), "mean_us": float(t.mean() * 1e6), "p95_us": float(np.percentile(t, 95) * 1e6)}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--calls", type=int, default=300)
    args = ap.parse_args()
    lib = SakuraLib()
    out = {}
    for n, nfit in ((4096, 5), (4096, 1), (1024, 5), (8192, 3)):
        data, mask = synth.make_spectra(64, n, seed=3, ripple=(n == 8192))
        d, m = aligned_copy(data[0]), aligned_copy(mask[0])
        coeff = aligned_copy(np.zeros(64))
        resid, fm = aligned_copy(np.zeros(n, np.float32)), aligned_copy(np.zeros(n, np.bool_))
        rms, lsq = C.c_float(), C.c_int()
        bound = aligned_copy(np.zeros(32, np.uint64))
        nw = aligned_copy(np.arange(21, dtype=np.uint64))
        st, cp = lib.create_polynomial_context(5, n)
        st, cs = lib.create_cubicspline_context(20, n)
        st, cw = lib.create_sinusoid_context(20, n)
        calls = {
            "polynomial order 5": lambda: lib.lib.sakura_LSQFitPolynomialFloat(
                cp, 5, n, d.ctypes.data, m.ctypes.data, 3.0, nfit, 6, coeff.ctypes.data, None, resid.ctypes.data,
                fm.ctypes.data, C.byref(rms), C.byref(lsq)),
        }
        if n == 8192:
            calls["cubic spline 20 pieces"] = lambda: lib.lib.sakura_LSQFitCubicSplineFloat(
                cs, 20, n, d.ctypes.data, m.ctypes.data, 3.0, nfit, coeff.ctypes.data, None, resid.ctypes.data,
                fm.ctypes.data, C.byref(rms), bound.ctypes.data, C.byref(lsq))
            calls["sinusoid nwave 0..20"] = lambda: lib.lib.sakura_LSQFitSinusoidFloat(
                cw, 21, nw.ctypes.data, n, d.ctypes.data, m.ctypes.data, 3.0, nfit, 41, coeff.ctypes.data, None,
                resid.ctypes.data, fm.ctypes.data, C.byref(rms), C.byref(lsq))
        for name, fn in calls.items():
            assert fn() == 0, lib.last_error()
            row = {}
            lib.set_host_pipeline_tuning(96, 2048)
            row["single_arena_path"] = time_calls(fn, args.calls)
            lib.set_host_pipeline_tuning(96, 0)
            row["staged_path"] = time_calls(fn, args.calls)
            lib.set_host_pipeline_tuning(96, 2048)
            row["kernel"] = lib.last_kernel_name()
            out[f"{name}, {n} ch, num_fitting_max {nfit}"] = row
        # the reference's own code, one thread, per row
        try:
            from oracle.refbind import Oracle
            ora = Oracle()
            ora.lsqfit_polynomial_batch(5, data, mask, 3.0, nfit, nthreads=1)
            t0 = time.perf_counter()
            ora.lsqfit_polynomial_batch(5, data, mask, 3.0, nfit, nthreads=1)
            out[f"polynomial order 5, {n} ch, num_fitting_max {nfit}"]["cpu_reference_us_per_row"] = \
                (time.perf_counter() - t0) / data.shape[0] * 1e6
            if n == 8192:
                t0 = time.perf_counter()
                ora.lsqfit_cubicspline_batch(20, data[:16], mask[:16], 3.0, nfit, nthreads=1)
                out[f"cubic spline 20 pieces, {n} ch, num_fitting_max {nfit}"]["cpu_reference_us_per_row"] = \
                    (time.perf_counter() - t0) / 16 * 1e6
                t0 = time.perf_counter()
                ora.lsqfit_sinusoid_batch(list(range(21)), data[:16], mask[:16], 3.0, nfit, nthreads=1)
                out[f"sinusoid nwave 0..20, {n} ch, num_fitting_max {nfit}"]["cpu_reference_us_per_row"] = \
                    (time.perf_counter() - t0) / 16 * 1e6
        except OSError as exc:
            out["cpu_reference"] = str(exc)
        for c in (cp, cs, cw):
            lib.destroy_context(c)
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
