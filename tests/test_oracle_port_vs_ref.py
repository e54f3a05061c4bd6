"""The plain-C restatement (oracle/sakura_oracle.c) must reproduce the reference's own object
code (oracle/_ref) BIT FOR BIT on seeded inputs: status, masks, rms, residual, best_fit,
coefficients, spline boundaries, statistics."""
import numpy as np
import pytest

from sakura_b200 import synth


def assert_identical(a, b):
    for k in a:
        if a[k] is None or not isinstance(a[k], np.ndarray):
            continue
        if a[k].dtype.kind == "f":
            assert np.array_equal(a[k], b[k], equal_nan=True), k
        else:
            assert np.array_equal(a[k], b[k]), k


@pytest.mark.parametrize("order,nfit,n", [(5, 3, 4096), (5, 5, 1024), (0, 2, 64), (7, 4, 333), (2, 1, 15)])
def test_polynomial(oracle_ref, oracle_port, order, nfit, n):
    data, mask = synth.make_spectra(64, n, seed=1 + order)
    if n >= 4 * (order + 1):
        synth.add_edge_case_rows(data, mask, order + 1)
    a = oracle_ref.lsqfit_polynomial_batch(order, data, mask, 3.0, nfit, want_best_fit=True)
    b = oracle_port.lsqfit_polynomial_batch(order, data, mask, 3.0, nfit, want_best_fit=True)
    assert_identical(a, b)


def test_chebyshev(oracle_ref, oracle_port):
    data, mask = synth.make_spectra(32, 2048, seed=7)
    a = oracle_ref.lsqfit_polynomial_batch(8, data, mask, 2.5, 4, lsqfit_type=1, want_best_fit=True)
    b = oracle_port.lsqfit_polynomial_batch(8, data, mask, 2.5, 4, lsqfit_type=1, want_best_fit=True)
    assert_identical(a, b)


@pytest.mark.parametrize("npiece,n", [(20, 8192), (3, 100), (1, 64)])
def test_cubic_spline(oracle_ref, oracle_port, npiece, n):
    data, mask = synth.make_spectra(24, n, seed=11, ripple=True)
    synth.add_edge_case_rows(data, mask, npiece + 3)
    a = oracle_ref.lsqfit_cubicspline_batch(npiece, data, mask, 3.0, 3, want_best_fit=True)
    b = oracle_port.lsqfit_cubicspline_batch(npiece, data, mask, 3.0, 3, want_best_fit=True)
    assert_identical(a, b)


@pytest.mark.parametrize("nwave", [list(range(21)), [1, 3, 4], [0]])
def test_sinusoid(oracle_ref, oracle_port, nwave):
    data, mask = synth.make_spectra(24, 2048, seed=13, ripple=True)
    a = oracle_ref.lsqfit_sinusoid_batch(nwave, data, mask, 3.0, 3, context_nwave=20, want_best_fit=True)
    b = oracle_port.lsqfit_sinusoid_batch(nwave, data, mask, 3.0, 3, context_nwave=20, want_best_fit=True)
    assert_identical(a, b)


@pytest.mark.parametrize("accurate", [True, False])
def test_statistics(oracle_ref, oracle_port, accurate):
    rng = np.random.default_rng(5)
    for n in (1, 4, 15, 31, 32, 33, 100, 256, 257, 1000, 4096, 5000, 70001):
        d = (rng.standard_normal(n) * 100).astype(np.float32)
        m = rng.random(n) > 0.3
        ra, _ = oracle_ref.statistics_batch(d, m, accurate=accurate)
        rb, _ = oracle_port.statistics_batch(d, m, accurate=accurate)
        for f in ra.dtype.names:
            assert (ra[f][0] == rb[f][0]) or (np.isnan(ra[f][0]) and np.isnan(rb[f][0])), (n, f)
