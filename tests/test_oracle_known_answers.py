"""Pins the oracle (both the reference's own object code and the C restatement) against the
known-answer tests the reference itself holds for this path, re-expressed without gtest:

* test/numeric_operation.cc:1915-1973  SolveSimultaneousEquationsByLU (50-pt quartic -> 1.0)
* test/numeric_operation.cc:498-607    GetLSQCoefficients vs naive sums, Chebyshev 65536 x 20
* test/numeric_operation.cc:1090-1191  UpdateLSQCoefficients vs full rebuild
* test/lsq.cc:1255-1410, 2096-2157     default ParamSets: coeff {10,0,..}, best_fit=data, residual=0
* test/lsq.cc:1441-1560                order 99 / 99 pieces / nwave 0..50 "performance" shapes
* test/lsq.cc:474-488, 1859-1922       effdataLB / effdataTL / noeffdata / clip_threshold_sigmaTL
* test/statistics.cc:294-700           iota vectors, all-invalid, NaN under mask, accuracy constants
* python-binding/test3.py:448-465      8 points, slope 0.5, masked +3 outlier
"""
import ctypes as C
import os

import numpy as np
import pytest

from oracle import refbind
from oracle.refbind import aligned_copy, aligned_empty

KINDS = ["port", "reference"]


@pytest.fixture(params=KINDS)
def ora(request):
    if request.param == "reference" and not os.path.exists(refbind.REF_SO):
        pytest.skip("reference library not built")
    if request.param == "port" and not os.path.exists(refbind.PORT_SO):
        refbind.build()
    return refbind.Oracle(request.param)


def almost(a, b, tol=1e-5):
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    return np.all(np.abs(a - b) <= np.maximum(np.abs(a), np.abs(b)) * tol + tol)  # lsq.cc:2096-2101


# ---- building blocks ---------------------------------------------------------------------
def _blocks(ora):
    """(get, update, solve) callables over aligned numpy arrays for either oracle kind."""
    vp, sz = C.c_void_p, C.c_size_t
    L = ora.lib
    if ora.kind == "reference":
        get, upd, sol = (L.sakura_GetLSQCoefficientsDouble, L.sakura_UpdateLSQCoefficientsDouble,
                         L.sakura_SolveSimultaneousEquationsByLUDouble)
    else:
        get, upd, sol = L.oracle_get_lsq_coefficients, L.oracle_update_lsq_coefficients, L.oracle_solve_by_lu
    get.argtypes = [sz, vp, vp, sz, vp, sz, vp, vp, vp]
    upd.argtypes = [sz, vp, vp, sz, vp, sz, vp, sz, vp, vp, vp]
    sol.argtypes = [sz, vp, vp, vp]
    for f in (get, upd, sol):
        f.restype = C.c_int
    p = lambda a: C.c_void_p(a.ctypes.data)  # noqa: E731
    return get, upd, sol, p


def test_solve_by_lu_quartic(ora):
    get, _, sol, p = _blocks(ora)
    n, K = 50, 5
    x = np.arange(n, dtype=np.float64)
    data = aligned_copy((1.0 + x + x * x + x ** 3 + x ** 4).astype(np.float32))
    mask = aligned_copy(np.ones(n, dtype=np.bool_))
    model = aligned_copy(np.vander(x, K, increasing=True))
    use = aligned_copy(np.arange(K, dtype=np.uint64))
    G = aligned_empty(K * K, np.float64)
    b = aligned_empty(K, np.float64)
    out = aligned_empty(K, np.float64)
    assert get(n, p(data), p(mask), K, p(model), K, p(use), p(G), p(b)) == 0
    assert sol(K, p(G), p(b), p(out)) == 0
    assert np.all((out - 1.0) / out <= 1e-7)  # numeric_operation.cc:1964-1967 (signed)
    assert np.all(np.abs(out - 1.0) <= 1e-6)


def _cheb_model(n, K):
    x = 2.0 * np.arange(n, dtype=np.float64) / (n - 1) - 1.0
    m = np.empty((n, K))
    m[:, 0] = 1.0
    m[:, 1] = x
    for i in range(2, K):
        m[:, i] = 2.0 * x * m[:, i - 1] - m[:, i - 2]
    return m


def _quintic(n):
    x = np.arange(n, dtype=np.float64)
    return (4.0 + 0.000056 * x - 0.000037 * x * x + 0.0000012 * x ** 3 + 0.0000009 * x ** 4
            + 0.0000006 * x ** 5).astype(np.float32)  # numeric_operation.cc:139-147


def test_get_lsq_coefficients_vs_naive(ora):
    get, _, _, p = _blocks(ora)
    n, K = 65536, 20
    data = aligned_copy(_quintic(n))
    mask = aligned_copy(np.ones(n, dtype=np.bool_))
    model = aligned_copy(_cheb_model(n, K))
    use = aligned_copy(np.arange(K, dtype=np.uint64))
    G = aligned_empty(K * K, np.float64)
    b = aligned_empty(K, np.float64)
    assert get(n, p(data), p(mask), K, p(model), K, p(use), p(G), p(b)) == 0
    Gref = model.T @ model
    bref = model.T @ data.astype(np.float64)
    assert almost(G.reshape(K, K), Gref, 1e-10)
    assert almost(b, bref, 1e-10)


def test_update_lsq_coefficients_vs_rebuild(ora):
    get, upd, _, p = _blocks(ora)
    n, K = 4096, 20
    data = aligned_copy(_quintic(n))
    mask = aligned_copy(np.ones(n, dtype=np.bool_))
    model = aligned_copy(_cheb_model(n, K))
    use = aligned_copy(np.arange(K, dtype=np.uint64))
    G = aligned_empty(K * K, np.float64)
    b = aligned_empty(K, np.float64)
    assert get(n, p(data), p(mask), K, p(model), K, p(use), p(G), p(b)) == 0
    excl = aligned_copy(np.array([10, 500, 1000, 2047, 4000], dtype=np.uint64))
    assert upd(n, p(data), p(mask), len(excl), p(excl), K, p(model), K, p(use), p(G), p(b)) == 0
    m2 = mask.copy()
    m2[excl.astype(int)] = False
    m2 = aligned_copy(m2)
    G2 = aligned_empty(K * K, np.float64)
    b2 = aligned_empty(K, np.float64)
    assert get(n, p(data), p(m2), K, p(model), K, p(use), p(G2), p(b2)) == 0
    assert almost(G, G2, 1e-7)
    assert almost(b, b2, 1e-7)


# ---- fits: default parameter sets (lsq.cc:1255-1410) ---------------------------------------
N15 = 15


def _const(n=N15, v=10.0):
    return np.full(n, v, dtype=np.float32), np.ones(n, dtype=np.bool_)


def _expect_const_fit(r, n, ncoef):
    assert r["status"][0] == 0 and r["lsqfit_status"][0] == 0
    c = r["coeff"].reshape(-1)
    expect = np.zeros_like(c)
    expect[0] = 10.0
    assert almost(c, expect)
    assert almost(r["best_fit"][0], np.full(n, 10.0))
    assert almost(r["residual"][0], np.zeros(n))
    assert np.all(r["final_mask"][0])


def test_default_polynomial(ora):
    d, m = _const()
    r = ora.lsqfit_polynomial_batch(3, d, m, 3.0, 2, want_best_fit=True)
    _expect_const_fit(r, N15, 4)
    r = ora.lsqfit_polynomial_batch(3, d, m, 3.0, 2, lsqfit_type=1, want_best_fit=True)  # contextVP chebyshev
    _expect_const_fit(r, N15, 4)
    r = ora.lsqfit_polynomial_batch(0, d, m, 3.0, 2, context_order=3, want_best_fit=True)  # orderLB
    _expect_const_fit(r, N15, 1)


def test_default_cubicspline(ora):
    d, m = _const()
    r = ora.lsqfit_cubicspline_batch(3, d, m, 3.0, 2, want_best_fit=True)
    assert r["status"][0] == 0
    c = r["coeff"][0]
    assert almost(c[:, 0], np.full(3, 10.0)) and almost(c[:, 1:], np.zeros((3, 3)))
    assert almost(r["best_fit"][0], np.full(N15, 10.0)) and almost(r["residual"][0], np.zeros(N15))
    assert r["boundary"][0, 0] == 0 and r["boundary"][0, -1] == N15
    r1 = ora.lsqfit_cubicspline_batch(1, d, m, 3.0, 2, context_npiece=3, want_best_fit=True)  # num_piecesLB
    assert r1["status"][0] == 0 and almost(r1["coeff"][0][0], [10.0, 0, 0, 0])


def test_default_sinusoid(ora):
    d, m = _const()
    r = ora.lsqfit_sinusoid_batch([0, 1, 2], d, m, 3.0, 2, context_nwave=3, want_best_fit=True)
    _expect_const_fit(r, N15, 5)


def test_performance_shapes_values(ora):
    # lsq.cc:1441-1457 order 99 on 100 points; :1479-1495 99 pieces on 102; :1519-1539 nwave 0..50 on 102
    d, m = _const(100)
    r = ora.lsqfit_polynomial_batch(99, d, m, 3.0, 1, want_best_fit=True)
    _expect_const_fit(r, 100, 100)
    d, m = _const(102)
    r = ora.lsqfit_cubicspline_batch(99, d, m, 3.0, 1, want_best_fit=True)
    assert r["status"][0] == 0
    assert almost(r["best_fit"][0], np.full(102, 10.0)) and almost(r["residual"][0], np.zeros(102))
    assert almost(r["coeff"][0][:, 0], np.full(99, 10.0))
    r = ora.lsqfit_sinusoid_batch(list(range(51)), d, m, 3.0, 1, want_best_fit=True)
    _expect_const_fit(r, 102, 101)


def test_effective_data_cases(ora):
    d, _ = _const()
    m = np.zeros(N15, dtype=np.bool_)
    m[:4] = True  # effdataLB: exactly order+1 valid
    r = ora.lsqfit_polynomial_batch(3, d, m, 3.0, 2)
    assert r["status"][0] == 0 and r["lsqfit_status"][0] == 0
    m[3] = False  # effdataTL
    r = ora.lsqfit_polynomial_batch(3, d, m, 3.0, 2)
    assert r["status"][0] == 1 and r["lsqfit_status"][0] == 2
    assert np.all(np.isnan(r["coeff"])) and not r["final_mask"].any()  # nothing written
    m[:] = False  # noeffdata
    r = ora.lsqfit_polynomial_batch(3, d, m, 3.0, 2)
    assert r["status"][0] == 1 and r["lsqfit_status"][0] == 2


def test_clip_threshold_sigma_too_small(ora):
    # lsq.cc:486-488, 1918-1922: sigma=0.1, 5 fits, data[0]=1000 -> clipping eats the row -> kNG
    d, m = _const()
    d[0] = 1000.0
    r = ora.lsqfit_polynomial_batch(3, d, m, 0.1, 5)
    assert r["status"][0] == 1 and r["lsqfit_status"][0] == 2
    assert np.all(np.isnan(r["coeff"])) and np.isnan(r["rms"][0])
    assert r["final_mask"][0].sum() < 4  # final_mask was already clipped in place


def test_num_fitting_max_zero_writes_nothing(ora):
    d, m = _const()
    r = ora.lsqfit_polynomial_batch(3, d, m, 3.0, 0)
    assert r["status"][0] == 0 and r["lsqfit_status"][0] == 0
    assert np.all(np.isnan(r["coeff"])) and np.isnan(r["rms"][0]) and not r["final_mask"].any()


def test_python_binding_example(ora):
    # python-binding/test3.py:448-465
    y = np.array([0.5 * i for i in range(8)], dtype=np.float32)
    m = np.ones(8, dtype=np.bool_)
    y[4] += 3.0
    m[4] = False
    r = ora.lsqfit_polynomial_batch(1, y, m, 5.0, 1, want_best_fit=True)
    assert r["status"][0] == 0
    assert np.allclose(r["residual"][0], [0, 0, 0, 0, 3, 0, 0, 0], atol=1e-6)
    assert np.allclose(r["coeff"][0], [0.0, 0.5], atol=1e-6)


# ---- statistics (test/statistics.cc) --------------------------------------------------------
@pytest.mark.parametrize("accurate", [True, False])
def test_statistics_iota(ora, accurate):
    for n in (256, 4, 1):
        d = np.arange(n, dtype=np.float32)
        m = np.ones(n, dtype=np.bool_)
        r, st = ora.statistics_batch(d, m, accurate=accurate)
        assert st[0] == 0
        assert r["count"][0] == n and r["sum"][0] == d.sum(dtype=np.float64)
        assert r["square_sum"][0] == (d.astype(np.float64) ** 2).sum()
        assert r["min"][0] == 0.0 and r["max"][0] == n - 1
        if n >= 32 or not accurate:  # the accurate variant mis-indexes short tails (documented quirk)
            assert r["index_of_min"][0] == 0 and r["index_of_max"][0] == n - 1


@pytest.mark.parametrize("accurate", [True, False])
def test_statistics_all_invalid_and_nan_under_mask(ora, accurate):
    n = 1024
    d = np.arange(n, dtype=np.float32)
    r, _ = ora.statistics_batch(d, np.zeros(n, dtype=np.bool_), accurate=accurate)
    assert r["count"][0] == 0 and r["sum"][0] == 0 and r["square_sum"][0] == 0
    assert np.isnan(r["min"][0]) and np.isnan(r["max"][0])
    assert r["index_of_min"][0] == -1 and r["index_of_max"][0] == -1
    m = np.ones(n, dtype=np.bool_)
    for bad in (np.nan, np.inf, -np.inf):
        d2 = d.copy()
        d2[[3, 500, 1023]] = bad
        m2 = m.copy()
        m2[[3, 500, 1023]] = False
        r, _ = ora.statistics_batch(d2, m2, accurate=accurate)
        keep = d[m2].astype(np.float64)
        assert r["count"][0] == keep.size and r["sum"][0] == keep.sum()
        assert r["min"][0] == keep.min() and r["max"][0] == keep.max()


def test_statistics_accuracy_constants(ora):
    # statistics.cc:594-700: data[i] = i-(N+8), odd i valid, N = 81,920,000 (bc-computed answers);
    # the accurate variant must match to EXPECT_DOUBLE_EQ (4 ulps)
    N = 81920000
    # `data[i] = i - float(data.size() + 8)`: size_t converted to float, float subtraction
    d = np.arange(N, dtype=np.uint64).astype(np.float32) - np.float32(N + 8)
    m = (np.arange(N) % 2 == 1)
    r, _ = ora.statistics_batch(d, m, accurate=True)
    assert r["count"][0] == N // 2
    assert r["sum"][0] == -1677721927680000.0
    expect_sq = 91625995824881782489120.0
    assert abs(r["square_sum"][0] - expect_sq) <= 4 * np.spacing(expect_sq)
