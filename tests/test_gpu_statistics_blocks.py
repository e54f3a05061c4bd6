"""GPU statistics kernels and the standalone normal-equation building blocks vs the oracle and
the reference's known-answer vectors (test/statistics.cc, test/numeric_operation.cc)."""
import ctypes as C
import os

import numpy as np
import pytest

from sakura_b200 import synth
from sakura_b200.capi import STATUS_OK, aligned_copy, aligned_empty

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def check_stats(rec, d, m, ref=None):
    keep = d[m]
    assert rec["count"] == keep.size
    if keep.size == 0:
        assert rec["sum"] == 0 and rec["square_sum"] == 0
        assert np.isnan(rec["min"]) and np.isnan(rec["max"])
        assert rec["index_of_min"] == -1 and rec["index_of_max"] == -1
        return
    k64 = keep.astype(np.float64)
    assert np.isclose(rec["sum"], k64.sum(), rtol=1e-12, atol=1e-6 * np.abs(k64).max())
    assert np.isclose(rec["square_sum"], (k64 * k64).sum(), rtol=1e-12)
    assert rec["min"] == keep.min() and rec["max"] == keep.max()  # bit-exact
    # any valid occurrence is acceptable (sakura.h:250)
    assert m[rec["index_of_min"]] and d[rec["index_of_min"]] == rec["min"]
    assert m[rec["index_of_max"]] and d[rec["index_of_max"]] == rec["max"]
    if ref is not None:
        assert rec["count"] == ref["count"]
        assert np.isclose(rec["sum"], ref["sum"], rtol=1e-6, atol=1e-9)
        assert np.isclose(rec["square_sum"], ref["square_sum"], rtol=1e-6)
        assert rec["min"] == ref["min"] and rec["max"] == ref["max"]


@pytest.mark.parametrize("accurate", [True, False])
def test_known_answer_vectors(lib, accurate):
    for n in (256, 4, 1):
        d = np.arange(n, dtype=np.float32)
        m = np.ones(n, dtype=np.bool_)
        st, rec = lib.compute_statistics(d, m, accurate=accurate)
        assert st == STATUS_OK
        assert rec["sum"] == d.sum(dtype=np.float64) and rec["square_sum"] == (d.astype(np.float64) ** 2).sum()
        assert rec["index_of_min"] == 0 and rec["index_of_max"] == n - 1
        check_stats(rec, d, m)
    st, rec = lib.compute_statistics(np.zeros(0, np.float32), np.zeros(0, np.bool_), accurate=accurate)
    assert st == STATUS_OK and rec["count"] == 0 and np.isnan(rec["min"]) and rec["index_of_min"] == -1
    d = np.arange(1024, dtype=np.float32)
    st, rec = lib.compute_statistics(d, np.zeros(1024, np.bool_), accurate=accurate)
    check_stats(rec, d, np.zeros(1024, np.bool_))
    # a single valid element at every tail position (statistics.cc:444-461)
    for pos in list(range(0, 40)) + [1000, 1023]:
        m = np.zeros(1024, np.bool_)
        m[pos] = True
        st, rec = lib.compute_statistics(d, m, accurate=accurate)
        assert rec["count"] == 1 and rec["min"] == pos and rec["index_of_max"] == pos
    # NaN / Inf under the mask must not leak (statistics.cc:463-513)
    for bad in (np.nan, np.inf, -np.inf):
        d2 = d.copy()
        d2[[0, 7, 512, 1023]] = bad
        m2 = np.ones(1024, np.bool_)
        m2[[0, 7, 512, 1023]] = False
        st, rec = lib.compute_statistics(d2, m2, accurate=accurate)
        check_stats(rec, d2, m2)


def test_parity_with_oracle_and_golden(lib, oracle):
    rng = np.random.default_rng(8)
    for n in (15, 31, 32, 33, 1000, 4096, 70001):
        d = (rng.standard_normal(n) * 100).astype(np.float32)
        m = rng.random(n) > 0.3
        ref, _ = oracle.statistics_batch(d, m, accurate=True)
        st, rec = lib.compute_statistics(d, m, accurate=True)
        check_stats(rec, d, m, ref[0])
    z = np.load(os.path.join(GOLD, "statistics.npz"))
    for n in z["sizes"]:
        st, rec = lib.compute_statistics(z[f"data_{n}"], z[f"mask_{n}"], accurate=True)
        check_stats(rec, z[f"data_{n}"], z[f"mask_{n}"], z[f"acc_{n}"][0])


def test_batched_rows(lib, oracle):
    data, mask = synth.make_spectra(2048, 4096, seed=12)
    rec = lib.compute_statistics_batch(data, mask)
    ref, _ = oracle.statistics_batch(data, mask, accurate=True)
    assert np.array_equal(rec["count"], ref["count"])
    assert np.allclose(rec["sum"], ref["sum"], rtol=1e-6, atol=1e-6)
    assert np.allclose(rec["square_sum"], ref["square_sum"], rtol=1e-6)
    assert np.array_equal(rec["min"], ref["min"]) and np.array_equal(rec["max"], ref["max"])
    rows = np.arange(2048)
    assert np.array_equal(data[rows, rec["index_of_min"]], rec["min"])
    assert np.array_equal(data[rows, rec["index_of_max"]], rec["max"])


def test_accuracy_constants_81_92M(lib):
    # statistics.cc:594-643: bc-computed sums; the accurate API must match to 4 ulps
    N = 81920000
    d = np.arange(N, dtype=np.uint64).astype(np.float32) - np.float32(N + 8)
    m = (np.arange(N) % 2 == 1)
    st, rec = lib.compute_statistics(d, m, accurate=True)
    assert st == STATUS_OK
    assert rec["count"] == N // 2 and rec["sum"] == -1677721927680000.0
    expect = 91625995824881782489120.0
    assert abs(rec["square_sum"] - expect) <= 4 * np.spacing(expect)
    assert rec["index_of_min"] == 1 and rec["min"] == d[1] and rec["max"] == d[N - 3]


def _cheb(n, K):
    x = 2.0 * np.arange(n, dtype=np.float64) / (n - 1) - 1.0
    mdl = np.empty((n, K))
    mdl[:, 0] = 1.0
    mdl[:, 1] = x
    for i in range(2, K):
        mdl[:, i] = 2.0 * x * mdl[:, i - 1] - mdl[:, i - 2]
    return mdl


def test_building_blocks(lib, oracle):
    L = lib.lib
    p = lambda a: a.ctypes.data  # noqa: E731
    # SolveSimultaneousEquationsByLU, numeric_operation.cc:1915-1973
    n, K = 50, 5
    x = np.arange(n, dtype=np.float64)
    data = aligned_copy((1.0 + x + x * x + x ** 3 + x ** 4).astype(np.float32))
    mask = aligned_copy(np.ones(n, np.bool_))
    model = aligned_copy(np.vander(x, K, increasing=True))
    use = aligned_copy(np.arange(K, dtype=np.uint64))
    G, b, out = aligned_empty(K * K, np.float64), aligned_empty(K, np.float64), aligned_empty(K, np.float64)
    assert L.sakura_GetLSQCoefficientsDouble(n, p(data), p(mask), K, p(model), K, p(use), p(G), p(b)) == 0
    assert L.sakura_SolveSimultaneousEquationsByLUDouble(K, p(G), p(b), p(out)) == 0
    assert np.all((out - 1.0) / out <= 1e-7) and np.all(np.abs(out - 1.0) < 1e-6)
    # GetLSQCoefficients vs naive sums (numeric_operation.cc:498-607), Update vs rebuild (:1090-1191)
    n, K = 65536, 20
    xx = np.arange(n, dtype=np.float64)
    data = aligned_copy((4.0 + 0.000056 * xx - 0.000037 * xx ** 2 + 0.0000012 * xx ** 3
                         + 0.0000009 * xx ** 4 + 0.0000006 * xx ** 5).astype(np.float32))
    mask = aligned_copy(np.ones(n, np.bool_))
    model = aligned_copy(_cheb(n, K))
    use = aligned_copy(np.arange(K, dtype=np.uint64))
    G, b = aligned_empty(K * K, np.float64), aligned_empty(K, np.float64)
    assert L.sakura_GetLSQCoefficientsDouble(n, p(data), p(mask), K, p(model), K, p(use), p(G), p(b)) == 0
    tol = lambda a, e: np.all(np.abs(a - e) <= np.maximum(np.abs(a), np.abs(e)) * 1e-10 + 1e-10)  # noqa: E731
    assert tol(G.reshape(K, K), model.T @ model) and tol(b, model.T @ data.astype(np.float64))
    excl = aligned_copy(np.array([10, 500, 1000, 2047, 4000], dtype=np.uint64))
    assert L.sakura_UpdateLSQCoefficientsDouble(n, p(data), p(mask), 5, p(excl), K, p(model), K, p(use),
                                                p(G), p(b)) == 0
    m2 = mask.copy()
    m2[excl.astype(int)] = False
    m2 = aligned_copy(m2)
    G2, b2 = aligned_empty(K * K, np.float64), aligned_empty(K, np.float64)
    assert L.sakura_GetLSQCoefficientsDouble(n, p(data), p(m2), K, p(model), K, p(use), p(G2), p(b2)) == 0
    assert np.allclose(G, G2, rtol=1e-7, atol=1e-7) and np.allclose(b, b2, rtol=1e-7, atol=1e-7)
    # too few valid channels -> kNG (numeric_operation.cc:221-224)
    m3 = aligned_copy(np.zeros(n, np.bool_))
    m3[:3] = True
    assert L.sakura_GetLSQCoefficientsDouble(n, p(data), p(m3), K, p(model), K, p(use), p(G2), p(b2)) == 1
    # 499 unknowns (numeric_operation.cc:1980-2056): status only
    Kb = 499
    rng = np.random.default_rng(0)
    A = rng.standard_normal((Kb, Kb))
    A = aligned_copy(A @ A.T + Kb * np.eye(Kb))
    rhs = aligned_copy(rng.standard_normal(Kb))
    sol = aligned_empty(Kb, np.float64)
    assert L.sakura_SolveSimultaneousEquationsByLUDouble(Kb, p(A), p(rhs), p(sol)) == 0
    assert np.allclose(A @ sol, rhs, atol=1e-8)
