"""GPU parity tests of the polynomial hot path (the C ABI on a B200) against the oracle on
identical seeded inputs, plus the reference's API-contract cases (libsakura/test/lsq.cc)."""
import ctypes as C
import os

import numpy as np
import pytest

from sakura_b200 import synth
from sakura_b200.capi import (LSQFIT_NOT_ENOUGH_DATA, LSQFIT_OK, LSQFIT_TYPE_CHEBYSHEV, STATUS_NG,
                              STATUS_OK, aligned_copy, aligned_empty)
from parity import compare_fit, forced_kernel

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def run_both(lib, oracle, order, data, mask, clip, nfit, lsqfit_type=0, ctx_order=None, **want):
    st, ctx = lib.create_polynomial_context(order if ctx_order is None else ctx_order, data.shape[-1],
                                            lsqfit_type=lsqfit_type)
    assert st == STATUS_OK
    g = lib.lsqfit_polynomial_batch(ctx, order, data, mask, clip, nfit, **want)
    kernel = lib.last_kernel_name()
    lib.destroy_context(ctx)
    o = oracle.lsqfit_polynomial_batch(order, data, mask, clip, nfit, lsqfit_type=lsqfit_type,
                                       context_order=ctx_order, **{k: v for k, v in want.items()})
    return g, o, kernel


def test_config_c1_1024x4096_poly5_3fits(lib, oracle):
    """BASELINE configs[0]: 1,024 x 4,096, order 5, 3 clip iterations at 3 sigma."""
    data, mask = synth.make_spectra(1024, 4096)
    names = synth.add_edge_case_rows(data, mask, 6)
    g, o, kernel = run_both(lib, oracle, 5, data, mask, 3.0, 3, want_best_fit=True)
    assert kernel == "poly_warp"  # the warp-per-spectrum kernel from two fits on
    reported = compare_fit(g, o, data, 3.0, poly_rescale_n=4096)
    assert reported == [], reported
    assert g["status"][names["K_minus_1"]] == STATUS_NG
    assert g["lsqfit_status"][names["all_false"]] == LSQFIT_NOT_ENOUGH_DATA
    assert g["status"][names["exactly_K"]] == STATUS_OK
    assert np.isfinite(g["residual"][names["nan_under_mask"]]).sum() == 4095  # NaN did not leak


@pytest.mark.parametrize("nfit", [1, 2, 5, 10])
def test_iteration_counts(lib, oracle, nfit):
    data, mask = synth.make_spectra(256, 4096, seed=3)
    g, o, _ = run_both(lib, oracle, 5, data, mask, 3.0, nfit)
    compare_fit(g, o, data, 3.0, poly_rescale_n=4096)


@pytest.mark.parametrize("order", [0, 1, 2, 3, 4, 6, 7])
def test_orders_fast_path(lib, oracle, order):
    data, mask = synth.make_spectra(96, 2048, seed=10 + order)
    synth.add_edge_case_rows(data, mask, order + 1)
    g, o, kernel = run_both(lib, oracle, order, data, mask, 3.0, 4, want_best_fit=True)
    assert kernel == "poly_warp"
    # cond(G) grows ~30x per order in the monomial basis (1.5e7 at order 5, 4e8 at 6, ~1e10 at 7): O(eps)
    # differences in the normal equations move the model by cond*eps, above all across flagged gaps.
    # Measured on these rows (profiles/r02/ref_vs_ref_orders.json, gpu_vs_ref_orders.txt), residual /
    # coefficients:
    #   order 6: two builds of the REFERENCE ITSELF (FMA / no FMA) 1.9e-6 / 1.8e-8, GPU vs reference
    #            3.8e-6 / 1.3e-7 (inside the 1e-5 coefficient gate);
    #   order 7: reference vs reference 1.2e-4 / 1.7e-6; this library's general kernel -- the reference's
    #            own table and arithmetic, sums grouped differently -- 7.4e-4; the moment form 1.3e-3 /
    #            1.8e-5 (2.4e-5 on valid channels).
    # At order 7 nothing short of the reference's exact summation order reproduces it to 1e-5; the bands
    # below are 2.5 x the measured deviation. DESIGN.md section 4.
    compare_fit(g, o, data, 3.0, poly_rescale_n=2048, resid_atol={6: 1e-5, 7: 3.3e-3}.get(order, 0.0),
                coeff_rtol={7: 4.5e-5}.get(order, 1e-5))


@pytest.mark.parametrize("n", [16, 64, 100, 333, 1000, 2048, 4096, 6000, 8192, 10000])
def test_channel_counts(lib, oracle, n):
    """Fast path for n % 4 == 0 and n <= 8192 (32-byte aligned rows need n % 32 == 0 in a tight
    host batch, so other sizes go row by row), general kernel otherwise."""
    order = 3
    data, mask = synth.make_spectra(8, n, seed=n)
    if n % 32 == 0:
        g, o, kernel = run_both(lib, oracle, order, data, mask, 3.0, 3)
        assert kernel == ("poly_warp" if n in (2048, 4096) else "poly_screen" if n == 8192 else "poly_fast" if n <= 8192 else "general")
        compare_fit(g, o, data, 3.0, poly_rescale_n=n)
    else:
        st, ctx = lib.create_polynomial_context(order, n)
        for r in range(data.shape[0]):
            g = lib.lsqfit_polynomial_batch(ctx, order, data[r], mask[r], 3.0, 3)
            o = oracle.lsqfit_polynomial_batch(order, data[r], mask[r], 3.0, 3)
            compare_fit(g, o, data[r:r + 1], 3.0, poly_rescale_n=n)
        lib.destroy_context(ctx)


def test_clip_sigma_variants(lib, oracle):
    data, mask = synth.make_spectra(128, 1024, seed=21)
    for clip in (1.5, 2.0, 5.0, 3.4e38):
        g, o, _ = run_both(lib, oracle, 4, data, mask, clip, 4)
        compare_fit(g, o, data, min(clip, 1e30), poly_rescale_n=1024)


def test_chebyshev_and_high_order_general_kernel(lib, oracle):
    data, mask = synth.make_spectra(64, 1024, seed=5)
    g, o, kernel = run_both(lib, oracle, 5, data, mask, 3.0, 3, lsqfit_type=LSQFIT_TYPE_CHEBYSHEV,
                            want_best_fit=True)
    # Chebyshev up to order 5 is fitted by the polynomial kernel (same model space) and only the
    # coefficients are converted to the Chebyshev basis
    assert kernel == "poly_fast"
    compare_fit(g, o, data, 3.0)
    for order in (0, 1, 3, 4, 7):
        g, o, kernel = run_both(lib, oracle, order, data, mask, 3.0, 4, lsqfit_type=LSQFIT_TYPE_CHEBYSHEV,
                                want_best_fit=True)
        # orders 6..47: Gram matrix from masked sums of T_m (T_j T_k = (T_{j+k} + T_{|j-k|}) / 2) in
        # the tensor-core kernel of the sinusoid fit
        assert kernel == ("poly_fast" if order <= 5 else "chebyshev_dmma")
        compare_fit(g, o, data, 3.0)
    # well-conditioned high order: Chebyshev 20 (the monomial basis is numerically rank deficient
    # beyond order ~10, where the rank-truncated solution depends on the last bit of G)
    nogap, nogap_mask = synth.make_spectra(64, 1024, seed=6, max_block=0)  # no wide flagged gap
    g, o, kernel = run_both(lib, oracle, 20, nogap, nogap_mask, 3.0, 3, lsqfit_type=LSQFIT_TYPE_CHEBYSHEV,
                            want_best_fit=True)
    assert kernel == "chebyshev_dmma"
    compare_fit(g, o, nogap, 3.0, resid_atol=1e-4, coeff_rtol=1e-3)
    g, o, kernel = run_both(lib, oracle, 8, nogap, nogap_mask, 3.0, 3, want_best_fit=True)
    assert kernel == "general"
    # monomial order 8: cond(G) ~ 1e12, so cond*eps ~ 1e-4 relative on the coefficients
    compare_fit(g, o, nogap, 3.0, poly_rescale_n=1024, coeff_rtol=1e-2, resid_atol=5e-3)
    # order < context order: num_coeff selects the bases (baseline.cc:1301-1302)
    g, o, _ = run_both(lib, oracle, 2, data, mask, 3.0, 3, ctx_order=5)
    compare_fit(g, o, data, 3.0, poly_rescale_n=1024)


def test_golden_vectors(lib):
    z = np.load(os.path.join(GOLD, "poly5.npz"))
    out = {k[4:]: z[k] for k in z.files if k.startswith("out_")}
    st, ctx = lib.create_polynomial_context(int(z["order"]), z["data"].shape[1])
    g = lib.lsqfit_polynomial_batch(ctx, int(z["order"]), z["data"], z["mask"], float(z["clip"]),
                                    int(z["nfit"]), want_best_fit=True)
    lib.destroy_context(ctx)
    assert compare_fit(g, out, z["data"], float(z["clip"]), poly_rescale_n=z["data"].shape[1]) == []
    z = np.load(os.path.join(GOLD, "cheb4.npz"))
    out = {k[4:]: z[k] for k in z.files if k.startswith("out_")}
    st, ctx = lib.create_polynomial_context(int(z["order"]), z["data"].shape[1],
                                            lsqfit_type=LSQFIT_TYPE_CHEBYSHEV)
    g = lib.lsqfit_polynomial_batch(ctx, int(z["order"]), z["data"], z["mask"], float(z["clip"]),
                                    int(z["nfit"]), want_best_fit=True)
    lib.destroy_context(ctx)
    assert compare_fit(g, out, z["data"], float(z["clip"])) == []


# ---- the reference's API-contract cases through the single-row entry point ------------------
class Row:
    def __init__(self, lib, n=15, order=3, lsqfit_type=0):
        self.lib, self.n, self.order = lib, n, order
        st, self.ctx = lib.create_polynomial_context(order, n, lsqfit_type=lsqfit_type)
        assert st == STATUS_OK
        self.data = aligned_copy(np.full(n, 10.0, dtype=np.float32))
        self.mask = aligned_copy(np.ones(n, dtype=np.bool_))
        self.coeff = aligned_empty(order + 1, np.float64)
        self.best_fit = aligned_empty(n, np.float32)
        self.residual = aligned_empty(n, np.float32)
        self.final_mask = aligned_empty(n, np.bool_)
        self.reset()

    def reset(self):
        self.coeff[:] = -7.0
        self.best_fit[:] = -7.0
        self.residual[:] = -7.0
        self.final_mask[:] = False
        self.rms = C.c_float(-7.0)
        self.lsq = C.c_int(-7)

    def call(self, order=None, nfit=2, clip=3.0, num_coeff=None, coeff="d", best_fit="d", residual="d",
             final_mask="d"):
        order = self.order if order is None else order
        pick = lambda v, d: d.ctypes.data if isinstance(v, str) else (None if v is None else v.ctypes.data)  # noqa
        return self.lib.lib.sakura_LSQFitPolynomialFloat(
            self.ctx, order, self.n, self.data.ctypes.data, self.mask.ctypes.data, clip, nfit,
            order + 1 if num_coeff is None else num_coeff, pick(coeff, self.coeff),
            pick(best_fit, self.best_fit), pick(residual, self.residual), pick(final_mask, self.final_mask),
            C.byref(self.rms), C.byref(self.lsq))


def near(a, b):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    return np.all(np.abs(a - b) <= np.maximum(np.abs(a), np.abs(b)) * 1e-5 + 1e-5)  # lsq.cc:2096-2101


def test_default_case_values(lib):
    for t in (0, 1):  # polynomial, "contextVP chebyshev"
        r = Row(lib, lsqfit_type=t)
        assert r.call() == STATUS_OK and r.lsq.value == LSQFIT_OK
        assert near(r.coeff, [10, 0, 0, 0]) and near(r.best_fit, 10.0) and near(r.residual, 0.0)
        assert r.final_mask.all() and abs(r.rms.value) < 1e-5


def test_null_outputs_and_order_lb(lib):
    r = Row(lib)
    for kw in (dict(coeff=None), dict(best_fit=None), dict(residual=None), dict(coeff=None, best_fit=None),
               dict(coeff=None, residual=None), dict(best_fit=None, residual=None),
               dict(coeff=None, best_fit=None, residual=None)):
        r.reset()
        assert r.call(**kw) == STATUS_OK
        if "coeff" in kw:
            assert np.all(r.coeff == -7.0)
        else:
            assert near(r.coeff, [10, 0, 0, 0])
        if "residual" in kw:
            assert np.all(r.residual == -7.0)
    r.reset()
    assert r.call(order=0) == STATUS_OK and near(r.coeff[:1], [10.0])  # orderLB
    # num_coeff is not checked against order when coeff is NULL (baseline.cc:1709-1713)
    assert r.call(coeff=None, num_coeff=2) == STATUS_OK


def test_in_place_aliasing(lib):
    # best_fit=data / residual=data / final_mask=mask (lsq.cc:518-529, 1988-1997)
    r = Row(lib)
    r.data[:] = np.linspace(1.0, 2.0, r.n, dtype=np.float32)
    keep = r.data.copy()
    assert r.call(best_fit=r.data, residual=None) == STATUS_OK
    assert near(r.data, keep)  # a cubic reproduces a straight line
    r.data[:] = keep
    assert r.call(residual=r.data, best_fit=None) == STATUS_OK
    assert near(r.data, 0.0)
    r.data[:] = keep
    r.mask[3] = False
    assert r.call(final_mask=r.mask) == STATUS_OK
    assert not r.mask[3] and r.mask.sum() == r.n - 1


def test_effective_data_and_failure_semantics(lib):
    r = Row(lib)
    r.mask[:] = False
    r.mask[:4] = True  # effdataLB
    assert r.call() == STATUS_OK and r.lsq.value == LSQFIT_OK
    r.reset()
    r.mask[3] = False  # effdataTL -> kNG before the loop: nothing written, not even final_mask
    assert r.call() == STATUS_NG and r.lsq.value == LSQFIT_NOT_ENOUGH_DATA
    assert np.all(r.coeff == -7.0) and np.all(r.residual == -7.0) and r.rms.value == -7.0
    assert not r.final_mask.any()
    r.reset()
    r.mask[:] = False  # noeffdata
    assert r.call() == STATUS_NG and r.lsq.value == LSQFIT_NOT_ENOUGH_DATA


def test_clip_threshold_too_small_fails_after_clipping(lib, oracle):
    # lsq.cc:486-488: sigma=0.1, 5 fits, data[0]=1000 -> clipping drives the row below K -> kNG;
    # final_mask has already been clipped in place, everything else is untouched
    r = Row(lib)
    r.data[0] = 1000.0
    r.final_mask[:] = True
    assert r.call(clip=0.1, nfit=5) == STATUS_NG and r.lsq.value == LSQFIT_NOT_ENOUGH_DATA
    assert np.all(r.coeff == -7.0) and np.all(r.residual == -7.0) and r.rms.value == -7.0
    o = oracle.lsqfit_polynomial_batch(3, r.data, r.mask, 0.1, 5)
    assert np.array_equal(r.final_mask, o["final_mask"][0])


def test_num_fitting_max_bounds(lib, oracle):
    r = Row(lib)
    assert r.call(nfit=0) == STATUS_OK and r.lsq.value == LSQFIT_OK  # "no fitting": nothing written
    assert np.all(r.coeff == -7.0) and r.rms.value == -7.0 and not r.final_mask.any()
    r.reset()
    assert r.call(nfit=65535) == STATUS_OK  # num_fitting_maxUB
    assert near(r.coeff, [10, 0, 0, 0])
    data, mask = synth.make_spectra(64, 1024, seed=77)
    g, o, _ = run_both(lib, oracle, 3, data, mask, 2.0, 100)  # performance_num_fitting_max shape
    compare_fit(g, o, data, 2.0, poly_rescale_n=1024)


def test_python_binding_example(lib):
    y = np.array([0.5 * i for i in range(8)], dtype=np.float32)
    m = np.ones(8, dtype=np.bool_)
    y[4] += 3.0
    m[4] = False
    st, ctx = lib.create_polynomial_context(1, 8)
    g = lib.lsqfit_polynomial_batch(ctx, 1, y, m, 5.0, 1, want_best_fit=True)
    lib.destroy_context(ctx)
    assert g["status"][0] == STATUS_OK
    assert np.allclose(g["residual"][0], [0, 0, 0, 0, 3, 0, 0, 0], atol=1e-6)


def test_device_pointer_api_matches_host_api(lib, oracle):
    import torch
    data, mask = synth.make_spectra(512, 4096, seed=31)
    R, n, K = data.shape[0], data.shape[1], 6
    st, ctx = lib.create_polynomial_context(5, n)
    h = lib.lsqfit_polynomial_batch(ctx, 5, data, mask, 3.0, 5)
    d_data = torch.from_numpy(data).cuda()
    d_mask = torch.from_numpy(mask).cuda()
    d_coeff = torch.full((R, K), float("nan"), dtype=torch.float64, device="cuda")
    d_res = torch.full((R, n), float("nan"), dtype=torch.float32, device="cuda")
    d_rms = torch.full((R,), float("nan"), dtype=torch.float32, device="cuda")
    d_st = torch.empty(R, dtype=torch.int32, device="cuda")
    d_ls = torch.empty(R, dtype=torch.int32, device="cuda")
    stream = torch.cuda.Stream()
    with torch.cuda.stream(stream):
        # in place: final_mask == mask is allowed on the device entry point
        rc = lib.lib.sakura_LSQFitPolynomialFloatBatchDevice(
            ctx, 5, R, n, d_data.data_ptr(), n, d_mask.data_ptr(), n, 3.0, 5, K, d_coeff.data_ptr(), None,
            d_res.data_ptr(), d_mask.data_ptr(), d_rms.data_ptr(), d_st.data_ptr(), d_ls.data_ptr(),
            stream.cuda_stream)
    assert rc == STATUS_OK
    stream.synchronize()
    assert np.array_equal(d_mask.cpu().numpy(), h["final_mask"])
    assert np.array_equal(d_res.cpu().numpy(), h["residual"], equal_nan=True)
    assert np.array_equal(d_coeff.cpu().numpy(), h["coeff"], equal_nan=True)
    assert np.array_equal(d_rms.cpu().numpy(), h["rms"], equal_nan=True)
    assert np.array_equal(d_st.cpu().numpy(), h["status"])
    # residual in place of data is accepted on the row-resident kernels (tests/test_gpu_device_api.py::
    # test_residual_in_place_of_data compares it with the out-of-place call); a partial overlap is not
    rc = lib.lib.sakura_LSQFitPolynomialFloatBatchDevice(
        ctx, 5, R - 1, n, d_data.data_ptr(), n, d_mask.data_ptr(), n, 3.0, 5, K, None, None,
        d_data.data_ptr() + 4 * n, d_mask.data_ptr(), d_rms.data_ptr(), d_st.data_ptr(), d_ls.data_ptr(), None)
    assert rc == 2
    lib.destroy_context(ctx)


def test_pipelined_host_path_equals_single_chunk(lib):
    """Batches larger than one pipeline chunk (96 MiB of data) go through the 3-slot
    copy/compute pipeline; results must not depend on the chunking, including failed rows."""
    R, n = 9000, 4096
    data, mask = synth.make_spectra(R, n, seed=41, lines=False)
    mask[[5, 6001, 8999]] = False  # rows that fail in different chunks
    st, ctx = lib.create_polynomial_context(5, n)
    big = lib.lsqfit_polynomial_batch(ctx, 5, data, mask, 3.0, 3)
    assert big["call_status"] == STATUS_OK
    for lo, hi in ((0, 16), (5990, 6010), (8990, 9000)):
        part = lib.lsqfit_polynomial_batch(ctx, 5, data[lo:hi], mask[lo:hi], 3.0, 3)
        for k in ("coeff", "residual", "final_mask", "rms", "status", "lsqfit_status"):
            assert np.array_equal(big[k][lo:hi], part[k], equal_nan=True), k
    assert big["status"][5] == STATUS_NG and np.isnan(big["rms"][5]) and not big["final_mask"][5].any()
    lib.destroy_context(ctx)


def test_results_are_deterministic(lib):
    data, mask = synth.make_spectra(300, 4096, seed=51)
    st, ctx = lib.create_polynomial_context(5, 4096)
    a = lib.lsqfit_polynomial_batch(ctx, 5, data, mask, 3.0, 5)
    b = lib.lsqfit_polynomial_batch(ctx, 5, data[::-1].copy(), mask[::-1].copy(), 3.0, 5)
    lib.destroy_context(ctx)
    for k in ("coeff", "residual", "final_mask", "rms"):
        assert np.array_equal(a[k], b[k][::-1], equal_nan=True), k


def test_chebyshev_high_order_large_shape(lib, oracle):
    # order 30 at 8192 channels with the usual flagged edges / gaps: the tensor-core path
    data, mask = synth.make_spectra(40, 8192, seed=77, ripple=True)
    g, o, kernel = run_both(lib, oracle, 30, data, mask, 3.0, 3, lsqfit_type=LSQFIT_TYPE_CHEBYSHEV,
                            want_best_fit=True)
    assert kernel == "chebyshev_dmma"
    compare_fit(g, o, data, 3.0, resid_atol=1e-4, coeff_rtol=1e-3)
