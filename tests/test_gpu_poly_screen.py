"""The screening kernel (csrc/lsqfit_poly_screen.cu) and the warp-per-spectrum kernel
(csrc/lsqfit_poly_warp.cu) against the per-channel kernel
(csrc/lsqfit_poly.cu) and the oracle: a row the screening kernel finishes must carry exactly the
clip decisions of the per-channel evaluation; rows it cannot decide are handed over."""
import os

import numpy as np
import pytest

from sakura_b200 import synth
from sakura_b200.capi import STATUS_OK
from parity import compare_fit, forced_kernel

pytestmark = pytest.mark.gpu


KERNEL_NAME = {"screen": "poly_screen", "warp": "poly_warp", "channel": "poly_fast"}


def fit(lib, order, data, mask, clip, nfit, which, **want):
    """`which`: "screen" / "warp" / "channel" (True / False = screen / channel)."""
    st, ctx = lib.create_polynomial_context(order, data.shape[-1])
    assert st == STATUS_OK
    with forced_kernel(which):
        g = lib.lsqfit_polynomial_batch(ctx, order, data, mask, clip, nfit, **want)
        kernel = lib.last_kernel_name()
        handed = lib.last_handover_count(ctx)
    lib.destroy_context(ctx)
    return g, kernel, handed


FAST = ["screen", "warp"]


def assert_same(a, b, data, loose=False):
    assert np.array_equal(a["status"], b["status"])
    assert np.array_equal(a["lsqfit_status"], b["lsqfit_status"])
    assert np.array_equal(a["final_mask"], b["final_mask"]), int((a["final_mask"] != b["final_mask"]).sum())
    good = a["status"] == 0
    ra, rb = a["rms"][good].astype(np.float64), b["rms"][good].astype(np.float64)
    # same clip decisions -> same fit; the last fit's model is evaluated from a different
    # expansion point in the two kernels, so a float residual may differ by one ulp of the model
    assert np.all(np.abs(ra - rb) <= 1e-6 * np.abs(rb) + 1e-30)
    # Rows fitted through a handful of valid channels are interpolations: between those channels the
    # model is not determined by the data to an ulp (the two kernels' power sums differ in the last
    # bit), so such rows are compared on the channels that constrain the fit, to 1e-5.
    dense = a["final_mask"][good].sum(axis=1) >= 64
    for key in ("residual", "best_fit"):
        if a.get(key) is None:
            continue
        x, y = a[key][good].astype(np.float64), b[key][good].astype(np.float64)
        model = np.abs(np.nan_to_num(data[good].astype(np.float64))) + np.abs(np.nan_to_num(y))
        both_nan = np.isnan(x) & np.isnan(y)
        sparse_ok = (np.abs(x - y) <= 1e-5 * model + 1e-30) | both_nan | ~a["final_mask"][good]
        assert np.all(sparse_ok[~dense]), key
        x, y, model, both_nan = x[dense], y[dense], model[dense], both_nan[dense]
        # orders 6, 7: cond(G) ~ 1e9..1e10 turns the O(eps) differences of the two kernels' moment
        # sums into 1e-6-level model differences (tests/test_gpu_polynomial.py widens likewise)
        assert np.all((np.abs(x - y) <= (2e-4 if loose else 2.4e-7) * model + 1e-30) | both_nan), key
    ca, cb = a["coeff"][good][dense], b["coeff"][good][dense]
    scale = np.abs(cb).max(axis=1, keepdims=True) + 1e-300
    assert np.all(np.abs(ca - cb) <= (1e-3 if loose else 1e-7) * np.abs(cb) + (1e-4 if loose else 1e-12) * scale)


@pytest.mark.parametrize("which,n,order,nfit",
                         [("screen", n, o, f) for n, o, f in [(4096, 5, 5), (4096, 5, 3), (4096, 5, 2), (4096, 5, 1),
                                                              (4096, 0, 4), (4096, 1, 4), (4096, 3, 6), (4096, 7, 3),
                                                              (8192, 5, 5), (8192, 2, 3)]]
                         + [("warp", n, o, f) for n, o, f in [(4096, 5, 5), (4096, 5, 3), (4096, 5, 2), (4096, 5, 1),
                                                              (4096, 0, 4), (4096, 1, 4), (4096, 2, 7), (4096, 3, 6),
                                                              (4096, 4, 5), (4096, 6, 4), (4096, 7, 3),
                                                              (2048, 5, 5), (2048, 2, 3), (2048, 7, 4), (2048, 0, 2)]])
def test_screen_equals_per_channel_kernel(lib, oracle, which, n, order, nfit):
    rows = 768
    data, mask = synth.make_spectra(rows, n, seed=100 + order + nfit)
    synth.add_edge_case_rows(data, mask, order + 1)
    s, kernel, handed = fit(lib, order, data, mask, 3.0, nfit, which, want_best_fit=True)
    assert kernel == KERNEL_NAME[which]
    p, kernel, _ = fit(lib, order, data, mask, 3.0, nfit, "channel", want_best_fit=True)
    assert kernel == "poly_fast"
    assert_same(s, p, data, loose=order >= 6)
    # the edge-case rows (constant data, exactly K valid channels, ...) are handed over by design;
    # ordinary rows almost never are
    # (orders 6 and 7: the monomial coefficients are large, the FP64 slack terms of the bounds grow
    # with their squares and more rows are handed over)
    assert 0 <= handed <= 16 + rows // (5 if order >= 6 else 50), handed
    if order <= 5:
        o = oracle.lsqfit_polynomial_batch(order, data, mask, 3.0, nfit, want_best_fit=True)
        reported = compare_fit(s, o, data, 3.0, poly_rescale_n=n)
        assert reported == [], reported


@pytest.mark.parametrize("which", FAST)
@pytest.mark.parametrize("scale,offset", [(1.0, 1000.0), (1e-6, 0.0), (1e6, 3e7), (1e-20, 0.0), (1.0, 1e6)])
def test_screen_scales_and_offsets(lib, oracle, which, scale, offset):
    """Large offsets widen the float-rounding band (more hand-overs, same results); tiny scales
    push the thresholds below the underflow guard (every row handed over)."""
    data, mask = synth.make_spectra(256, 4096, seed=77)
    data = (data * np.float32(scale) + np.float32(offset)).astype(np.float32)
    s, kernel, handed = fit(lib, 5, data, mask, 3.0, 5, which)
    assert kernel == KERNEL_NAME[which]
    p, _, _ = fit(lib, 5, data, mask, 3.0, 5, "channel")
    assert_same(s, p, data)
    # ... and both against the reference itself (status, final_mask bit-exact; the value gates scale with
    # the data: 1e-5 relative, one float ulp of the model)
    o = oracle.lsqfit_polynomial_batch(5, data, mask, 3.0, 5)
    for got in (s, p):
        assert np.array_equal(got["status"], o["status"])
        assert np.array_equal(got["lsqfit_status"], o["lsqfit_status"])
        assert np.array_equal(got["final_mask"], o["final_mask"]), int((got["final_mask"] != o["final_mask"]).sum())
        good = o["status"] == 0
        # (1e-6 relative; with an offset of 1e6 one float ulp of the data is 0.06 sigma and the float model
        # of a few channels per row rounds the other way: the ulp-of-data floor of tests/parity.py)
        floor = 2.4e-7 * np.abs(data[good]).max(axis=1)
        assert np.all(np.abs(got["rms"][good].astype(np.float64) - o["rms"][good]) <= 1e-6 * o["rms"][good] + floor)
        model = np.abs(data[good].astype(np.float64) - o["residual"][good])
        tol = 1e-5 * np.abs(o["residual"][good].astype(np.float64)) + 2.4e-7 * model
        assert np.all(np.abs(got["residual"][good].astype(np.float64) - o["residual"][good]) <= tol)
    if scale == 1e-20:
        assert handed == 256


@pytest.mark.parametrize("which", FAST)
@pytest.mark.parametrize("clip", [0.5, 1.5, 2.5, 6.0])
def test_screen_clip_levels(lib, oracle, which, clip):
    """Low thresholds clip a large part of the row per pass (long clipped lists, several rounds
    per warp); many iterations exercise the stale passes."""
    data, mask = synth.make_spectra(128, 4096, seed=int(clip * 10))
    s, _, handed = fit(lib, 4, data, mask, clip, 12, which)
    p, _, _ = fit(lib, 4, data, mask, clip, 12, "channel")
    assert_same(s, p, data)
    # ... and against the reference. At clip 0.5 a pass removes 60 % of the channels that are left: after
    # six passes a row is down to a dozen channels, the quartic interpolates them and the "residual" is the
    # rounding noise of the model, so what 0.5 x rms(noise) clips is no longer a property of the data. The
    # reference comparison therefore stops after four fits there (~230 channels left); the kernels are
    # compared with each other over all twelve above.
    nfit_ref = 4 if clip < 1.0 else 12
    if nfit_ref != 12:
        s, _, _ = fit(lib, 4, data, mask, clip, nfit_ref, which)
        p, _, _ = fit(lib, 4, data, mask, clip, nfit_ref, "channel")
    o = oracle.lsqfit_polynomial_batch(4, data, mask, clip, nfit_ref)
    compare_fit(s, o, data, clip, poly_rescale_n=4096)
    compare_fit(p, o, data, clip, poly_rescale_n=4096)


@pytest.mark.parametrize("which", FAST)
def test_screen_dense_flags_and_runs(lib, oracle, which):
    """Rows with most channels flagged (direct power sums instead of the complement), long
    flagged runs, rows that run out of channels while clipping."""
    rng = np.random.default_rng(5)
    data, mask = synth.make_spectra(192, 4096, seed=9)
    mask[:64] &= rng.random((64, 4096)) < 0.3
    for r in range(64, 128):
        a = int(rng.integers(0, 3000))
        mask[r, a:a + int(rng.integers(300, 1000))] = False
    mask[128:160] &= rng.random((32, 4096)) < 0.004   # ~16 valid channels: clipping exhausts them
    s, _, handed = fit(lib, 5, data, mask, 3.0, 6, which, want_best_fit=True)
    p, _, _ = fit(lib, 5, data, mask, 3.0, 6, "channel", want_best_fit=True)
    assert_same(s, p, data)
    assert (s["status"] != 0).sum() == (p["status"] != 0).sum()
    o = oracle.lsqfit_polynomial_batch(5, data, mask, 3.0, 6, want_best_fit=True)
    compare_fit(s, o, data, 3.0, poly_rescale_n=4096)
    compare_fit(p, o, data, 3.0, poly_rescale_n=4096)


@pytest.mark.parametrize("which,rows,n,nfit,clip", [("screen", 1 << 17, 4096, 5, 3.0), ("screen", 1 << 16, 4096, 9, 2.5),
                                                    ("screen", 1 << 15, 8192, 6, 3.0), ("warp", 1 << 17, 4096, 5, 3.0),
                                                    ("warp", 1 << 16, 4096, 9, 2.5), ("warp", 1 << 17, 2048, 6, 3.0),
                                                    ("warp", 1 << 16, 4096, 2, 3.5)])
def test_screen_large_device_batches(lib, which, rows, n, nfit, clip):
    """Device-resident batches: identical clip decisions with and without screening on every row,
    bit-identical results from run to run (fixed reduction orders, no atomics in the data path)."""
    import torch
    import bench
    from sakura_b200.device_api import DeviceFits
    dev = DeviceFits(lib)
    data, mask = bench.generate_on_device(torch, rows, n, seed=rows + nfit, device=torch.device("cuda", 0))
    st, ctx = lib.create_polynomial_context(5, n)
    assert st == STATUS_OK
    with forced_kernel(which):
        a = dev.polynomial(ctx, 5, data, mask, clip, nfit)
        assert lib.last_kernel_name() == KERNEL_NAME[which]
        handed = lib.last_handover_count(ctx)
        b = dev.polynomial(ctx, 5, data, mask, clip, nfit)
    with forced_kernel("channel"):
        c = dev.polynomial(ctx, 5, data, mask, clip, nfit)
        assert lib.last_kernel_name() == "poly_fast"
    lib.destroy_context(ctx)
    assert 0 <= handed <= rows // 200, handed
    for key in ("final_mask", "status", "lsqfit_status", "rms", "coeff", "residual"):
        assert torch.equal(a[key], b[key]), key                   # run-to-run
    assert torch.equal(a["final_mask"], c["final_mask"])
    assert torch.equal(a["status"], c["status"])
    rel = ((a["rms"] - c["rms"]).abs() / c["rms"].abs()).max().item()
    assert rel <= 1e-6, rel


@pytest.mark.parametrize("which", FAST)
def test_screen_non_finite_and_degenerate_rows(lib, oracle, which):
    """NaN / Inf in a valid channel, constant rows, rows of zeros, huge and denormal values: the
    screening kernel must hand such rows over (or decide them identically); whatever the
    per-channel kernel produces for them is what comes out."""
    data, mask = synth.make_spectra(96, 4096, seed=21)
    data[3, 100] = np.nan
    mask[3, 100] = True
    data[5, 2000] = np.inf
    mask[5, 2000] = True
    data[7, 17] = -np.inf
    mask[7, 17] = True
    data[9] = 2.5
    data[11] = 0.0
    data[13] *= np.float32(1e30)
    data[15] *= np.float32(1e-42)
    data[17, ::2] = 0.0                       # half of the row exactly on the model
    data[19] = np.arange(4096, dtype=np.float32)  # exactly polynomial: residual is rounding noise
    s, kernel, handed = fit(lib, 5, data, mask, 3.0, 5, which, want_best_fit=True)
    assert kernel == KERNEL_NAME[which]
    p, _, _ = fit(lib, 5, data, mask, 3.0, 5, "channel", want_best_fit=True)
    assert np.array_equal(s["status"], p["status"])
    assert np.array_equal(s["lsqfit_status"], p["lsqfit_status"])
    assert np.array_equal(s["final_mask"], p["final_mask"])
    special = [3, 5, 7, 9, 11, 13, 15, 19]
    for key in ("rms", "residual", "coeff", "best_fit"):
        a, b = s[key][special], p[key][special]
        assert np.array_equal(np.isnan(a), np.isnan(b)), key
        with np.errstate(over="ignore"):
            assert np.array_equal(np.nan_to_num(a).view(np.uint8), np.nan_to_num(b).view(np.uint8)), key
    assert handed >= len(special) - 1
    # the reference itself on the same rows: statuses and clip decisions bit-exact, NaN where it has NaN
    # (a NaN / Inf in a valid channel poisons the normal equations: Eigen's full-pivot LU then decides what
    # comes out, the per-channel kernel restates that LU)
    o = oracle.lsqfit_polynomial_batch(5, data, mask, 3.0, 5, want_best_fit=True)
    for got in (s, p):
        assert np.array_equal(got["status"], o["status"]), (got["status"][special], o["status"][special])
        assert np.array_equal(got["lsqfit_status"], o["lsqfit_status"])
        assert np.array_equal(got["final_mask"], o["final_mask"]), np.argwhere(got["final_mask"] != o["final_mask"])[:8]
        for key in ("rms", "residual", "best_fit"):
            assert np.array_equal(np.isnan(got[key]), np.isnan(o[key])), key
    ordinary = np.ones(data.shape[0], dtype=bool)
    ordinary[special + [17]] = False
    sub = lambda d: {k: (v[ordinary] if isinstance(v, np.ndarray) else v) for k, v in d.items()}
    so = sub(o)
    so.pop("refit")
    gs = sub(s)
    gs["call_status"] = 0
    compare_fit(gs, so, data[ordinary], 3.0, poly_rescale_n=4096)
