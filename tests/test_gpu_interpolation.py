"""GPU parity of sakura_Interpolate{X,Y}AxisFloat (csrc/interpolation.cu) with the
oracle: bit-exact values and masks, untouched outputs for arrays without a valid base point,
argument rules, device-array variant at the e2e shape (Tsys[time][channel] -> one row per spectrum)."""
import ctypes as C

import numpy as np
import pytest

from sakura_b200.capi import STATUS_OK, aligned_copy, aligned_empty
from test_oracle_interpolation import LINEAR, NEAREST, POLYNOMIAL, SPLINE, known_answer_cases, random_problem

pytestmark = pytest.mark.gpu
INVALID = 2


@pytest.mark.parametrize("case", known_answer_cases(), ids=lambda c: c[0])
def test_reference_known_answers(lib, case):
    name, method, xb, yb, mb, xi, expected, expected_mask = case
    out, om, st = lib.interpolate("y", method, xb, yb, mb, xi)
    assert st == STATUS_OK and lib.last_kernel_name() == "interpolate"
    assert np.array_equal(om, expected_mask)
    if expected is not None:
        assert np.allclose(out[om], expected[om], rtol=1e-5, atol=1e-6)
    if not om.any():
        assert np.isnan(out).all()


@pytest.mark.parametrize("base_desc", [False, True])
@pytest.mark.parametrize("interp_desc", [False, True])
def test_random_masked_problems_against_oracle(lib, oracle, base_desc, interp_desc):
    rng = np.random.default_rng(5 + 2 * base_desc + interp_desc)
    for nb, na, ni, frac in ((2, 3, 9, 1.0), (5, 7, 40, 0.7), (31, 16, 200, 0.5), (1, 4, 6, 1.0),
                             (64, 33, 129, 0.9), (8, 4096, 300, 0.97)):
        xb, yb, mb, xi = random_problem(rng, nb, na, ni, base_desc, interp_desc, frac)
        for method in (NEAREST, LINEAR):
            for axis in ("y", "x"):
                bd, bm = (yb, mb) if axis == "y" else (yb.T.copy(), mb.T.copy())
                g, gm, sg = lib.interpolate(axis, method, xb, bd, bm, xi)
                o, om, so = oracle.interpolate(axis, method, xb, bd, bm, xi)
                assert sg == so == 0
                assert np.array_equal(gm, om), (nb, na, method, axis)
                assert np.array_equal(g.view(np.uint32), o.view(np.uint32)), (nb, na, method, axis)


@pytest.mark.parametrize("nb", [12, 13, 14, 15, 16, 17])
def test_y_axis_table_kernel_boundary(lib, oracle, nb):
    """The per-array interval table of the Y-axis kernel is used up to 15 base points, where it needs
    more than 48 KiB of shared memory (opt-in attribute); 16 and 17 go through the bisection kernel."""
    rng = np.random.default_rng(100 + nb)
    xb, yb, mb, xi = random_problem(rng, nb, 300, 257, False, False, 0.9)
    for method in (NEAREST, LINEAR):
        g, gm, sg = lib.interpolate("y", method, xb, yb, mb, xi)
        o, om, so = oracle.interpolate("y", method, xb, yb, mb, xi)
        assert sg == so == 0, lib.last_error()
        assert np.array_equal(gm, om)
        assert np.array_equal(g.view(np.uint32), o.view(np.uint32))


def test_argument_rules(lib):
    L = lib.lib
    xb = aligned_copy(np.array([0.0, 1.0]))
    yb = aligned_copy(np.array([[1.0, -0.5], [-1.0, 0.2]], dtype=np.float32))
    mb = aligned_copy(np.ones((2, 2), bool))
    xi = aligned_copy(np.array([0.25, 0.5]))
    out = aligned_empty((2, 2), np.float32)
    om = aligned_empty((2, 2), np.bool_)
    p = lambda a, off=0: C.c_void_p(a.ctypes.data + off)
    f = L.sakura_InterpolateYAxisFloat
    assert f(LINEAR, 0, 2, p(xb), 2, p(yb), p(mb), 2, p(xi), p(out), p(om)) == STATUS_OK
    assert f(LINEAR, 0, 0, p(xb), 2, p(yb), p(mb), 2, p(xi), p(out), p(om)) == INVALID      # test :69
    assert f(LINEAR, 0, 2, p(xb), 2, p(yb), p(mb), 0, p(xi), p(out), p(om)) == STATUS_OK    # test :82
    assert f(LINEAR, 0, 2, p(xb), 0, p(yb), p(mb), 2, p(xi), p(out), p(om)) == STATUS_OK
    assert f(LINEAR, 0, 2, p(xb, 8), 2, p(yb), p(mb), 2, p(xi), p(out), p(om)) == INVALID   # test :95
    assert f(LINEAR, 0, 2, p(xb), 2, p(yb), p(mb), 2, p(xi), p(out, 4), p(om)) == INVALID
    assert f(LINEAR, 0, 2, p(xb), 2, None, p(mb), 2, p(xi), p(out), p(om)) == INVALID
    assert f(4, 0, 2, p(xb), 2, p(yb), p(mb), 2, p(xi), p(out), p(om)) == INVALID           # test :53
    # not implemented on the GPU path: rejected, never approximated
    assert f(SPLINE, 0, 2, p(xb), 2, p(yb), p(mb), 2, p(xi), p(out), p(om)) == INVALID
    assert f(POLYNOMIAL, 2, 2, p(xb), 2, p(yb), p(mb), 2, p(xi), p(out), p(om)) == INVALID
    assert "not implemented" in lib.last_error()
    assert f(POLYNOMIAL, 0, 2, p(xb), 2, p(yb), p(mb), 2, p(xi), p(out), p(om)) == STATUS_OK  # = nearest


def test_device_variant_tsys_rows(lib, oracle):
    """Tsys measured at a few times, interpolated to the time of every ON spectrum: the rows the
    calibration multiplies with (bench/e2eReduce/src/main.cc:86-104)."""
    import torch
    rng = np.random.default_rng(9)
    nb, nchan, nrow = 6, 4096, 512
    t_base = np.sort(rng.uniform(0, 100, nb))
    t_rows = np.sort(rng.uniform(-5, 105, nrow))
    tsys = (150 + 10 * rng.standard_normal((nb, nchan))).astype(np.float32)
    valid = rng.random((nb, nchan)) < 0.98
    d = lambda a: torch.from_numpy(np.ascontiguousarray(a)).cuda()
    d_tb, d_tr, d_ts, d_v = d(t_base), d(t_rows), d(tsys), d(valid)
    out = torch.full((nrow, nchan), float("nan"), dtype=torch.float32, device="cuda")
    om = torch.zeros((nrow, nchan), dtype=torch.bool, device="cuda")
    vp = C.c_void_p
    rc = lib.lib.sakura_InterpolateYAxisFloatDevice(LINEAR, 0, nb, vp(d_tb.data_ptr()), nchan, vp(d_ts.data_ptr()),
                                                    vp(d_v.data_ptr()), nrow, vp(d_tr.data_ptr()),
                                                    vp(out.data_ptr()), vp(om.data_ptr()), None)
    assert rc == STATUS_OK
    torch.cuda.synchronize()
    o, omask, so = oracle.interpolate("y", LINEAR, t_base, tsys, valid, t_rows)
    assert so == 0
    assert np.array_equal(om.cpu().numpy(), omask)
    assert np.array_equal(out.cpu().numpy().view(np.uint32), o.view(np.uint32))
