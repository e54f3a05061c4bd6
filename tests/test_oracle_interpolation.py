"""Interpolation: the oracles against the reference's own known-answer tests
(libsakura/test/interpolation_y.cc:109-530, re-expressed without gtest) and the C restatement
against the reference's object code, bit for bit, on random masked inputs in both layouts and
all four sort-order combinations."""
import numpy as np
import pytest

NEAREST, LINEAR, POLYNOMIAL, SPLINE = 0, 1, 2, 3


def oracles(request):
    kinds = [request.getfixturevalue("oracle_port")]
    try:
        kinds.append(request.getfixturevalue("oracle_ref"))
    except pytest.skip.Exception:
        pass
    return kinds


def grid13():
    x = np.empty(13)
    x[0], x[-1] = -1.0, 10.0
    x[1:-1] = np.linspace(0.0, 2.0, 11)  # EquallySpacedGrid, test :292
    return x


def known_answer_cases():
    """(name, method, x_base, y_base [nb, na], mask, x_interp, expected [ni, na], expected mask)"""
    cases = []
    xb2 = [0.0, 1.0]
    yb2 = np.array([[1.0, -0.5], [-1.0, 0.2]], dtype=np.float32)
    xi5 = [-1.0, 0.1, 0.5, 0.9, 1.2]
    near = np.array([[1.0, -0.5], [1.0, -0.5], [1.0, -0.5], [-1.0, 0.2], [-1.0, 0.2]], dtype=np.float32)
    cases.append(("Nearest", NEAREST, xb2, yb2, np.ones((2, 2), bool), xi5, near, np.ones((5, 2), bool)))
    cases.append(("AllDataMasked", NEAREST, xb2, yb2, np.zeros((2, 2), bool), xi5, None, np.zeros((5, 2), bool)))
    cases.append(("PolynomialOrder0", POLYNOMIAL, xb2, yb2, np.ones((2, 2), bool), xi5, near, np.ones((5, 2), bool)))
    xb3 = [0.0, 1.0, 2.0]
    yb3 = np.array([[1.0, 0.0], [-1.0, 0.5], [0.0, -0.2]], dtype=np.float32)
    xi = grid13()
    exp = np.empty((13, 2), dtype=np.float32)
    exp[0], exp[-1] = yb3[0], yb3[-1]
    for i in range(1, 12):
        l = 0 if xi[i] < 1.0 else 1
        frac = (xi[i] - xb3[l]) / (xb3[l + 1] - xb3[l])
        exp[i] = yb3[l] + frac * (yb3[l + 1] - yb3[l])
    cases.append(("Linear", LINEAR, xb3, yb3, np.ones((3, 2), bool), xi, exp, np.ones((13, 2), bool)))
    ym = np.array([[1.0, 0.0], [-1.0, -1.0e10], [0.0, -0.2]], dtype=np.float32)
    mm = np.ones((3, 2), bool)
    mm[1, 1] = False
    em = np.array([1.0, 0.0, 1.0, 0.0, 0.6, -0.02, 0.2, -0.04, -0.2, -0.06, -0.6, -0.08, -1.0, -0.1, -0.8, -0.12,
                   -0.6, -0.14, -0.4, -0.16, -0.2, -0.18, 0.0, -0.2, 0.0, -0.2], dtype=np.float32).reshape(13, 2)
    cases.append(("LinearMask", LINEAR, xb3, ym, mm, xi, em, np.ones((13, 2), bool)))
    ye = np.array([[1.0e10, 0.0], [-1.0, 0.5], [0.0, -1.0e10]], dtype=np.float32)
    me = np.ones((3, 2), bool)
    me[0, 0] = me[2, 1] = False
    ee = np.array([-1.0, 0.0, -1.0, 0.0, -1.0, 0.1, -1.0, 0.2, -1.0, 0.3, -1.0, 0.4, -1.0, 0.5, -0.8, 0.5, -0.6, 0.5,
                   -0.4, 0.5, -0.2, 0.5, 0.0, 0.5, 0.0, 0.5], dtype=np.float32).reshape(13, 2)
    cases.append(("LinearMaskEdge", LINEAR, xb3, ye, me, xi, ee, np.ones((13, 2), bool)))
    ma = np.ones((3, 2), bool)
    ma[:, 1] = False
    ema = np.ones((13, 2), bool)
    ema[:, 1] = False
    cases.append(("LinearMaskArray", LINEAR, xb3, ym, ma, xi, em, ema))
    return cases


@pytest.mark.parametrize("case", known_answer_cases(), ids=lambda c: c[0])
def test_reference_known_answers(request, case):
    name, method, xb, yb, mb, xi, expected, expected_mask = case
    for ora in oracles(request):
        out, om, st = ora.interpolate("y", method, xb, yb, mb, xi)
        assert st == 0, (ora.kind, name)
        assert np.array_equal(om, expected_mask), (ora.kind, name)
        if expected is not None:
            ok = om  # arrays without a valid base point keep their previous content
            assert np.allclose(out[ok], expected[ok], rtol=1e-5, atol=1e-6), (ora.kind, name)
        if not om.any():
            assert np.isnan(out).all()  # untouched
        # descending base and / or interpolated positions give the same answer per position
        out2, om2, st2 = ora.interpolate("y", method, xb[::-1], yb[::-1], mb[::-1], np.asarray(xi)[::-1])
        assert st2 == 0 and np.array_equal(om2, om[::-1])
        assert np.array_equal(np.nan_to_num(out2), np.nan_to_num(out[::-1]))
        # the X-axis layout is the transpose
        out3, om3, st3 = ora.interpolate("x", method, xb, yb.T.copy(), mb.T.copy(), xi)
        assert st3 == 0 and np.array_equal(om3, om.T)
        assert np.array_equal(np.nan_to_num(out3), np.nan_to_num(out.T))


def test_status_rules(request):
    for ora in oracles(request):
        yb = np.ones((1, 5), dtype=np.float32)
        assert ora.interpolate("y", NEAREST, [], np.ones((0, 5), np.float32), np.ones((0, 5), bool), [0.5])[2] == 2
        out, om, st = ora.interpolate("y", NEAREST, [0.0], yb, np.ones((1, 5), bool), [])
        assert st == 0  # nothing to do (interpolation.cc:1311-1316)
        # a single base point: every position takes its value (test :109-123)
        out, om, st = ora.interpolate("y", LINEAR, [0.3], yb * 2, np.ones((1, 5), bool), [-1.0, 0.3, 7.0])
        assert st == 0 and om.all() and np.all(out == 2.0)


def random_problem(rng, nb, na, ni, base_desc, interp_desc, valid_fraction):
    xb = np.sort(rng.choice(np.arange(10 * nb), nb, replace=False).astype(np.float64)) * 0.37
    lo, hi = xb[0] - 2.0, xb[-1] + 2.0
    xi = np.unique(np.concatenate([rng.uniform(lo, hi, ni), rng.choice(xb, min(nb, 3))]))
    xi = np.concatenate([xi, 0.5 * (xb[:-1] + xb[1:])[:3]]) if nb > 1 else xi
    xi = np.unique(xi)
    yb = rng.standard_normal((nb, na)).astype(np.float32) * 100
    mb = rng.random((nb, na)) < valid_fraction
    mb[:, 0] = False                      # an array without any valid point
    if na > 2:
        mb[:, 1] = False
        mb[nb // 2, 1] = True             # an array with exactly one
    if base_desc:
        xb, yb, mb = xb[::-1].copy(), yb[::-1].copy(), mb[::-1].copy()
    if interp_desc:
        xi = xi[::-1].copy()
    return xb, yb, mb, xi


@pytest.mark.parametrize("base_desc", [False, True])
@pytest.mark.parametrize("interp_desc", [False, True])
def test_port_equals_reference(oracle_port, oracle_ref, base_desc, interp_desc):
    rng = np.random.default_rng(17 + 2 * base_desc + interp_desc)
    for nb, na, ni, frac in ((2, 3, 9, 1.0), (5, 7, 40, 0.7), (31, 16, 200, 0.5), (1, 4, 6, 1.0), (64, 33, 129, 0.9)):
        xb, yb, mb, xi = random_problem(rng, nb, na, ni, base_desc, interp_desc, frac)
        for method in (NEAREST, LINEAR):
            for axis in ("y", "x"):
                bd, bm = (yb, mb) if axis == "y" else (yb.T.copy(), mb.T.copy())
                a, am, sa = oracle_port.interpolate(axis, method, xb, bd, bm, xi)
                b, bm2, sb = oracle_ref.interpolate(axis, method, xb, bd, bm, xi)
                assert sa == sb == 0
                assert np.array_equal(am, bm2), (nb, na, method, axis)
                assert np.array_equal(a.view(np.uint32), b.view(np.uint32)), (nb, na, method, axis)
