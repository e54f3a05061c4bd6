"""Oracle for the position-switch calibration pinned to the reference's own
known-answer tests (libsakura/test/normalization.cc) and, bit for bit, to the reference's
compiled code (oracle/_ref) where present."""
import numpy as np
import pytest

from sakura_b200.capi import aligned_copy, aligned_empty


def kinds():
    from oracle.refbind import Oracle
    out = []
    for kind in ("reference", "port"):
        try:
            out.append(Oracle(kind))
        except Exception:
            if kind == "port":
                raise
    return out


@pytest.mark.parametrize("ora", kinds(), ids=lambda o: o.kind)
def test_known_answers(ora):
    # BasicTest (normalization.cc:112-122): factors (0.5, 2), target (1, 1), reference (1, 0.5)
    res, st = ora.calibrate_batch([0.5, 2.0], [1.0, 1.0], [1.0, 0.5])
    assert st[0] == 0 and np.array_equal(res[0], np.float32([0.0, 2.0]))
    # InPlaceTest (:124-133, 380-383): result overwrites the input
    res, st = ora.calibrate_batch([0.5, 2.0], [1.0, 1.0], [1.0, 0.5], in_place=True)
    assert st[0] == 0 and np.array_equal(res[0], np.float32([0.0, 2.0]))
    # SingleScalingFactor (:135-144): constant 0.5 -> (0, 0.5)
    res, st = ora.calibrate_batch(0.5, [1.0, 1.0], [1.0, 0.5])
    assert st[0] == 0 and np.array_equal(res[0], np.float32([0.0, 0.5]))
    # ZeroDivision (:88-107): +inf and -inf, status kOK
    res, st = ora.calibrate_batch(1.0, [1.0, -1.0], [0.0, 0.0])
    assert st[0] == 0 and res[0, 0] == np.inf and res[0, 1] == -np.inf
    # IntrinsicsTest (:158-169): ten elements, factor 10, target 2, reference 1 -> 10
    res, st = ora.calibrate_batch(np.full(10, 10.0), np.full(10, 2.0), np.ones(10))
    assert st[0] == 0 and np.array_equal(res[0], np.full(10, 10.0, np.float32))


def test_port_bit_identical_to_reference():
    from oracle.refbind import Oracle
    try:
        ref = Oracle("reference")
    except Exception:
        pytest.skip("oracle/_ref not built")
    port = Oracle("port")
    g = np.random.Generator(np.random.Philox(77))
    for n in (1, 7, 8, 9, 64, 1000, 4096):
        rows = 5 if n % 8 == 0 else 1  # rows of a tight batch are aligned only when n % 8 == 0
        t = g.standard_normal((rows, n)).astype(np.float32) * 100
        r = (g.standard_normal((rows, n)).astype(np.float32) * 50 + 300).astype(np.float32)
        s = g.uniform(50, 400, (rows, n)).astype(np.float32)
        t[0, 0], r[rows // 2, 0], s[rows - 1, 0] = np.nan, 0.0, np.inf
        r[rows - 1, n // 2] = 1e-42  # subnormal reference
        for scaling in (s, 123.25):
            for in_place in (False, True):
                a, sa = ref.calibrate_batch(scaling, t, r, in_place=in_place)
                b, sb = port.calibrate_batch(scaling, t, r, in_place=in_place)
                assert np.array_equal(sa, sb) and np.all(sa == 0)
                assert np.array_equal(a.view(np.uint32), b.view(np.uint32)), (n, in_place)
