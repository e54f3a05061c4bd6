"""Boolean-filter builders: the oracles against the reference's own known-answer
tests (libsakura/test/bool_filter_collection.cc:38-200, 930-1000, re-expressed without gtest), and
the C restatement against the reference's object code on random and pathological inputs."""
import numpy as np
import pytest

F_DATA = np.array([0.0, -0.5, -1.0, -0.5, 0.0, 0.5, 1.0, 0.5, 0.0], dtype=np.float32)  # test :717
I_DATA = np.array([0, -5, -10, -5, 0, 5, 10, 5, 0], dtype=np.int32)                     # test :737
BOUNDS = {np.float32: ([-0.75, 0.5], [-0.25, 0.75]), np.int32: ([-7, 5], [-3, 7])}
RANGE_ANSWERS = {  # test :84-125; the 2 real ranges come last, padding ranges lie at 1000 + 3 i
    "ranges_inclusive": {0: [0] * 9, 1: [0, 1, 0, 1, 0, 0, 0, 0, 0], "many": [0, 1, 0, 1, 0, 1, 0, 1, 0]},
    "ranges_exclusive": {0: [0] * 9, 1: [0, 1, 0, 1, 0, 0, 0, 0, 0], "many": [0, 1, 0, 1, 0, 0, 0, 0, 0]},
}
COMPARE_ANSWERS = {  # test :187-200, threshold 0
    "gt": [0, 0, 0, 0, 0, 1, 1, 1, 0], "ge": [1, 0, 0, 0, 1, 1, 1, 1, 1],
    "lt": [0, 1, 1, 1, 0, 0, 0, 0, 0], "le": [1, 1, 1, 1, 1, 0, 0, 0, 1],
}


def bounds_for(dtype, num_condition):
    """GetBounds of the reference test (:674-692)."""
    lo0, hi0 = BOUNDS[dtype]
    pad = max(num_condition - 2, 0)
    lo = [1000 + 3 * i for i in range(pad)] + lo0[:num_condition - pad]
    hi = [1000 + 3 * i + 2 for i in range(pad)] + hi0[:num_condition - pad]
    return np.array(lo, dtype=dtype), np.array(hi, dtype=dtype)


def repeat_to(base, n):
    return np.resize(base, n) if n else base[:0]


def oracles(request):
    kinds = [request.getfixturevalue("oracle_port")]
    try:
        kinds.append(request.getfixturevalue("oracle_ref"))
    except pytest.skip.Exception:
        pass
    return kinds


@pytest.mark.parametrize("dtype,base", [(np.float32, F_DATA), (np.int32, I_DATA)])
@pytest.mark.parametrize("n", [5, 9, 11, 16, 0, 1 << 12])
def test_known_answers_ranges_and_thresholds(request, dtype, base, n):
    data = repeat_to(base, n)
    for ora in oracles(request):
        for op, table in RANGE_ANSWERS.items():
            for nc in range(0, 18):
                lo, hi = bounds_for(dtype, nc)
                got, st = ora.bool_filter(op, data, lo, hi)
                assert st == 0
                answer = np.array(table[nc] if nc in table else table["many"], dtype=bool)
                assert np.array_equal(got, repeat_to(answer, n)), (ora.kind, op, nc)
        for op, answer in COMPARE_ANSWERS.items():
            got, st = ora.bool_filter(op, data, threshold=0)
            assert st == 0
            assert np.array_equal(got, repeat_to(np.array(answer, dtype=bool), n)), (ora.kind, op)


def test_known_answers_simple_filters(request):
    f = np.array([0.0, np.inf, -1.0, -0.5, -np.inf, np.nan, 1.0, 0.5, 0.0], dtype=np.float32)
    u8 = np.array([0, 1, 2, 4, 8, 16, 32, 64, 128], dtype=np.uint8)
    u32 = np.array([0, 1, 1 << 1, 1 << 3, 1 << 7, 1 << 11, 1 << 16, 1 << 24, 1 << 30], dtype=np.uint32)
    b = np.array([0, 1, 0, 0, 1, 1, 0, 0, 1], dtype=bool)
    for ora in oracles(request):
        for n in (5, 9, 11, 16, 64, 77, 0):
            got, st = ora.bool_filter("finite", repeat_to(f, n))
            assert st == 0 and np.array_equal(got, repeat_to(np.isfinite(f), n))
            got, st = ora.bool_filter("nonzero", repeat_to(u8, n))
            assert st == 0 and np.array_equal(got, repeat_to(u8 != 0, n))
            got, st = ora.bool_filter("nonzero", repeat_to(u32, n))
            assert st == 0 and np.array_equal(got, repeat_to(u32 != 0, n))
            got, st = ora.bool_filter("invert", repeat_to(b, n))
            assert st == 0 and np.array_equal(got, repeat_to(~b, n))


def test_port_equals_reference_on_random_and_pathological_inputs(oracle_port, oracle_ref):
    rng = np.random.default_rng(11)
    for n in (1, 7, 8, 9, 63, 1000, 4099):
        f = rng.standard_normal(n).astype(np.float32)
        i = rng.integers(-50, 50, n).astype(np.int32)
        # values whose product form under- or overflows: the scalar remainder differs from the SIMD body
        f[-3:] = np.array([1e-30, -1e-30, 3e38], dtype=np.float32)[: min(3, n)]
        if n > 20:  # NaN is "in range" for the SIMD body's unordered comparisons, not for the remainder
            f[5], f[6], f[7], f[-5] = np.nan, np.inf, -np.inf, np.nan
        for nc in (0, 1, 2, 5, 16, 17, 40):
            lo_f = np.sort(rng.standard_normal(nc).astype(np.float32))
            hi_f = lo_f + rng.random(nc).astype(np.float32)
            lo_i = rng.integers(-40, 30, nc).astype(np.int32)
            hi_i = lo_i + rng.integers(0, 9, nc).astype(np.int32)
            if nc:
                lo_f[0], hi_f[0] = np.float32(0.0), np.float32(0.0)
            for op in ("ranges_inclusive", "ranges_exclusive"):
                a, sa = oracle_port.bool_filter(op, f, lo_f, hi_f)
                b, sb = oracle_ref.bool_filter(op, f, lo_f, hi_f)
                assert sa == sb == 0 and np.array_equal(a, b), (op, n, nc, "float")
                a, sa = oracle_port.bool_filter(op, i, lo_i, hi_i)
                b, sb = oracle_ref.bool_filter(op, i, lo_i, hi_i)
                assert sa == sb == 0 and np.array_equal(a, b), (op, n, nc, "int")
        for op in ("gt", "ge", "lt", "le"):
            for thr in (0, 1):
                assert np.array_equal(oracle_port.bool_filter(op, f, threshold=thr)[0],
                                      oracle_ref.bool_filter(op, f, threshold=thr)[0])
                assert np.array_equal(oracle_port.bool_filter(op, i, threshold=thr)[0],
                                      oracle_ref.bool_filter(op, i, threshold=thr)[0])
        raw = rng.integers(0, 2 ** 32, n, dtype=np.uint64).astype(np.uint32)
        assert np.array_equal(oracle_port.bool_filter("finite", raw.view(np.float32))[0],
                              oracle_ref.bool_filter("finite", raw.view(np.float32))[0])
        u8 = rng.integers(0, 3, n).astype(np.uint8)
        for op in ("nonzero", "invert"):
            assert np.array_equal(oracle_port.bool_filter(op, u8)[0].view(np.uint8),
                                  oracle_ref.bool_filter(op, u8)[0].view(np.uint8))
