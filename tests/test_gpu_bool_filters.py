"""GPU parity of the boolean-filter builders (csrc/bool_filters.cu) with the
oracle: bit-exact, including the scalar-remainder / generic product form of the reference, the
argument rules, the device-array variants with the fused AND, and unaligned device arrays."""
import ctypes as C

import numpy as np
import pytest

from sakura_b200.capi import STATUS_OK, aligned_copy, aligned_empty
from test_oracle_bool_filters import (BOUNDS, COMPARE_ANSWERS, F_DATA, I_DATA, RANGE_ANSWERS, bounds_for,
                                      repeat_to)

pytestmark = pytest.mark.gpu
INVALID = 2


@pytest.mark.parametrize("dtype,base", [(np.float32, F_DATA), (np.int32, I_DATA)])
@pytest.mark.parametrize("n", [5, 9, 11, 16, 0, 4099])
def test_reference_known_answers(lib, dtype, base, n):
    data = repeat_to(base, n)
    for op, table in RANGE_ANSWERS.items():
        for nc in range(0, 18):
            lo, hi = bounds_for(dtype, nc)
            got, st = lib.bool_filter(op, data, lo, hi)
            assert st == STATUS_OK
            answer = np.array(table[nc] if nc in table else table["many"], dtype=bool)
            assert np.array_equal(got, repeat_to(answer, n)), (op, nc)
    for op, answer in COMPARE_ANSWERS.items():
        got, st = lib.bool_filter(op, data, threshold=0)
        assert st == STATUS_OK and np.array_equal(got, repeat_to(np.array(answer, dtype=bool), n))
    assert lib.last_kernel_name() == "bool_filter" or n == 0


@pytest.mark.parametrize("n", [1, 7, 8, 9, 63, 1000, 4099, 1 << 20])
def test_random_and_pathological_inputs_against_oracle(lib, oracle, n):
    rng = np.random.default_rng(n)
    f = rng.standard_normal(n).astype(np.float32)
    i = rng.integers(-50, 50, n).astype(np.int32)
    f[-3:] = np.array([1e-30, -1e-30, 3e38], dtype=np.float32)[: min(3, n)]
    if n > 20:
        f[5], f[6], f[7], f[-5] = np.nan, np.inf, -np.inf, np.nan
    for nc in (0, 1, 2, 16, 17, 40):
        lo_f = np.sort(rng.standard_normal(nc).astype(np.float32))
        hi_f = lo_f + rng.random(nc).astype(np.float32)
        lo_i = rng.integers(-40, 30, nc).astype(np.int32)
        hi_i = lo_i + rng.integers(0, 9, nc).astype(np.int32)
        if nc:
            lo_f[0], hi_f[0] = np.float32(0.0), np.float32(0.0)
        for op in ("ranges_inclusive", "ranges_exclusive"):
            for d, lo, hi in ((f, lo_f, hi_f), (i, lo_i, hi_i)):
                g, sg = lib.bool_filter(op, d, lo, hi)
                o, so = oracle.bool_filter(op, d, lo, hi)
                assert sg == so == 0
                assert np.array_equal(g, o), (op, nc, d.dtype, np.nonzero(g != o)[0][:5])
    for op in ("gt", "ge", "lt", "le"):
        for d in (f, i):
            assert np.array_equal(lib.bool_filter(op, d, threshold=1)[0], oracle.bool_filter(op, d, threshold=1)[0])
    raw = rng.integers(0, 2 ** 32, n, dtype=np.uint64).astype(np.uint32).view(np.float32)
    assert np.array_equal(lib.bool_filter("finite", raw)[0], oracle.bool_filter("finite", raw)[0])
    u8 = rng.integers(0, 3, n).astype(np.uint8)
    u32 = rng.integers(0, 2, n).astype(np.uint32) * np.uint32(1 << 31)
    assert np.array_equal(lib.bool_filter("nonzero", u8)[0], oracle.bool_filter("nonzero", u8)[0])
    assert np.array_equal(lib.bool_filter("nonzero", u32)[0], oracle.bool_filter("nonzero", u32)[0])
    assert np.array_equal(lib.bool_filter("invert", u8)[0].view(np.uint8),
                          oracle.bool_filter("invert", u8)[0].view(np.uint8))


def test_argument_rules(lib):
    """NULL or unaligned data / result / bound arrays => kInvalidArgument, also with zero conditions
    (bool_filter_collection.cc:535-600); InvertBool may run in place (test :1114)."""
    L = lib.lib
    d = aligned_copy(F_DATA)
    r = aligned_empty(16, np.bool_)
    lo = aligned_copy(np.array([-0.75, 0.5], dtype=np.float32))
    hi = aligned_copy(np.array([-0.25, 0.75], dtype=np.float32))
    p = lambda a, off=0: C.c_void_p(a.ctypes.data + off)
    fn = L.sakura_SetTrueIfInRangesInclusiveFloat
    assert fn(9, p(d), 2, p(lo), p(hi), p(r)) == STATUS_OK
    assert fn(9, None, 2, p(lo), p(hi), p(r)) == INVALID
    assert fn(9, p(d), 2, p(lo), p(hi), None) == INVALID
    assert fn(9, p(d), 2, None, p(hi), p(r)) == INVALID
    assert fn(9, p(d), 0, None, None, p(r)) == INVALID
    assert fn(9, p(d, 4), 2, p(lo), p(hi), p(r)) == INVALID
    assert fn(9, p(d), 2, p(lo, 4), p(hi), p(r)) == INVALID
    assert fn(9, p(d), 2, p(lo), p(hi), p(r, 1)) == INVALID
    assert L.sakura_SetTrueIfGreaterThanFloat(9, None, 0.0, p(r)) == INVALID
    assert L.sakura_SetFalseIfNanOrInfFloat(9, p(d, 4), p(r)) == INVALID
    assert L.sakura_Uint8ToBool(9, p(r), None) == INVALID
    b = aligned_copy(np.array([0, 1, 0, 0, 1, 1, 0, 0, 1, 1], dtype=bool))
    expect = ~b.copy()
    assert L.sakura_InvertBool(10, p(b), p(b)) == STATUS_OK
    assert np.array_equal(b, expect)


def test_device_variants_fused_and_unaligned(lib, oracle):
    import torch
    L = lib.lib
    rng = np.random.default_rng(3)
    n = 100003
    f = rng.standard_normal(n + 8).astype(np.float32)
    f[17], f[99] = np.nan, np.inf
    prev = rng.random(n + 8) < 0.7
    lo = np.array([-1.0, 0.5], dtype=np.float32)
    hi = np.array([-0.25, 2.0], dtype=np.float32)
    d_f = torch.from_numpy(f).cuda()
    d_prev = torch.from_numpy(prev).cuda()
    vp = C.c_void_p
    for off in (0, 1, 4):  # element offsets: 16-byte aligned, fully unaligned, 16-byte aligned floats only
        d_out = torch.zeros(n + 8, dtype=torch.bool, device="cuda")
        src = f[off:off + n]
        # ranges, fused with the AND into an existing mask
        rc = L.sakura_SetTrueIfInRangesFloatDevice(False, n, vp(d_f.data_ptr() + 4 * off), 2,
                                                   vp(lo.ctypes.data), vp(hi.ctypes.data),
                                                   vp(d_prev.data_ptr() + off), vp(d_out.data_ptr() + off), None)
        assert rc == STATUS_OK
        torch.cuda.synchronize()
        want = oracle.bool_filter("ranges_exclusive", src, lo, hi)[0] & prev[off:off + n]
        assert np.array_equal(d_out.cpu().numpy()[off:off + n], want), off
        # NaN/Inf flags ANDed in place into the running mask
        d_mask = d_prev.clone()
        rc = L.sakura_SetFalseIfNanOrInfFloatDevice(n, vp(d_f.data_ptr() + 4 * off), vp(d_mask.data_ptr() + off),
                                                    vp(d_mask.data_ptr() + off), None)
        assert rc == STATUS_OK
        torch.cuda.synchronize()
        want = oracle.bool_filter("finite", src)[0] & prev[off:off + n]
        assert np.array_equal(d_mask.cpu().numpy()[off:off + n], want)
    # edge flags as e2eReduce builds them (main.cc:182-207): channel ids strictly inside (edge-1, n-edge)
    nchan, edge = 4096, 32
    ids = torch.arange(nchan, dtype=torch.int32, device="cuda")
    out = torch.zeros(nchan, dtype=torch.bool, device="cuda")
    lo_i = np.array([edge - 1], dtype=np.int32)
    hi_i = np.array([nchan - edge], dtype=np.int32)
    assert L.sakura_SetTrueIfInRangesIntDevice(False, nchan, vp(ids.data_ptr()), 1, vp(lo_i.ctypes.data),
                                               vp(hi_i.ctypes.data), None, vp(out.data_ptr()), None) == STATUS_OK
    m = out.cpu().numpy()
    assert m[:edge].sum() == 0 and m[-edge:].sum() == 0 and m[edge:-edge].all()
    # uint8 flags -> bool -> inverted, compare >= with fused AND
    u8 = torch.from_numpy(rng.integers(0, 3, n).astype(np.uint8)).cuda()
    b1 = torch.zeros(n, dtype=torch.bool, device="cuda")
    assert L.sakura_Uint8ToBoolDevice(n, vp(u8.data_ptr()), None, vp(b1.data_ptr()), None) == STATUS_OK
    assert L.sakura_InvertBoolDevice(n, vp(b1.data_ptr()), None, vp(b1.data_ptr()), None) == STATUS_OK
    assert np.array_equal(b1.cpu().numpy(), u8.cpu().numpy() == 0)
    assert L.sakura_SetTrueIfComparesFloatDevice(1, n, vp(d_f.data_ptr()), C.c_float(0.25), vp(b1.data_ptr()),
                                                 vp(b1.data_ptr()), None) == STATUS_OK
    assert np.array_equal(b1.cpu().numpy(), (u8.cpu().numpy() == 0) & (f[:n] >= np.float32(0.25)))
