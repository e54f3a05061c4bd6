"""GPU parity of the bit operations (csrc/bool_filters.cu::BitOperationKernel) with the oracle:
bit-exact, the reference's known answers, argument rules, in place, device variant."""
import ctypes as C

import numpy as np
import pytest

from sakura_b200.capi import STATUS_OK, aligned_copy, aligned_empty
from test_oracle_bit_operations import DATA, EDIT, answers

pytestmark = pytest.mark.gpu
INVALID = 2


@pytest.mark.parametrize("dtype", [np.uint8, np.uint32])
@pytest.mark.parametrize("n", [8, 5, 11, 16, 0, 4099])
def test_known_answers(lib, dtype, n):
    data = np.resize(np.array(DATA, dtype=dtype), n)
    edit = np.resize(np.array(EDIT), n)
    for op, ans in answers(dtype).items():
        got, st = lib.bit_operation(op, 2, data, edit)
        assert st == STATUS_OK
        assert np.array_equal(got, np.resize(np.array(ans, dtype=dtype), n)), op
        got, st = lib.bit_operation(op, 2, data, edit, in_place=True)   # test :683
        assert st == STATUS_OK and np.array_equal(got, np.resize(np.array(ans, dtype=dtype), n)), op


def test_random_against_oracle(lib, oracle):
    rng = np.random.default_rng(8)
    for dtype, hi in ((np.uint8, 256), (np.uint32, 2 ** 32)):
        for n in (1, 7, 64, 1001, 1 << 18):
            data = rng.integers(0, hi, n, dtype=np.uint64).astype(dtype)
            edit = rng.integers(0, 2, n).astype(bool)
            for op in answers(dtype):
                for bm in (0, 0x80, hi - 1, 0x5A):
                    g, sg = lib.bit_operation(op, bm, data, edit)
                    o, so = oracle.bit_operation(op, bm, data, edit)
                    assert sg == so == 0 and np.array_equal(g, o), (dtype, n, op, bm)


def test_argument_rules_and_device_variant(lib, oracle):
    import torch
    L = lib.lib
    d = aligned_copy(np.arange(64, dtype=np.uint8))
    m = aligned_copy(np.ones(64, dtype=bool))
    r = aligned_empty(64, np.uint8)
    p = lambda a, off=0: C.c_void_p(a.ctypes.data + off)
    f = L.sakura_OperateBitwiseOrUint8
    assert f(128, 64, p(d), p(m), p(r)) == STATUS_OK
    assert f(128, 64, None, p(m), p(r)) == INVALID          # test :696
    assert f(128, 64, p(d), None, p(r)) == INVALID
    assert f(128, 64, p(d), p(m), None) == INVALID
    assert f(128, 63, p(d, 1), p(m), p(r)) == INVALID       # test :709
    assert f(128, 63, p(d), p(m, 1), p(r)) == INVALID
    assert L.sakura_OperateBitwiseNotUint32(4, p(d, 4), p(m), p(r)) == INVALID
    # mask -> flag bytes as bench/e2eReduce/src/main.cc:59-78: invert the mask, OR bit 7 where set
    rng = np.random.default_rng(2)
    n = 100003
    mask = rng.random(n + 3) < 0.9
    flags = rng.integers(0, 4, n + 3).astype(np.uint8)
    for off in (0, 1):   # aligned and unaligned device arrays
        d_mask = torch.from_numpy(mask).cuda()
        d_flags = torch.from_numpy(flags).cuda()
        inv = torch.empty_like(d_mask)
        vp = C.c_void_p
        assert L.sakura_InvertBoolDevice(n, vp(d_mask.data_ptr() + off), None, vp(inv.data_ptr() + off), None) == STATUS_OK
        assert L.sakura_OperateBitwiseUint8Device(4, 0x80, n, vp(d_flags.data_ptr() + off), vp(inv.data_ptr() + off),
                                                  vp(d_flags.data_ptr() + off), None) == STATUS_OK
        torch.cuda.synchronize()
        want, so = oracle.bit_operation("or", 0x80, flags[off:off + n], ~mask[off:off + n])
        assert so == 0
        assert np.array_equal(d_flags.cpu().numpy()[off:off + n], want)
    assert lib.last_kernel_name() == "bit_operation"
