"""Bit operations with an edit mask (libsakura/src/bit_operation.cc): the oracles against the
reference's known answers (libsakura/test/bit_operation.cc:114-230: bit_mask 2, data 0 1 2 3 0 1 2 3,
edit_mask F F F F T T T T) and the C restatement against the reference's object code."""
import numpy as np
import pytest

DATA = [0, 1, 2, 3, 0, 1, 2, 3]
EDIT = [False] * 4 + [True] * 4
ALL = {np.uint8: 0xFF, np.uint32: 0xFFFFFFFF}


def answers(dtype):
    a = ALL[dtype]
    return {
        "and": [0, 1, 2, 3, 0, 0, 2, 2],
        "converse_nonimplication": [0, 1, 2, 3, 2, 2, 0, 0],
        "implication": [0, 1, 2, 3, a, a - 1, a, a - 1],
        "or": [0, 1, 2, 3, 2, 3, 2, 3],
        "xor": [0, 1, 2, 3, 2, 3, 0, 1],
        "not": [0, 1, 2, 3, a, a - 1, a - 2, a - 3],
    }


def oracles(request):
    kinds = [request.getfixturevalue("oracle_port")]
    try:
        kinds.append(request.getfixturevalue("oracle_ref"))
    except pytest.skip.Exception:
        pass
    return kinds


@pytest.mark.parametrize("dtype", [np.uint8, np.uint32])
@pytest.mark.parametrize("n", [8, 5, 11, 16, 0, 4099])
def test_known_answers(request, dtype, n):
    data = np.resize(np.array(DATA, dtype=dtype), n)
    edit = np.resize(np.array(EDIT), n)
    for ora in oracles(request):
        for op, ans in answers(dtype).items():
            got, st = ora.bit_operation(op, 2, data, edit)
            assert st == 0
            assert np.array_equal(got, np.resize(np.array(ans, dtype=dtype), n)), (ora.kind, op)


def test_port_equals_reference_random(oracle_port, oracle_ref):
    rng = np.random.default_rng(3)
    for dtype, hi in ((np.uint8, 256), (np.uint32, 2 ** 32)):
        for n in (1, 7, 64, 1001):
            data = rng.integers(0, hi, n, dtype=np.uint64).astype(dtype)
            edit = rng.integers(0, 2, n).astype(bool)
            for op in answers(dtype):
                for bm in (0, 1, 0x80, hi - 1, 0x5A):
                    a, sa = oracle_port.bit_operation(op, bm, data, edit)
                    b, sb = oracle_ref.bit_operation(op, bm, data, edit)
                    assert sa == sb == 0 and np.array_equal(a, b), (dtype, n, op, bm)
