"""The host-side mask codec of the pipelined paths (bool bytes <-> one bit per channel), no GPU involved:
sakura_b200_PackBoolMaskRows / sakura_b200_UnpackBoolMaskRows against numpy's packbits."""
import ctypes as C

import numpy as np
import pytest

from sakura_b200.capi import STATUS_OK


def _bind(hostlib):
    L = hostlib.lib
    L.sakura_b200_PackBoolMaskRows.argtypes = [C.c_size_t, C.c_size_t, C.c_void_p, C.c_size_t, C.c_void_p]
    L.sakura_b200_PackBoolMaskRows.restype = C.c_int
    L.sakura_b200_UnpackBoolMaskRows.argtypes = [C.c_size_t, C.c_size_t, C.c_void_p, C.c_void_p, C.c_size_t,
                                                 C.c_void_p, C.c_int32]
    L.sakura_b200_UnpackBoolMaskRows.restype = C.c_int
    return L


@pytest.mark.parametrize("rows,n,stride", [(1, 32, 32), (5, 4096, 4096), (1000, 1056, 1088), (3000, 2048, 2048)])
def test_pack_unpack_round_trip(hostlib, rows, n, stride):
    L = _bind(hostlib)
    rng = np.random.default_rng(rows + n)
    raw = np.zeros((rows, stride), np.uint8)
    raw[:, :n] = rng.integers(0, 256, (rows, n), dtype=np.uint8) * (rng.random((rows, n)) < 0.7)  # any non-zero byte is true
    raw[:, n:] = 0xAB                                                                            # padding is never read
    bits = np.zeros((rows, n // 32), np.uint32)
    assert L.sakura_b200_PackBoolMaskRows(rows, n, raw.ctypes.data, stride, bits.ctypes.data) == STATUS_OK
    want = np.packbits(raw[:, :n] != 0, axis=1, bitorder="little").view(np.uint32)
    assert np.array_equal(bits, want)
    out = np.full((rows, stride), 7, np.uint8)
    flags = rng.integers(0, 8, rows).astype(np.int32)
    assert L.sakura_b200_UnpackBoolMaskRows(rows, n, bits.ctypes.data, out.ctypes.data, stride, flags.ctypes.data,
                                            2) == STATUS_OK
    sel = (flags & 2) != 0
    assert np.array_equal(out[sel][:, :n], (raw[sel][:, :n] != 0).astype(np.uint8))
    assert np.all(out[~sel] == 7) and np.all(out[:, n:] == 7)       # unselected rows and the padding untouched
    out2 = np.full((rows, stride), 7, np.uint8)
    assert L.sakura_b200_UnpackBoolMaskRows(rows, n, bits.ctypes.data, out2.ctypes.data, stride, None, 0) == STATUS_OK
    assert np.array_equal(out2[:, :n], (raw[:, :n] != 0).astype(np.uint8))


def test_codec_argument_rules(hostlib):
    L = _bind(hostlib)
    a = np.zeros(64, np.uint8)
    b = np.zeros(2, np.uint32)
    assert L.sakura_b200_PackBoolMaskRows(1, 48, a.ctypes.data, 48, b.ctypes.data) != STATUS_OK     # not a multiple of 32
    assert L.sakura_b200_PackBoolMaskRows(1, 64, a.ctypes.data, 32, b.ctypes.data) != STATUS_OK     # stride < num_data
    assert L.sakura_b200_PackBoolMaskRows(1, 64, None, 64, b.ctypes.data) != STATUS_OK
    assert L.sakura_b200_PackBoolMaskRows(0, 64, None, 64, None) == STATUS_OK


def test_concurrent_callers_share_the_worker_pool(hostlib):
    """One pipeline thread per GPU packs / unpacks at the same time (MultiDevice entry points): several callers
    in flight on the shared workers, every result exact."""
    import threading
    L = _bind(hostlib)
    rows, n = 1500, 4096
    errors = []

    def worker(seed):
        rng = np.random.default_rng(seed)
        for _ in range(6):
            raw = (rng.random((rows, n)) < 0.5).astype(np.uint8) * 3
            bits = np.zeros((rows, n // 32), np.uint32)
            back = np.zeros((rows, n), np.uint8)
            if L.sakura_b200_PackBoolMaskRows(rows, n, raw.ctypes.data, n, bits.ctypes.data) != STATUS_OK:
                errors.append("pack status")
            if L.sakura_b200_UnpackBoolMaskRows(rows, n, bits.ctypes.data, back.ctypes.data, n, None, 0) != STATUS_OK:
                errors.append("unpack status")
            if not np.array_equal(back, (raw != 0).astype(np.uint8)):
                errors.append("mismatch")

    threads = [threading.Thread(target=worker, args=(s,)) for s in range(8)]
    for t in threads:
        t.start()
    for t in threads:
        t.join()
    assert errors == []
