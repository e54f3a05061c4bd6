"""The host-pointer paths of the C ABI (csrc/capi_fit.cu): the single-arena path of small batches
(the reference's single-row signatures), the copy/compute pipeline (all three fit types) and the
...MultiDevice entry points must deliver, row for row and bit for bit, what one upload + one kernel
+ one download deliver -- which the other test files compare with the oracle -- and are compared
with the oracle here once more on rows that fail, publish partially or publish nothing."""
import ctypes as C

import numpy as np
import pytest

from sakura_b200 import synth
from sakura_b200.capi import (LSQFIT_NOT_ENOUGH_DATA, LSQFIT_OK, STATUS_INVALID_ARGUMENT, STATUS_NG, STATUS_OK,
                              aligned_copy)
from parity import compare_fit

pytestmark = pytest.mark.gpu

KEYS = ("status", "lsqfit_status", "final_mask", "rms", "coeff", "residual", "best_fit", "boundary")


def assert_bitwise(a, b):
    assert a["call_status"] == STATUS_OK and b["call_status"] == STATUS_OK
    for key in KEYS:
        if a.get(key) is None:
            assert b.get(key) is None, key
            continue
        x, y = np.ascontiguousarray(a[key]), np.ascontiguousarray(b[key])
        assert x.shape == y.shape, key
        assert np.array_equal(x.view(np.uint8), y.view(np.uint8)), (key, int((x.view(np.uint8) != y.view(np.uint8)).sum()))


class Tuning:
    """Forces a host path: 'serial' (one chunk), 'pipeline' (1 MiB chunks), 'packed' (single arena)."""

    def __init__(self, lib, path):
        self.lib, self.path = lib, path

    def __enter__(self):
        chunk, packed = {"serial": (4096, 0), "pipeline": (1, 0), "packed": (4096, 1 << 20)}[self.path]
        self.lib.set_host_pipeline_tuning(chunk, packed)

    def __exit__(self, *exc):
        self.lib.set_host_pipeline_tuning(96, 2048)


def run(lib, kind, path, data, mask, nfit=4, clip=3.0, devices=None):
    n = data.shape[-1]
    with Tuning(lib, path):
        if kind == "poly":
            st, ctx = lib.create_polynomial_context(5, n)
            assert st == STATUS_OK
            g = lib.lsqfit_polynomial_batch(ctx, 5, data, mask, clip, nfit, want_best_fit=True, devices=devices)
        elif kind == "poly9":     # order 9: the general kernel (scratch of the context shared by the chunks)
            st, ctx = lib.create_polynomial_context(9, n)
            assert st == STATUS_OK
            g = lib.lsqfit_polynomial_batch(ctx, 9, data, mask, clip, nfit, want_best_fit=True, devices=devices)
        elif kind == "spline":
            st, ctx = lib.create_cubicspline_context(6, n)
            assert st == STATUS_OK
            g = lib.lsqfit_cubicspline_batch(ctx, 6, data, mask, clip, nfit, want_best_fit=True, devices=devices)
        else:
            st, ctx = lib.create_sinusoid_context(4, n)
            assert st == STATUS_OK
            g = lib.lsqfit_sinusoid_batch(ctx, [0, 1, 2, 4], data, mask, clip, nfit, want_best_fit=True,
                                          devices=devices)
        kernel = lib.last_kernel_name()
        lib.destroy_context(ctx)
    return g, kernel


def rows_with_failures(rows, n, K, seed):
    data, mask = synth.make_spectra(rows, n, seed=seed, ripple=True)
    for first in (0, rows // 2, rows - 6):   # failing rows in the first, a middle and the last chunk
        synth.add_edge_case_rows(data, mask, K, first_row=first)
    return data, mask


@pytest.mark.parametrize("kind,n,rows", [("poly", 4096, 1500), ("poly", 1056, 3000), ("poly9", 1024, 700),
                                         ("spline", 2048, 1200), ("sin", 2048, 1500), ("sin", 1056, 900)])
def test_pipeline_equals_single_chunk(lib, kind, n, rows):
    """Many 1 MiB chunks through the four slots against one upload / kernel / download."""
    K = {"poly": 6, "poly9": 10, "spline": 9, "sin": 7}[kind]
    data, mask = rows_with_failures(rows, n, K, seed=n + rows)
    a, ka = run(lib, kind, "serial", data, mask)
    b, kb = run(lib, kind, "pipeline", data, mask)
    assert ka == kb
    assert (a["status"] != STATUS_OK).sum() >= 6      # K-1 valid channels / all flagged, three times
    assert_bitwise(a, b)
    # untouched outputs of the failing rows are still the caller's NaN / False prefill
    bad = a["lsqfit_status"] == LSQFIT_NOT_ENOUGH_DATA
    assert bad.sum() >= 3
    assert np.all(np.isnan(b["rms"][bad])) and np.all(np.isnan(b["residual"][bad])) and not b["final_mask"][bad].any()


@pytest.mark.parametrize("kind,n", [("poly", 4096), ("poly", 2048), ("poly", 1056), ("poly9", 512), ("spline", 2048),
                                    ("sin", 2048)])
@pytest.mark.parametrize("rows", [1, 7])
def test_packed_equals_single_chunk(lib, kind, n, rows):
    K = {"poly": 6, "poly9": 10, "spline": 9, "sin": 7}[kind]
    data, mask = synth.make_spectra(max(rows, 6), n, seed=3 * n + rows, ripple=True)
    if rows >= 6:
        synth.add_edge_case_rows(data, mask, K)
    data, mask = data[:rows], mask[:rows]
    a, _ = run(lib, kind, "serial", data, mask)
    b, _ = run(lib, kind, "packed", data, mask)
    assert_bitwise(a, b)


@pytest.mark.parametrize("kind,n,rows", [("poly", 4096, 1500), ("spline", 2048, 1200)])
def test_pipeline_with_and_without_bit_packed_masks(lib, kind, n, rows, monkeypatch):
    """The pipeline moves the masks over PCIe as bits by default (host_bits.h); SAKURA_B200_PACK_MASKS=0 keeps
    them as bytes. Both against the single-chunk path, with mask bytes other than 0 / 1 ("any non-zero
    byte is true") and failing rows whose final_mask must stay the caller's."""
    K = {"poly": 6, "spline": 9}[kind]
    data, mask = rows_with_failures(rows, n, K, seed=7 * n + rows)
    odd = (mask.view(np.uint8) * np.uint8(0x84)).view(np.bool_)      # true = 0x84 instead of 0x01
    a, _ = run(lib, kind, "serial", data, mask)
    for packed in ("1", "0"):
        monkeypatch.setenv("SAKURA_B200_PACK_MASKS", packed)
        # (as bytes the general kernel hands a row's mask bytes through as they are: 0 / 1 input there)
        b, _ = run(lib, kind, "pipeline", data, odd if packed == "1" else mask)
        assert_bitwise(a, b)
        bad = a["lsqfit_status"] == LSQFIT_NOT_ENOUGH_DATA
        assert bad.sum() >= 3 and not b["final_mask"][bad].any()
        assert set(np.unique(b["final_mask"].view(np.uint8))) <= {0, 1}


def test_host_paths_against_oracle(lib, oracle):
    data, mask = rows_with_failures(1200, 4096, 6, seed=99)
    o = oracle.lsqfit_polynomial_batch(5, data, mask, 3.0, 4, want_best_fit=True)
    for path in ("pipeline", "serial"):
        g, _ = run(lib, "poly", path, data, mask)
        assert compare_fit(g, o, data, 3.0, poly_rescale_n=4096) == []
    g, _ = run(lib, "poly", "packed", data[:12], mask[:12])
    o12 = {k: (v[:12] if isinstance(v, np.ndarray) else v) for k, v in o.items() if k != "refit"}
    compare_fit(g, o12, data[:12], 3.0, poly_rescale_n=4096)


def test_single_row_reference_signatures_use_the_packed_path(lib, oracle):
    """sakura_LSQFitPolynomialFloat (the reference's own signature) = a batch of one row: every output
    against the oracle's, including a row that fails (outputs untouched)."""
    n = 4096
    data, mask = synth.make_spectra(8, n, seed=5)
    names = synth.add_edge_case_rows(data, mask, 6)
    st, ctx = lib.create_polynomial_context(5, n)
    assert st == STATUS_OK
    o = oracle.lsqfit_polynomial_batch(5, data, mask, 3.0, 5, want_best_fit=True)
    for r in range(8):
        d, m = aligned_copy(data[r]), aligned_copy(mask[r])
        coeff = aligned_copy(np.full(6, np.nan))
        best, resid = aligned_copy(np.full(n, np.nan, np.float32)), aligned_copy(np.full(n, np.nan, np.float32))
        fm = aligned_copy(np.zeros(n, np.bool_))
        rms, lsq = C.c_float(np.nan), C.c_int(-1)
        st = lib.lib.sakura_LSQFitPolynomialFloat(ctx, 5, n, d.ctypes.data, m.ctypes.data, 3.0, 5, 6, coeff.ctypes.data,
                                                  best.ctypes.data, resid.ctypes.data, fm.ctypes.data, C.byref(rms),
                                                  C.byref(lsq))
        assert st == o["status"][r] and lsq.value == o["lsqfit_status"][r], r
        if st == STATUS_OK:
            g = {"call_status": 0, "status": np.array([st], np.int32), "lsqfit_status": np.array([lsq.value], np.int32),
                 "final_mask": fm[None], "rms": np.array([rms.value], np.float32), "coeff": coeff[None],
                 "residual": resid[None], "best_fit": best[None]}
            orow = {k: (v[r:r + 1] if isinstance(v, np.ndarray) else v) for k, v in o.items() if k != "refit"}
            compare_fit(g, orow, data[r:r + 1], 3.0, poly_rescale_n=n)
        else:
            assert r in (names["K_minus_1"], names["all_false"])
            assert np.isnan(rms.value) and np.all(np.isnan(resid)) and np.all(np.isnan(coeff)) and not fm.any()
    lib.destroy_context(ctx)


@pytest.mark.parametrize("kind,n,rows", [("poly", 4096, 1500), ("spline", 2048, 700), ("sin", 2048, 1100)])
def test_multi_device_equals_single_device(lib, kind, n, rows):
    """Rows split over every visible GPU (one thread, context replica and pipeline per GPU) against the
    current-device call; with one visible GPU this still runs the replica / shard code."""
    import torch
    K = {"poly": 6, "spline": 9, "sin": 7}[kind]
    data, mask = rows_with_failures(rows, n, K, seed=11 * n + rows)
    a, _ = run(lib, kind, "serial", data, mask)
    for devices in ("all", [0], list(range(torch.cuda.device_count()))[::-1]):
        for path in ("serial", "pipeline"):
            b, _ = run(lib, kind, path, data, mask, devices=devices)
            assert_bitwise(a, b)
    assert torch.cuda.current_device() == 0


def test_multi_device_argument_rules(lib):
    import torch
    data, mask = synth.make_spectra(8, 1024, seed=1)
    st, ctx = lib.create_polynomial_context(3, 1024)
    assert st == STATUS_OK
    nvis = torch.cuda.device_count()
    for devices in ([nvis], [-1], [0, 0]):
        g = lib.lsqfit_polynomial_batch(ctx, 3, data, mask, 3.0, 2, devices=devices)
        assert g["call_status"] == STATUS_INVALID_ARGUMENT, devices
    # more devices than rows: the surplus devices get nothing
    g = lib.lsqfit_polynomial_batch(ctx, 3, data[:1], mask[:1], 3.0, 2, devices="all")
    h = lib.lsqfit_polynomial_batch(ctx, 3, data[:1], mask[:1], 3.0, 2)
    assert_bitwise(g, h)
    lib.destroy_context(ctx)


@pytest.mark.parametrize("packed", ["1", "0"])
def test_pipeline_in_place_outputs(lib, packed, monkeypatch):
    """residual in place of data and final_mask in place of mask (the reference allows both, sakura.h:1755-1766)
    through the chunked pipeline: a chunk's results come back while later chunks are still being read from the
    same arrays (and, with bit-packed masks, while worker threads pack / expand other rows of them)."""
    monkeypatch.setenv("SAKURA_B200_PACK_MASKS", packed)
    rows, n = 1500, 4096
    data, mask = rows_with_failures(rows, n, 6, seed=4242)
    a, _ = run(lib, "poly", "serial", data, mask)
    d2, m2 = aligned_copy(data), aligned_copy(mask)
    st, ctx = lib.create_polynomial_context(5, n)
    assert st == STATUS_OK
    with Tuning(lib, "pipeline"):
        out = {"residual": d2, "final_mask": m2}
        b = lib.lsqfit_polynomial_batch(ctx, 5, d2, m2, 3.0, 4, out=out)
    lib.destroy_context(ctx)
    assert b["call_status"] == STATUS_OK
    assert np.array_equal(a["status"], b["status"])
    ok = a["status"] == STATUS_OK
    assert np.array_equal(a["residual"][ok].view(np.uint32), d2[ok].view(np.uint32))
    assert np.array_equal(a["final_mask"][ok], m2[ok])
    # failing rows: data untouched; the mask only where the reference writes final_mask for them
    bad = a["lsqfit_status"] == LSQFIT_NOT_ENOUGH_DATA
    assert np.array_equal(d2[bad].view(np.uint32), data[bad].view(np.uint32))
