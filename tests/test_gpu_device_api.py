"""Device-array binding (sakura_b200/device_api.py): torch CUDA tensors through the *BatchDevice
entry points give the same results as host arrays through the *Batch entry points (same kernels),
including padded rows (row stride > channels)."""
import numpy as np
import pytest

from sakura_b200 import synth
from sakura_b200.capi import STAT_DTYPE, STATUS_OK

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def dev(lib):
    from sakura_b200.device_api import DeviceFits
    return DeviceFits(lib)


def to_cuda(data, mask, pad=0):
    import torch
    R, n = data.shape
    d = torch.zeros((R, n + pad), dtype=torch.float32, device="cuda")
    m = torch.zeros((R, n + pad), dtype=torch.bool, device="cuda")
    d[:, :n] = torch.from_numpy(data)
    m[:, :n] = torch.from_numpy(mask)
    return d[:, :n], m[:, :n]


def same(res, host, keys):
    import torch
    torch.cuda.synchronize()
    for k in keys:
        got = res[k].cpu().numpy()
        ref = host[k]
        if k == "boundary":
            ref = ref.astype(np.int64)
        assert np.array_equal(got, ref, equal_nan=True) if got.dtype.kind == "f" else np.array_equal(got, ref), k


@pytest.mark.parametrize("pad", [0, 32])
def test_polynomial_device_arrays(lib, dev, pad):
    import torch
    data, mask = synth.make_spectra(64, 4096, seed=5)
    st, ctx = lib.create_polynomial_context(5, 4096)
    assert st == STATUS_OK
    host = lib.lsqfit_polynomial_batch(ctx, 5, data, mask, 3.0, 5, want_best_fit=True)
    d, m = to_cuda(data, mask, pad)
    out = None
    if pad:
        # the C ABI has one row stride for data/best_fit/residual and one for mask/final_mask
        out = {"residual": torch.zeros((64, 4096 + pad), dtype=torch.float32, device="cuda")[:, :4096],
               "best_fit": torch.zeros((64, 4096 + pad), dtype=torch.float32, device="cuda")[:, :4096],
               "final_mask": torch.zeros((64, 4096 + pad), dtype=torch.bool, device="cuda")[:, :4096]}
    res = dev.polynomial(ctx, 5, d, m, 3.0, 5, want_best_fit=True, out=out)
    assert res["call_status"] == STATUS_OK
    same(res, host, ["coeff", "best_fit", "residual", "final_mask", "rms", "status", "lsqfit_status"])
    lib.destroy_context(ctx)


def test_mismatched_output_stride_is_rejected(lib, dev):
    import torch
    data, mask = synth.make_spectra(8, 256, seed=6)
    st, ctx = lib.create_polynomial_context(2, 256)
    d, m = to_cuda(data, mask, 16)
    with pytest.raises(ValueError):
        dev.polynomial(ctx, 2, d, m, 3.0, 2)  # default outputs are dense, inputs are padded
    with pytest.raises(TypeError):
        dev.polynomial(ctx, 2, data, mask, 3.0, 2)  # host arrays
    lib.destroy_context(ctx)


def test_spline_and_sinusoid_device_arrays(lib, dev):
    data, mask = synth.make_spectra(32, 2048, seed=7, ripple=True)
    d, m = to_cuda(data, mask)
    st, ctx = lib.create_cubicspline_context(10, 2048)
    host = lib.lsqfit_cubicspline_batch(ctx, 10, data, mask, 3.0, 3, want_best_fit=True)
    res = dev.cubicspline(ctx, 10, d, m, 3.0, 3, want_best_fit=True)
    assert res["call_status"] == STATUS_OK
    same(res, host, ["coeff", "best_fit", "residual", "final_mask", "rms", "boundary", "status",
                     "lsqfit_status"])
    lib.destroy_context(ctx)
    st, ctx = lib.create_sinusoid_context(4, 2048)
    host = lib.lsqfit_sinusoid_batch(ctx, [0, 1, 2, 4], data, mask, 3.0, 3, want_best_fit=True)
    res = dev.sinusoid(ctx, [0, 1, 2, 4], d, m, 3.0, 3, want_best_fit=True)
    assert res["call_status"] == STATUS_OK
    same(res, host, ["coeff", "best_fit", "residual", "final_mask", "rms", "status", "lsqfit_status"])
    lib.destroy_context(ctx)


def test_statistics_device_arrays(lib, dev):
    import torch
    data, mask = synth.make_spectra(16, 1024, seed=8)
    d, m = to_cuda(data, mask)
    rec = dev.statistics(d, m)
    torch.cuda.synchronize()
    got = rec.cpu().numpy().view(STAT_DTYPE).reshape(-1)
    ref = lib.compute_statistics_batch(data, mask)
    for f in STAT_DTYPE.names:
        assert np.array_equal(got[f], ref[f]), f
