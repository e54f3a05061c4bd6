"""Device-array binding (sakura_b200/device_api.py): torch CUDA tensors through the *BatchDevice
entry points give the same results as host arrays through the *Batch entry points (same kernels),
including padded rows (row stride > channels) -- and, independently of the host path, the results the
ORACLE computes for the same inputs (tests/parity.py gates)."""
import numpy as np
import pytest

from sakura_b200 import synth
from sakura_b200.capi import STAT_DTYPE, STATUS_OK
from parity import compare_fit

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


def as_host(res):
    """Device results as the dict of numpy arrays tests/parity.py compares."""
    import torch
    torch.cuda.synchronize()
    out = {k: (v.cpu().numpy() if hasattr(v, "cpu") else v) for k, v in res.items()}
    if "boundary" in out:
        out["boundary"] = out["boundary"].astype(np.uint64)
    return out


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
def test_polynomial_device_arrays(lib, dev, oracle, pad):
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
    o = oracle.lsqfit_polynomial_batch(5, data, mask, 3.0, 5, want_best_fit=True)
    compare_fit(as_host(res), o, data, 3.0, poly_rescale_n=4096)
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


def test_spline_and_sinusoid_device_arrays(lib, dev, oracle):
    data, mask = synth.make_spectra(32, 2048, seed=7, ripple=True)
    d, m = to_cuda(data, mask)
    st, ctx = lib.create_cubicspline_context(10, 2048)
    host = lib.lsqfit_cubicspline_batch(ctx, 10, data, mask, 3.0, 3, want_best_fit=True)
    res = dev.cubicspline(ctx, 10, d, m, 3.0, 3, want_best_fit=True)
    assert res["call_status"] == STATUS_OK
    same(res, host, ["coeff", "best_fit", "residual", "final_mask", "rms", "boundary", "status",
                     "lsqfit_status"])
    o = oracle.lsqfit_cubicspline_batch(10, data, mask, 3.0, 3, want_best_fit=True)
    got = as_host(res)
    assert np.array_equal(got["boundary"], o["boundary"])
    # (spline coefficients are ill-conditioned: gated through residual / mask / rms)
    compare_fit(got, o, data, 3.0, check_coeff=False, resid_atol=5e-5)
    lib.destroy_context(ctx)
    st, ctx = lib.create_sinusoid_context(4, 2048)
    host = lib.lsqfit_sinusoid_batch(ctx, [0, 1, 2, 4], data, mask, 3.0, 3, want_best_fit=True)
    res = dev.sinusoid(ctx, [0, 1, 2, 4], d, m, 3.0, 3, want_best_fit=True)
    assert res["call_status"] == STATUS_OK
    same(res, host, ["coeff", "best_fit", "residual", "final_mask", "rms", "status", "lsqfit_status"])
    o = oracle.lsqfit_sinusoid_batch([0, 1, 2, 4], data, mask, 3.0, 3, want_best_fit=True)
    compare_fit(as_host(res), o, data, 3.0)
    lib.destroy_context(ctx)


def test_statistics_device_arrays(lib, dev, oracle):
    import torch
    data, mask = synth.make_spectra(16, 1024, seed=8)
    d, m = to_cuda(data, mask)
    rec = dev.statistics(d, m)
    torch.cuda.synchronize()
    got = rec.cpu().numpy().view(STAT_DTYPE).reshape(-1)
    ref = lib.compute_statistics_batch(data, mask)
    for f in STAT_DTYPE.names:
        assert np.array_equal(got[f], ref[f]), f
    o, _ = oracle.statistics_batch(data, mask)
    assert np.array_equal(got["count"], o["count"])
    for f in ("min", "max", "index_of_min", "index_of_max"):
        assert np.array_equal(got[f], o[f]), f
    for f in ("sum", "square_sum"):
        assert np.all(np.abs(got[f] - o[f]) <= 1e-6 * np.abs(o[f]) + 1e-12), f


@pytest.mark.parametrize("n,nfit,kernel", [(4096, 5, "poly_warp"), (8192, 4, "poly_screen"), (1024, 3, "poly_fast"),
                                           (4096, 1, "poly_fast")])
def test_residual_in_place_of_data(lib, dev, n, nfit, kernel):
    """residual == data (the reference documents in-place use, sakura.h:1755-1766): the row-resident
    kernels read a row completely before writing it; results equal the out-of-place call bit for bit.
    Any other overlap, or a shape the general kernel fits, is rejected."""
    import torch
    from sakura_b200.capi import STATUS_INVALID_ARGUMENT
    data, mask = synth.make_spectra(300, n, seed=n + nfit)
    synth.add_edge_case_rows(data, mask, 6)
    d, m = to_cuda(data, mask)
    st, ctx = lib.create_polynomial_context(5, n)
    ref = dev.polynomial(ctx, 5, d, m, 3.0, nfit)
    assert ref["call_status"] == STATUS_OK and lib.last_kernel_name() == kernel
    torch.cuda.synchronize()
    work = d.clone()
    out = dev.polynomial(ctx, 5, work, m, 3.0, nfit, out={"residual": work})
    assert out["call_status"] == STATUS_OK and lib.last_kernel_name() == kernel
    torch.cuda.synchronize()
    good = ref["status"] == 0
    for k in ("status", "lsqfit_status", "final_mask", "rms", "coeff"):
        assert torch.equal(out[k][good] if k not in ("status", "lsqfit_status") else out[k],
                           ref[k][good] if k not in ("status", "lsqfit_status") else ref[k]), k
    assert torch.equal(work[good].view(torch.int32), ref["residual"][good].view(torch.int32))
    assert torch.equal(work[~good].view(torch.int32), d[~good].view(torch.int32))   # failing rows: data untouched
    # best_fit in place as well, but not both
    work = d.clone()
    out = dev.polynomial(ctx, 5, work, m, 3.0, nfit, want_best_fit=True, out={"best_fit": work})
    assert out["call_status"] == STATUS_OK
    torch.cuda.synchronize()
    assert torch.equal((d[good] - work[good]).view(torch.int32), ref["residual"][good].view(torch.int32))
    work = d.clone()
    assert dev.polynomial(ctx, 5, work, m, 3.0, nfit, want_best_fit=True,
                          out={"best_fit": work, "residual": work})["call_status"] == STATUS_INVALID_ARGUMENT
    # a partial overlap (residual one row further down the same buffer)
    big = torch.empty((301, n), dtype=torch.float32, device="cuda")
    big[:300] = d
    assert dev.polynomial(ctx, 5, big[:300], m, 3.0, nfit, out={"residual": big[1:301]})["call_status"] == STATUS_INVALID_ARGUMENT
    lib.destroy_context(ctx)
    # order 9 goes through the general kernel, which re-reads data: in place is refused
    st, ctx9 = lib.create_polynomial_context(9, n)
    work = d.clone()
    assert dev.polynomial(ctx9, 9, work, m, 3.0, 2, out={"residual": work})["call_status"] == STATUS_INVALID_ARGUMENT
    lib.destroy_context(ctx9)
