"""BASELINE.json full sizes on the GPU, checked through size-independent properties (the oracle
would need minutes on 4e9 channels) plus an oracle comparison of a row sample:

* every row succeeds; final_mask only ever removes channels; channel n-1 is never clipped;
* rms equals the statistics of the written residual over final_mask (recomputed with torch);
* idempotence: refitting with final_mask as the mask and one fit reproduces the coefficients;
* linearity: adding a polynomial of the fitted order to the data shifts the coefficients by it
  and leaves masks / residuals unchanged (to float rounding);
* a sample of rows equals the oracle.
"""
import numpy as np
import pytest

from sakura_b200.capi import STATUS_OK
from parity import compare_fit

pytestmark = pytest.mark.gpu


def device_fit(lib, ctx, torch, data, mask, order, clip, nfit, want_residual=True):
    R, n = data.shape
    K = order + 1
    out = {
        "coeff": torch.empty(R, K, dtype=torch.float64, device=data.device),
        "residual": torch.empty(R, n, dtype=torch.float32, device=data.device) if want_residual else None,
        "final_mask": torch.empty(R, n, dtype=torch.bool, device=data.device),
        "rms": torch.empty(R, dtype=torch.float32, device=data.device),
        "status": torch.empty(R, dtype=torch.int32, device=data.device),
        "lsqfit_status": torch.empty(R, dtype=torch.int32, device=data.device),
    }
    rc = lib.lib.sakura_LSQFitPolynomialFloatBatchDevice(
        ctx, order, R, n, data.data_ptr(), n, mask.data_ptr(), n, clip, nfit, K, out["coeff"].data_ptr(), None,
        out["residual"].data_ptr() if want_residual else None, out["final_mask"].data_ptr(),
        out["rms"].data_ptr(), out["status"].data_ptr(), out["lsqfit_status"].data_ptr(), None)
    assert rc == STATUS_OK, lib.last_error()
    torch.cuda.synchronize()
    return out


@pytest.mark.parametrize("rows,n,nfit", [(1 << 20, 4096, 5), (1 << 19, 2048, 3)])
def test_full_size_properties(lib, oracle, rows, n, nfit):
    import torch
    import bench
    dev = torch.device("cuda", 0)
    data, mask = bench.generate_on_device(torch, rows, n, seed=7, device=dev)
    st, ctx = lib.create_polynomial_context(5, n)
    out = device_fit(lib, ctx, torch, data, mask, 5, 3.0, nfit)
    assert lib.last_kernel_name() == "poly_warp"
    assert int((out["status"] != 0).sum()) == 0
    fm = out["final_mask"]
    assert not bool((fm & ~mask).any())  # clipping only removes channels
    assert bool((fm[:, n - 1] == mask[:, n - 1]).all())  # last channel never examined
    assert int((mask & ~fm).sum()) > rows  # and it did clip something
    # rms from the written residual
    chunk = 1 << 16
    worst = 0.0
    for r0 in range(0, rows, chunk):
        r = out["residual"][r0:r0 + chunk].double()
        v = fm[r0:r0 + chunk]
        cnt = v.sum(dim=1).double()
        mean = (r * v).sum(dim=1) / cnt
        var = ((r * r) * v).sum(dim=1) / cnt - mean * mean
        rms = var.abs().sqrt()
        rel = ((rms - out["rms"][r0:r0 + chunk].double()).abs() / rms).max().item()
        worst = max(worst, rel)
    assert worst < 1e-6, worst
    # idempotence: the final mask is a fixed point of one more fit
    sub = slice(0, 1 << 17)
    again = device_fit(lib, ctx, torch, data[sub], fm[sub].clone(), 5, 3.0, 1, want_residual=False)
    c0, c1 = out["coeff"][sub], again["coeff"]
    scale = torch.tensor([float(n - 1) ** k for k in range(6)], dtype=torch.float64, device=dev)
    d = ((c0 - c1) * scale).abs()
    assert bool(((d <= 1e-6) | (d <= 1e-5 * (c0 * scale).abs())).all())
    assert bool((again["final_mask"] == fm[sub]).all())
    # linearity in the data: add 3 + 2x - x^2 (exactly representable steps are not needed: compare loosely)
    x = torch.linspace(0.0, 1.0, n, dtype=torch.float64, device=dev)
    shift = (3.0 + 2.0 * x - x * x)
    sub2 = slice(0, 1 << 15)
    shifted = (data[sub2].double() + shift).float()
    sh = device_fit(lib, ctx, torch, shifted, mask[sub2], 5, 3.0, nfit)
    dc = (sh["coeff"] - out["coeff"][sub2]) * scale
    expect = torch.tensor([3.0, 2.0, -1.0, 0.0, 0.0, 0.0], dtype=torch.float64, device=dev)
    assert float((dc - expect).abs().max()) < 2e-3  # float rounding of the shifted data, cond(G)~1e7
    same = (sh["final_mask"] == fm[sub2]).float().mean().item()
    assert same > 0.99999
    # oracle on a sample of rows
    take = torch.arange(0, rows, rows // 512, device=dev)[:512]
    hd, hm = data[take].cpu().numpy(), mask[take].cpu().numpy()
    o = oracle.lsqfit_polynomial_batch(5, hd, hm, 3.0, nfit)
    g = {k: (v[take].cpu().numpy() if v is not None else None) for k, v in out.items()}
    g["call_status"] = 0
    g["best_fit"] = None
    compare_fit(g, o, hd, 3.0, poly_rescale_n=n)
    lib.destroy_context(ctx)


def _rms_from_residual(torch, out, rows):
    """rms of baseline.cc:1200-1207 recomputed with torch from the residual the kernel wrote."""
    worst = 0.0
    chunk = 1 << 14
    for r0 in range(0, rows, chunk):
        r = out["residual"][r0:r0 + chunk].double()
        v = out["final_mask"][r0:r0 + chunk]
        cnt = v.sum(dim=1).double()
        mean = (r * v).sum(dim=1) / cnt
        var = ((r * r) * v).sum(dim=1) / cnt - mean * mean
        rms = var.abs().sqrt()
        worst = max(worst, ((rms - out["rms"][r0:r0 + chunk].double()).abs() / rms).max().item())
    return worst


@pytest.mark.parametrize("kind", ["spline", "sinusoid"])
def test_full_size_properties_c3_c4(lib, oracle, kind):
    """BASELINE configs[2] / configs[3] at full size (262 144 x 8192; 20 pieces / wave numbers 0..20,
    3 fits) on device arrays: size-independent properties plus the oracle on a sample of rows."""
    import torch
    import bench
    from sakura_b200.device_api import DeviceFits
    rows, n, nfit, clip = 1 << 18, 8192, 3, 3.0
    dev = torch.device("cuda", 0)
    fits = DeviceFits(lib)
    data, mask = bench.generate_on_device(torch, rows, n, seed=31, device=dev, ripple=True)
    if kind == "spline":
        st, ctx = lib.create_cubicspline_context(20, n)
        fit = lambda d, m, f, **kw: fits.cubicspline(ctx, 20, d, m, clip, f, **kw)
        expect_kernel = "spline_moments"
    else:
        st, ctx = lib.create_sinusoid_context(20, n)
        waves = list(range(21))
        fit = lambda d, m, f, **kw: fits.sinusoid(ctx, waves, d, m, clip, f, **kw)
        expect_kernel = "sinusoid_dmma"
    assert st == STATUS_OK
    out = fit(data, mask, nfit)
    torch.cuda.synchronize()
    assert out["call_status"] == STATUS_OK, lib.last_error()
    assert lib.last_kernel_name() == expect_kernel
    assert int((out["status"] != 0).sum()) == 0
    fm = out["final_mask"]
    assert not bool((fm & ~mask).any())
    assert int((mask & ~fm).sum()) > rows
    if kind == "spline":
        b = out["boundary"]
        assert bool((b[:, 0] == 0).all()) and bool((b[:, -1] == n).all())
        assert bool((b[:, 1:] > b[:, :-1]).all())
    assert _rms_from_residual(torch, out, rows) < 1e-6
    # idempotence: the final mask is a fixed point of one more fit and reproduces the residual
    # (spline: the piece boundaries follow the mask, so only the rms is compared there)
    sub = slice(0, 1 << 15)
    again = fit(data[sub], fm[sub].clone(), 1)
    torch.cuda.synchronize()
    assert bool((again["final_mask"] == fm[sub]).all())
    if kind == "sinusoid":
        tol = 1e-5 * out["residual"][sub].abs() + 2.4e-7 * (data[sub] - out["residual"][sub]).abs() + 1e-6
        assert bool(((again["residual"] - out["residual"][sub]).abs() <= tol).all())
        assert float(((again["rms"] - out["rms"][sub]).abs() / out["rms"][sub]).max()) < 1e-6
    # run-to-run: bit-identical
    twice = fit(data[sub], mask[sub], nfit)
    torch.cuda.synchronize()
    for key in ("final_mask", "rms", "residual", "coeff", "status"):
        assert torch.equal(twice[key], out[key][sub]), key
    # the oracle on a sample of rows
    take = torch.arange(0, rows, rows // 96, device=dev)[:96]
    hd, hm = data[take].cpu().numpy(), mask[take].cpu().numpy()
    g = {k: (v[take].cpu().numpy() if isinstance(v, torch.Tensor) else v) for k, v in out.items()}
    g["best_fit"] = None
    if kind == "spline":
        o = oracle.lsqfit_cubicspline_batch(20, hd, hm, clip, nfit)
        g["boundary"] = g["boundary"].astype(np.uint64)
        assert np.array_equal(g["boundary"], o["boundary"])
        reported = compare_fit(g, o, hd, clip, check_coeff=False, resid_atol=5e-5)
    else:
        o = oracle.lsqfit_sinusoid_batch(list(range(21)), hd, hm, clip, nfit)
        reported = compare_fit(g, o, hd, clip)
    assert reported == [], reported
    lib.destroy_context(ctx)
