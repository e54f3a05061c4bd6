"""GPU calibration: bit-exact against the oracle; the fit with the calibration
fused into its load is identical to calibrating, NaN/Inf-flagging and fitting in sequence."""
import numpy as np
import pytest

from parity import forced_kernel

from sakura_b200 import synth
from sakura_b200.capi import STATUS_OK, aligned_copy

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def dev(lib):
    from sakura_b200.device_api import DeviceFits
    return DeviceFits(lib)


def on_off(rows, n, seed):
    """ON and OFF spectra of a position-switch pair and Tsys-like scaling factors."""
    g = np.random.Generator(np.random.Philox(seed))
    data, mask = synth.make_spectra(rows, n, seed=seed)
    off = (300.0 + 20.0 * g.standard_normal((rows, n))).astype(np.float32)
    on = (off * (1.0 + data / 250.0)).astype(np.float32)
    tsys = g.uniform(80.0, 250.0, (rows, n)).astype(np.float32)
    return on, off, tsys, mask


def bits(a):
    """Bit patterns with every NaN mapped to one value (IEEE 754 leaves NaN payloads open: x86
    propagates the operand's payload, the GPU returns the canonical quiet NaN)."""
    a = np.ascontiguousarray(a, dtype=np.float32)
    b = a.view(np.uint32).copy()
    b[np.isnan(a)] = 0x7fc00000
    return b


def test_single_array_entry_points(lib, oracle):
    # the reference's known answers (test/normalization.cc) and random data, host pointers
    st, res = lib.calibrate([0.5, 2.0], [1.0, 1.0], [1.0, 0.5])
    assert st == STATUS_OK and np.array_equal(res, np.float32([0.0, 2.0]))
    st, res = lib.calibrate(0.5, [1.0, 1.0], [1.0, 0.5], in_place=True)
    assert st == STATUS_OK and np.array_equal(res, np.float32([0.0, 0.5]))
    st, res = lib.calibrate(1.0, [1.0, -1.0], [0.0, 0.0])
    assert st == STATUS_OK and res[0] == np.inf and res[1] == -np.inf
    assert lib.last_kernel_name() == "calibrate"
    on, off, tsys, _ = on_off(1, 4099, 3)
    on[0, 5], off[0, 9] = np.nan, 0.0
    for scaling in (tsys[0], 117.5):
        st, res = lib.calibrate(scaling, on[0], off[0])
        ref, sr = oracle.calibrate_batch(scaling, on, off)
        assert st == STATUS_OK and sr[0] == 0
        assert np.array_equal(bits(res), bits(ref[0]))
    # argument rules (normalization.cc:224-243): zero length is kOK, NULL / unaligned are invalid
    assert lib.lib.sakura_CalibrateDataWithConstScalingFloat(1.0, 0, None, None, None) == STATUS_OK
    a = aligned_copy(np.ones(16, np.float32))
    assert lib.lib.sakura_CalibrateDataWithConstScalingFloat(1.0, 8, a.ctypes.data, None, a.ctypes.data) == 2
    assert lib.lib.sakura_CalibrateDataWithConstScalingFloat(1.0, 8, a.ctypes.data + 4, a.ctypes.data,
                                                             a.ctypes.data) == 2
    assert lib.lib.sakura_CalibrateDataWithArrayScalingFloat(8, None, a.ctypes.data, a.ctypes.data,
                                                             a.ctypes.data) == 2


def test_batch_device_matches_oracle(lib, dev, oracle):
    import torch
    on, off, tsys, _ = on_off(37, 2048, 4)
    off[3, 100] = 0.0
    on[4, 7] = np.inf
    d, r, s = (torch.from_numpy(x).cuda() for x in (on, off, tsys))
    for scaling, sc_dev in ((tsys, s), (93.0, 93.0), (tsys[:1], s[:1])):
        out = dev.calibrate(sc_dev, d, r)
        torch.cuda.synchronize()
        ref, st = oracle.calibrate_batch(scaling, on, off)
        assert np.all(st == 0)
        assert np.array_equal(bits(out.cpu().numpy()), bits(ref))
    # one reference row shared by all spectra, result in place over the data
    out = dev.calibrate(2.0, d.clone(), r[:1])
    ref, _ = oracle.calibrate_batch(2.0, on, off[:1])
    assert np.array_equal(bits(out.cpu().numpy()), bits(ref))
    dd = d.clone()
    dev.calibrate(s, dd, r, out=dd)
    ref, _ = oracle.calibrate_batch(tsys, on, off)
    assert np.array_equal(bits(dd.cpu().numpy()), bits(ref))


@pytest.mark.parametrize("n,order,kernel,in_kernel", [
    (4096, 5, "poly_warp", True), (2048, 5, "poly_warp", True), (2048, 2, "poly_warp", True),
    (4096, 5, "poly_warp", False), (2048, 5, "poly_warp", False),
    (1008, 3, "poly_fast", True), (8192, 7, "poly_screen", True), (1024, 9, "general", False)])
def test_fused_calibration_equals_sequence(lib, dev, oracle, n, order, kernel, in_kernel, monkeypatch):
    # in_kernel: the fit kernel calibrates while it loads (for the warp-kernel shapes only on request:
    # there the library's default is a calibration pass at HBM speed in front of the plain fit)
    import torch
    if kernel == "poly_warp" and in_kernel:
        monkeypatch.setenv("SAKURA_B200_CAL_FUSED", "1")
    rows = 48
    on, off, tsys, mask = on_off(rows, n, 11 + n)
    off[2, 17] = 0.0        # division by zero -> Inf in the calibrated spectrum
    on[5, 33] = np.nan      # NaN in the input
    mask[2, 17] = True
    mask[5, 33] = True
    d, r, s, m = (torch.from_numpy(x).cuda() for x in (on, off, tsys, mask))
    st, ctx = lib.create_polynomial_context(order, n)
    assert st == STATUS_OK
    fused = dev.calibrate_and_polynomial(ctx, order, s, d, r, m, 3.0, 4, flag_nan_or_inf=True,
                                         want_best_fit=True)
    assert fused["call_status"] == STATUS_OK
    assert lib.last_kernel_name() == kernel
    cal = dev.calibrate(s, d, r)
    m2 = m & torch.isfinite(cal)
    # (bit-for-bit: the sequence must run the kernel family the fused call ran)
    with forced_kernel({"poly_screen": "screen", "poly_warp": "warp"}.get(kernel, "channel")):
        seq = dev.polynomial(ctx, order, cal, m2, 3.0, 4, want_best_fit=True)
    torch.cuda.synchronize()
    for k in ("status", "lsqfit_status", "final_mask", "rms", "coeff", "residual", "best_fit"):
        a, b = fused[k].cpu().numpy(), seq[k].cpu().numpy()
        assert np.array_equal(a, b, equal_nan=True) if a.dtype.kind == "f" else np.array_equal(a, b), k
    # and the whole chain against the CPU reference: calibrate, flag NaN/Inf, fit
    ref_cal, _ = oracle.calibrate_batch(tsys, on, off)
    assert np.array_equal(bits(cal.cpu().numpy()), bits(ref_cal))
    o = oracle.lsqfit_polynomial_batch(order, ref_cal, mask & np.isfinite(ref_cal), 3.0, 4)
    good = o["status"] == 0
    assert np.array_equal(fused["status"].cpu().numpy(), o["status"])
    assert np.array_equal(fused["final_mask"].cpu().numpy()[good], o["final_mask"][good])
    lib.destroy_context(ctx)
