"""Batched fits on arrays that already live on the GPU.

The reference's python binding (libsakura/python-binding/python-binding.cc:1610-1757) wraps the
single-spectrum polynomial fit for numpy buffers; spline and sinusoid are TODOs there (:1649-1651).
Here every fit of the path is callable on anything that exposes ``__cuda_array_interface__``
(torch CUDA tensors, CuPy, Numba device arrays): rows [R, n] with a contiguous last axis, any row
stride. No data moves through the host; the calls are asynchronous on the given stream (default:
torch's current stream). Outputs are torch tensors on the inputs' device unless the caller passes
its own arrays in ``out``.

    lib = SakuraLib(); dev = DeviceFits(lib)
    st, ctx = lib.create_polynomial_context(5, n)
    res = dev.polynomial(ctx, 5, data_cuda, mask_cuda, 3.0, 5)
    res["final_mask"], res["residual"], res["coeff"], res["rms"], res["status"]
"""
from __future__ import annotations

import ctypes as C
from typing import Optional, Sequence

import numpy as np

from .capi import STATUS_OK, SakuraLib, aligned_copy

_ITEMSIZE = {"<f4": 4, "|b1": 1, "|u1": 1, "<f8": 8, "<i4": 4, "<u8": 8, "<i8": 8}


class DeviceArray:
    """Pointer, shape and row stride (in elements) of a 1-D/2-D device array."""

    def __init__(self, obj, typestrs: Sequence[str], what: str):
        cai = getattr(obj, "__cuda_array_interface__", None)
        if cai is None:
            raise TypeError(f"{what}: object has no __cuda_array_interface__ (is it on the GPU?)")
        if cai["typestr"] not in typestrs:
            raise TypeError(f"{what}: dtype {cai['typestr']} not in {tuple(typestrs)}")
        shape = tuple(cai["shape"])
        if len(shape) == 1:
            shape = (1,) + shape
        if len(shape) != 2:
            raise ValueError(f"{what}: expected [rows, channels], got shape {shape}")
        item = _ITEMSIZE[cai["typestr"]]
        strides = cai.get("strides")
        if strides is None:
            row_stride = shape[1]
        else:
            strides = tuple(strides)
            if len(strides) == 1:
                strides = (shape[1] * item,) + strides
            if strides[1] != item or strides[0] % item != 0:
                raise ValueError(f"{what}: the channel axis must be contiguous")
            row_stride = strides[0] // item
        self.ptr = int(cai["data"][0])
        self.shape = shape
        self.row_stride = int(row_stride) if shape[0] > 1 else max(int(row_stride), shape[1])
        self.obj = obj  # keeps the memory alive for the duration of the call


def _torch():
    import torch
    return torch


class DeviceFits:
    def __init__(self, lib: Optional[SakuraLib] = None):
        self.lib = lib if lib is not None else SakuraLib()

    # ------------------------------------------------------------------ helpers
    @staticmethod
    def _stream(stream) -> C.c_void_p:
        if stream is None:
            stream = _torch().cuda.current_stream().cuda_stream
        elif hasattr(stream, "cuda_stream"):
            stream = stream.cuda_stream
        return C.c_void_p(int(stream))

    @staticmethod
    def _device_of(obj):
        torch = _torch()
        if isinstance(obj, torch.Tensor):
            return obj.device
        return torch.device("cuda", torch.cuda.current_device())

    def _outputs(self, out, device, R, n, coeff_shape, want_coeff, want_best_fit, want_residual,
                 boundary_cols=0):
        torch = _torch()
        res = dict(out) if out is not None else {}

        def need(key, shape, dtype, wanted=True):
            if key in res and res[key] is not None:
                return
            res[key] = torch.empty(shape, dtype=dtype, device=device) if wanted else None

        need("coeff", (R,) + tuple(coeff_shape), torch.float64, want_coeff)
        need("best_fit", (R, n), torch.float32, want_best_fit)
        need("residual", (R, n), torch.float32, want_residual)
        need("final_mask", (R, n), torch.bool)
        need("rms", (R,), torch.float32)
        need("status", (R,), torch.int32)
        need("lsqfit_status", (R,), torch.int32)
        if boundary_cols:
            need("boundary", (R, boundary_cols), torch.int64)
        return res

    @staticmethod
    def _p(obj, typestrs, what, R=None, n=None):
        """Device pointer (and row stride) of an optional output array."""
        if obj is None:
            return None, 0
        a = DeviceArray(obj, typestrs, what)
        if R is not None and (a.shape[0] != R or (n is not None and a.shape[1] != n)):
            raise ValueError(f"{what}: shape {a.shape} does not match [{R}, {n}]")
        return C.c_void_p(a.ptr), a.row_stride

    def _row_outputs(self, res, R, n, data_stride, mask_stride):
        bf, bfs = self._p(res["best_fit"], ("<f4",), "best_fit", R, n)
        rs, rss = self._p(res["residual"], ("<f4",), "residual", R, n)
        fm, fms = self._p(res["final_mask"], ("|b1", "|u1"), "final_mask", R, n)
        for name, ptr, stride, want in (("best_fit", bf, bfs, data_stride),
                                        ("residual", rs, rss, data_stride),
                                        ("final_mask", fm, fms, mask_stride)):
            if ptr is not None and R > 1 and stride != want:
                raise ValueError(f"{name}: row stride {stride} must equal the input's ({want}); the "
                                 "C ABI uses one stride for data/best_fit/residual and one for "
                                 "mask/final_mask")
        return bf, rs, fm

    @staticmethod
    def _flat(obj, typestrs, what, count):
        a = DeviceArray(obj, typestrs, what)
        if a.shape[0] * a.shape[1] < count:
            raise ValueError(f"{what}: needs {count} elements")
        return C.c_void_p(a.ptr)

    # ------------------------------------------------------------------ fits
    def polynomial(self, ctx, order: int, data, mask, clip_threshold_sigma: float,
                   num_fitting_max: int, num_coeff: Optional[int] = None, want_coeff: bool = True,
                   want_best_fit: bool = False, want_residual: bool = True, out: Optional[dict] = None,
                   stream=None) -> dict:
        """sakura_LSQFitPolynomialFloatBatchDevice on device arrays; see the module docstring."""
        d = DeviceArray(data, ("<f4",), "data")
        m = DeviceArray(mask, ("|b1", "|u1"), "mask")
        if d.shape != m.shape:
            raise ValueError("data and mask shapes differ")
        R, n = d.shape
        K = order + 1 if num_coeff is None else num_coeff
        res = self._outputs(out, self._device_of(data), R, n, (K,), want_coeff, want_best_fit,
                            want_residual)
        bf, rs, fm = self._row_outputs(res, R, n, d.row_stride, m.row_stride)
        co = self._flat(res["coeff"], ("<f8",), "coeff", R * K) if res["coeff"] is not None else None
        res["call_status"] = self.lib.lib.sakura_LSQFitPolynomialFloatBatchDevice(
            ctx, order, R, n, C.c_void_p(d.ptr), d.row_stride, C.c_void_p(m.ptr), m.row_stride,
            clip_threshold_sigma, num_fitting_max, K, co, bf, rs, fm,
            self._flat(res["rms"], ("<f4",), "rms", R), self._flat(res["status"], ("<i4",), "status", R),
            self._flat(res["lsqfit_status"], ("<i4",), "lsqfit_status", R), self._stream(stream))
        return res

    def _calibration_args(self, scaling, reference, R, n):
        ref = DeviceArray(reference, ("<f4",), "reference")
        if ref.shape[1] != n or ref.shape[0] not in (1, R):
            raise ValueError("reference: expected [rows, channels] or one shared row")
        ref_stride = 0 if (ref.shape[0] == 1 and R > 1) else ref.row_stride
        if np.isscalar(scaling):
            return None, 0, float(scaling), C.c_void_p(ref.ptr), ref_stride, ref, None
        sc = DeviceArray(scaling, ("<f4",), "scaling")
        if sc.shape[1] != n or sc.shape[0] not in (1, R):
            raise ValueError("scaling: expected a float, [rows, channels] or one shared row")
        sc_stride = 0 if (sc.shape[0] == 1 and R > 1) else sc.row_stride
        return C.c_void_p(sc.ptr), sc_stride, 0.0, C.c_void_p(ref.ptr), ref_stride, ref, sc

    def calibrate(self, scaling, data, reference, out=None, stream=None):
        """result = scaling * (data - reference) / reference on device arrays
        (sakura_CalibrateDataFloatBatchDevice); `scaling` is a float or an array, `reference` and
        an array `scaling` may be one row shared by all rows."""
        d = DeviceArray(data, ("<f4",), "data")
        R, n = d.shape
        sc, scs, const, ref, refs, _k1, _k2 = self._calibration_args(scaling, reference, R, n)
        if out is None:
            out = _torch().empty((R, n), dtype=_torch().float32, device=self._device_of(data))
        o = DeviceArray(out, ("<f4",), "out")
        if o.shape != d.shape:
            raise ValueError("out: shape differs from data")
        st = self.lib.lib.sakura_CalibrateDataFloatBatchDevice(
            R, n, sc, scs, const, C.c_void_p(d.ptr), d.row_stride, ref, refs, C.c_void_p(o.ptr),
            o.row_stride, self._stream(stream))
        if st != STATUS_OK:
            raise RuntimeError(f"sakura_CalibrateDataFloatBatchDevice -> {st}: {self.lib.last_error()}")
        return out

    def calibrate_and_polynomial(self, ctx, order: int, scaling, data, reference, mask,
                                 clip_threshold_sigma: float, num_fitting_max: int,
                                 flag_nan_or_inf: bool = False, num_coeff: Optional[int] = None,
                                 want_coeff: bool = True, want_best_fit: bool = False,
                                 want_residual: bool = True, out: Optional[dict] = None,
                                 stream=None) -> dict:
        """sakura_CalibrateAndLSQFitPolynomialFloatBatchDevice: the fit of the calibrated spectra,
        calibrated (and NaN/Inf-flagged) inside the fit kernel's load."""
        d = DeviceArray(data, ("<f4",), "data")
        m = DeviceArray(mask, ("|b1", "|u1"), "mask")
        if d.shape != m.shape:
            raise ValueError("data and mask shapes differ")
        R, n = d.shape
        K = order + 1 if num_coeff is None else num_coeff
        sc, scs, const, ref, refs, _k1, _k2 = self._calibration_args(scaling, reference, R, n)
        res = self._outputs(out, self._device_of(data), R, n, (K,), want_coeff, want_best_fit,
                            want_residual)
        bf, rs, fm = self._row_outputs(res, R, n, d.row_stride, m.row_stride)
        co = self._flat(res["coeff"], ("<f8",), "coeff", R * K) if res["coeff"] is not None else None
        res["call_status"] = self.lib.lib.sakura_CalibrateAndLSQFitPolynomialFloatBatchDevice(
            ctx, order, R, n, sc, scs, const, C.c_void_p(d.ptr), d.row_stride, ref, refs,
            bool(flag_nan_or_inf), C.c_void_p(m.ptr), m.row_stride, clip_threshold_sigma,
            num_fitting_max, K, co, bf, rs, fm, self._flat(res["rms"], ("<f4",), "rms", R),
            self._flat(res["status"], ("<i4",), "status", R),
            self._flat(res["lsqfit_status"], ("<i4",), "lsqfit_status", R), self._stream(stream))
        return res

    def cubicspline(self, ctx, num_pieces: int, data, mask, clip_threshold_sigma: float,
                    num_fitting_max: int, want_coeff: bool = True, want_best_fit: bool = False,
                    want_residual: bool = True, out: Optional[dict] = None, stream=None) -> dict:
        """sakura_LSQFitCubicSplineFloatBatchDevice; coeff is [R, num_pieces, 4], boundary [R, np+1]."""
        d = DeviceArray(data, ("<f4",), "data")
        m = DeviceArray(mask, ("|b1", "|u1"), "mask")
        if d.shape != m.shape:
            raise ValueError("data and mask shapes differ")
        R, n = d.shape
        res = self._outputs(out, self._device_of(data), R, n, (num_pieces, 4), want_coeff,
                            want_best_fit, want_residual, boundary_cols=num_pieces + 1)
        bf, rs, fm = self._row_outputs(res, R, n, d.row_stride, m.row_stride)
        co = None
        if res["coeff"] is not None:
            co = C.c_void_p(int(res["coeff"].__cuda_array_interface__["data"][0]))
        bd = C.c_void_p(int(res["boundary"].__cuda_array_interface__["data"][0]))
        res["call_status"] = self.lib.lib.sakura_LSQFitCubicSplineFloatBatchDevice(
            ctx, num_pieces, R, n, C.c_void_p(d.ptr), d.row_stride, C.c_void_p(m.ptr), m.row_stride,
            clip_threshold_sigma, num_fitting_max, co, bf, rs, fm,
            self._flat(res["rms"], ("<f4",), "rms", R), bd,
            self._flat(res["status"], ("<i4",), "status", R),
            self._flat(res["lsqfit_status"], ("<i4",), "lsqfit_status", R), self._stream(stream))
        return res

    def sinusoid(self, ctx, nwave: Sequence[int], data, mask, clip_threshold_sigma: float,
                 num_fitting_max: int, num_coeff: Optional[int] = None, want_coeff: bool = True,
                 want_best_fit: bool = False, want_residual: bool = True, out: Optional[dict] = None,
                 stream=None) -> dict:
        """sakura_LSQFitSinusoidFloatBatchDevice; `nwave` is a host sequence (as in the C ABI)."""
        d = DeviceArray(data, ("<f4",), "data")
        m = DeviceArray(mask, ("|b1", "|u1"), "mask")
        if d.shape != m.shape:
            raise ValueError("data and mask shapes differ")
        R, n = d.shape
        nw = aligned_copy(np.asarray(nwave, dtype=np.uint64))
        if num_coeff is None:
            num_coeff = 2 * len(nw) - 1 if nw[0] == 0 else 2 * len(nw)
        res = self._outputs(out, self._device_of(data), R, n, (num_coeff,), want_coeff, want_best_fit,
                            want_residual)
        bf, rs, fm = self._row_outputs(res, R, n, d.row_stride, m.row_stride)
        co = self._flat(res["coeff"], ("<f8",), "coeff", R * num_coeff) if res["coeff"] is not None else None
        res["call_status"] = self.lib.lib.sakura_LSQFitSinusoidFloatBatchDevice(
            ctx, len(nw), nw.ctypes.data_as(C.c_void_p), R, n, C.c_void_p(d.ptr), d.row_stride,
            C.c_void_p(m.ptr), m.row_stride, clip_threshold_sigma, num_fitting_max, num_coeff, co, bf,
            rs, fm, self._flat(res["rms"], ("<f4",), "rms", R),
            self._flat(res["status"], ("<i4",), "status", R),
            self._flat(res["lsqfit_status"], ("<i4",), "lsqfit_status", R), self._stream(stream))
        return res

    # ---- mask builders: flat device arrays, optional AND into an existing mask ----
    def _filter_out(self, data, out):
        if out is None:
            out = _torch().empty(tuple(data.shape), dtype=_torch().bool, device=self._device_of(data))
        return out

    @staticmethod
    def _addr(obj):
        return None if obj is None else C.c_void_p(obj.data_ptr() if hasattr(obj, "data_ptr") else
                                                    obj.__cuda_array_interface__["data"][0])

    def _check(self, st, what):
        if st != STATUS_OK:
            raise RuntimeError(f"{what} -> {st}: {self.lib.last_error()}")

    def in_ranges(self, data, lower, upper, inclusive: bool = False, and_with=None, out=None, stream=None):
        """sakura_SetTrueIfInRanges{Float,Int}Device on a contiguous float32 / int32 device array of any
        shape; `lower` / `upper` are host sequences; `and_with` (may be `out`) is ANDed in."""
        is_int = "int" in str(data.dtype)
        dt = np.int32 if is_int else np.float32
        lo = np.ascontiguousarray(np.asarray(lower, dtype=dt).reshape(-1))
        hi = np.ascontiguousarray(np.asarray(upper, dtype=dt).reshape(-1))
        out = self._filter_out(data, out)
        fn = self.lib.lib.sakura_SetTrueIfInRangesIntDevice if is_int else self.lib.lib.sakura_SetTrueIfInRangesFloatDevice
        self._check(fn(bool(inclusive), int(np.prod(data.shape)), self._addr(data), lo.size,
                       C.c_void_p(lo.ctypes.data), C.c_void_p(hi.ctypes.data), self._addr(and_with),
                       self._addr(out), self._stream(stream)), "sakura_SetTrueIfInRanges*Device")
        return out

    def finite(self, data, and_with=None, out=None, stream=None):
        """sakura_SetFalseIfNanOrInfFloatDevice."""
        out = self._filter_out(data, out)
        self._check(self.lib.lib.sakura_SetFalseIfNanOrInfFloatDevice(
            int(np.prod(data.shape)), self._addr(data), self._addr(and_with), self._addr(out),
            self._stream(stream)), "sakura_SetFalseIfNanOrInfFloatDevice")
        return out

    def uint8_to_bool(self, data, invert: bool = False, and_with=None, out=None, stream=None):
        """sakura_Uint8ToBoolDevice (and sakura_InvertBoolDevice on the result when `invert`)."""
        out = self._filter_out(data, out)
        n = int(np.prod(data.shape))
        self._check(self.lib.lib.sakura_Uint8ToBoolDevice(n, self._addr(data), None if invert else self._addr(and_with),
                                                        self._addr(out), self._stream(stream)), "sakura_Uint8ToBoolDevice")
        if invert:
            self._check(self.lib.lib.sakura_InvertBoolDevice(n, self._addr(out), self._addr(and_with), self._addr(out),
                                                           self._stream(stream)), "sakura_InvertBoolDevice")
        return out

    def interpolate_y(self, method: int, base_position, base_data, base_mask, interpolated_position,
                      out=None, out_mask=None, stream=None):
        """sakura_InterpolateYAxisFloatDevice: base_data [num_base, num_array] -> [num_interpolated,
        num_array] (float64 positions on the device). Returns (data, mask)."""
        torch = _torch()
        nb, na = base_data.shape
        ni = int(interpolated_position.shape[0])
        if out is None:
            out = torch.empty((ni, na), dtype=torch.float32, device=self._device_of(base_data))
        if out_mask is None:
            out_mask = torch.empty((ni, na), dtype=torch.bool, device=self._device_of(base_data))
        self._check(self.lib.lib.sakura_InterpolateYAxisFloatDevice(
            method, 0, nb, self._addr(base_position), na, self._addr(base_data), self._addr(base_mask), ni,
            self._addr(interpolated_position), self._addr(out), self._addr(out_mask), self._stream(stream)),
            "sakura_InterpolateYAxisFloatDevice")
        return out, out_mask

    def statistics(self, data, mask, stream=None):
        """sakura_ComputeStatisticsFloatBatchDevice; returns a torch uint8 tensor [R, 48] whose rows
        are sakura_StatisticsResultFloat records (view it with capi.STAT_DTYPE after .cpu())."""
        d = DeviceArray(data, ("<f4",), "data")
        m = DeviceArray(mask, ("|b1", "|u1"), "mask")
        if d.shape != m.shape:
            raise ValueError("data and mask shapes differ")
        R, n = d.shape
        torch = _torch()
        from .capi import STAT_DTYPE
        rec = torch.empty((R, STAT_DTYPE.itemsize), dtype=torch.uint8, device=self._device_of(data))
        st = self.lib.lib.sakura_ComputeStatisticsFloatBatchDevice(
            R, n, C.c_void_p(d.ptr), d.row_stride, C.c_void_p(m.ptr), m.row_stride,
            C.c_void_p(rec.data_ptr()), self._stream(stream))
        if st != STATUS_OK:
            raise RuntimeError(f"sakura_ComputeStatisticsFloatBatchDevice -> {st}: "
                               f"{self.lib.last_error()}")
        return rec
