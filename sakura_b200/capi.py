"""ctypes binding of the C ABI in include/sakura_b200.h.

Host-side mirror of the reference's python-binding for this path
(libsakura/python-binding/python-binding.cc:1610-1757 `create_baseline_context`,
`lsqfit_polynomial`, `compute_statistics`; aligned buffers :505-563), extended with the
batched entry points.  numpy arrays in, numpy arrays out; torch tensors are accepted by
the `*_device` methods through `data_ptr()`.

The binding never computes anything itself: every call goes through
libsakura_b200.so, and loading fails loudly when the library cannot be built/loaded.
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Optional, Sequence

import numpy as np

ALIGNMENT = 32

STATUS_OK, STATUS_NG, STATUS_INVALID_ARGUMENT, STATUS_NO_MEMORY, STATUS_UNKNOWN = 0, 1, 2, 3, 99
LSQFIT_OK, LSQFIT_NG, LSQFIT_NOT_ENOUGH_DATA = 0, 1, 2
LSQFIT_TYPE_POLYNOMIAL, LSQFIT_TYPE_CHEBYSHEV = 0, 1


class StatisticsResult(C.Structure):
    """sakura_StatisticsResultFloat (sakura.h:208-237)."""
    _fields_ = [("count", C.c_size_t), ("sum", C.c_double), ("square_sum", C.c_double),
                ("min", C.c_float), ("max", C.c_float),
                ("index_of_min", C.c_ssize_t), ("index_of_max", C.c_ssize_t)]


STAT_DTYPE = np.dtype([("count", np.uint64), ("sum", np.float64), ("square_sum", np.float64),
                       ("min", np.float32), ("max", np.float32),
                       ("index_of_min", np.int64), ("index_of_max", np.int64)], align=True)
assert STAT_DTYPE.itemsize == C.sizeof(StatisticsResult) == 48


def aligned_empty(shape, dtype, align: int = 64) -> np.ndarray:
    """Uninitialised C-contiguous array whose base address is `align`-byte aligned
    (the reference's new_uninitialized_aligned_ndarray, python-binding.cc:505-563)."""
    dtype = np.dtype(dtype)
    shape = (shape,) if np.isscalar(shape) else tuple(shape)
    nbytes = int(np.prod(shape, dtype=np.int64)) * dtype.itemsize
    raw = np.empty(nbytes + align, dtype=np.uint8)
    off = (-raw.ctypes.data) % align
    return raw[off:off + nbytes].view(dtype).reshape(shape)


def aligned_copy(a, dtype=None, align: int = 64) -> np.ndarray:
    a = np.asarray(a)
    out = aligned_empty(a.shape, dtype or a.dtype, align)
    out[...] = a
    return out


def _is_aligned(a: np.ndarray) -> bool:
    return a.ctypes.data % ALIGNMENT == 0


def _ptr(a: Optional[np.ndarray]):
    return None if a is None else C.c_void_p(a.ctypes.data)


def default_library_path() -> str:
    # SAKURA_B200_LIB: another build of the same library (A/B timing of two builds in one session)
    override = os.environ.get("SAKURA_B200_LIB")
    if override:
        return override
    return os.path.join(os.path.dirname(os.path.abspath(__file__)), "libsakura_b200.so")


class SakuraLib:
    """A loaded library that exports the sakura.h LSQ-fit / statistics C API."""

    def __init__(self, path: Optional[str] = None, has_batch: bool = True):
        if path is None:
            path = default_library_path()
            if not os.path.exists(path):
                from . import build as _build
                _build.build()
        if not os.path.exists(path):
            raise OSError(f"{path} is missing: libsakura_b200.so must be built "
                          "(python -m sakura_b200.build); there is no CPU fallback")
        self.path = path
        self.lib = C.CDLL(path)
        self.has_batch = has_batch
        self._declare()
        st = self.lib.sakura_Initialize(None, None)
        if st != STATUS_OK:
            raise RuntimeError("sakura_Initialize failed")

    # ------------------------------------------------------------------ prototypes
    def _declare(self) -> None:
        L = self.lib
        vp, sz, u16, f32 = C.c_void_p, C.c_size_t, C.c_uint16, C.c_float
        L.sakura_Initialize.argtypes = [vp, vp]
        L.sakura_Initialize.restype = C.c_int
        L.sakura_GetAlignment.restype = sz
        L.sakura_IsAligned.argtypes = [vp]
        L.sakura_IsAligned.restype = C.c_bool
        L.sakura_CreateLSQFitContextPolynomialFloat.argtypes = [C.c_int, u16, sz, C.POINTER(vp)]
        L.sakura_CreateLSQFitContextCubicSplineFloat.argtypes = [u16, sz, C.POINTER(vp)]
        L.sakura_CreateLSQFitContextSinusoidFloat.argtypes = [u16, sz, C.POINTER(vp)]
        L.sakura_DestroyLSQFitContextFloat.argtypes = [vp]
        L.sakura_GetNumberOfCoefficientsFloat.argtypes = [vp, u16, C.POINTER(sz)]
        L.sakura_LSQFitPolynomialFloat.argtypes = [vp, u16, sz, vp, vp, f32, u16, sz, vp, vp, vp, vp,
                                                   C.POINTER(f32), C.POINTER(C.c_int)]
        L.sakura_LSQFitCubicSplineFloat.argtypes = [vp, sz, sz, vp, vp, f32, u16, vp, vp, vp, vp,
                                                    C.POINTER(f32), vp, C.POINTER(C.c_int)]
        L.sakura_LSQFitSinusoidFloat.argtypes = [vp, sz, vp, sz, vp, vp, f32, u16, sz, vp, vp, vp, vp,
                                                 C.POINTER(f32), C.POINTER(C.c_int)]
        L.sakura_SubtractPolynomialFloat.argtypes = [vp, sz, vp, sz, vp, vp]
        L.sakura_SubtractCubicSplineFloat.argtypes = [vp, sz, vp, sz, vp, vp, vp]
        L.sakura_SubtractSinusoidFloat.argtypes = [vp, sz, vp, sz, vp, sz, vp, vp]
        L.sakura_ComputeStatisticsFloat.argtypes = [sz, vp, vp, vp]
        L.sakura_ComputeAccurateStatisticsFloat.argtypes = [sz, vp, vp, vp]
        L.sakura_GetLSQCoefficientsDouble.argtypes = [sz, vp, vp, sz, vp, sz, vp, vp, vp]
        L.sakura_UpdateLSQCoefficientsDouble.argtypes = [sz, vp, vp, sz, vp, sz, vp, sz, vp, vp, vp]
        L.sakura_SolveSimultaneousEquationsByLUDouble.argtypes = [sz, vp, vp, vp]
        for name in ("sakura_CreateLSQFitContextPolynomialFloat",
                     "sakura_CreateLSQFitContextCubicSplineFloat",
                     "sakura_CreateLSQFitContextSinusoidFloat", "sakura_DestroyLSQFitContextFloat",
                     "sakura_GetNumberOfCoefficientsFloat", "sakura_LSQFitPolynomialFloat",
                     "sakura_LSQFitCubicSplineFloat", "sakura_LSQFitSinusoidFloat",
                     "sakura_SubtractPolynomialFloat", "sakura_SubtractCubicSplineFloat",
                     "sakura_SubtractSinusoidFloat", "sakura_ComputeStatisticsFloat",
                     "sakura_ComputeAccurateStatisticsFloat", "sakura_GetLSQCoefficientsDouble",
                     "sakura_UpdateLSQCoefficientsDouble",
                     "sakura_SolveSimultaneousEquationsByLUDouble"):
            getattr(L, name).restype = C.c_int
        if not self.has_batch:
            return
        i32p = vp
        L.sakura_LSQFitPolynomialFloatBatch.argtypes = [vp, u16, sz, sz, vp, sz, vp, sz, f32, u16, sz,
                                                        vp, vp, vp, vp, vp, i32p, i32p]
        L.sakura_LSQFitPolynomialFloatBatchDevice.argtypes = \
            L.sakura_LSQFitPolynomialFloatBatch.argtypes + [vp]
        L.sakura_LSQFitCubicSplineFloatBatch.argtypes = [vp, sz, sz, sz, vp, sz, vp, sz, f32, u16,
                                                         vp, vp, vp, vp, vp, vp, i32p, i32p]
        L.sakura_LSQFitCubicSplineFloatBatchDevice.argtypes = \
            L.sakura_LSQFitCubicSplineFloatBatch.argtypes + [vp]
        L.sakura_LSQFitSinusoidFloatBatch.argtypes = [vp, sz, vp, sz, sz, vp, sz, vp, sz, f32, u16, sz,
                                                      vp, vp, vp, vp, vp, i32p, i32p]
        L.sakura_LSQFitSinusoidFloatBatchDevice.argtypes = \
            L.sakura_LSQFitSinusoidFloatBatch.argtypes + [vp]
        for kind in ("Polynomial", "CubicSpline", "Sinusoid"):
            fn = getattr(L, f"sakura_LSQFit{kind}FloatBatchMultiDevice")
            fn.argtypes = getattr(L, f"sakura_LSQFit{kind}FloatBatch").argtypes + [sz, vp]
            fn.restype = C.c_int
        L.sakura_ComputeStatisticsFloatBatch.argtypes = [sz, sz, vp, sz, vp, sz, vp]
        L.sakura_ComputeStatisticsFloatBatchDevice.argtypes = [sz, sz, vp, sz, vp, sz, vp, vp]
        L.sakura_CalibrateDataWithArrayScalingFloat.argtypes = [sz, vp, vp, vp, vp]
        L.sakura_CalibrateDataWithConstScalingFloat.argtypes = [f32, sz, vp, vp, vp]
        L.sakura_CalibrateDataFloatBatchDevice.argtypes = [sz, sz, vp, sz, f32, vp, sz, vp, sz, vp, sz, vp]
        L.sakura_CalibrateAndLSQFitPolynomialFloatBatchDevice.argtypes = \
            [vp, u16, sz, sz, vp, sz, f32, vp, sz, vp, sz, C.c_bool, vp, sz, f32, u16, sz,
             vp, vp, vp, vp, vp, i32p, i32p, vp]
        for name in ("sakura_CalibrateDataWithArrayScalingFloat",
                     "sakura_CalibrateDataWithConstScalingFloat",
                     "sakura_CalibrateDataFloatBatchDevice",
                     "sakura_CalibrateAndLSQFitPolynomialFloatBatchDevice"):
            getattr(L, name).restype = C.c_int
        for name in ("sakura_LSQFitPolynomialFloatBatch", "sakura_LSQFitPolynomialFloatBatchDevice",
                     "sakura_LSQFitCubicSplineFloatBatch", "sakura_LSQFitCubicSplineFloatBatchDevice",
                     "sakura_LSQFitSinusoidFloatBatch", "sakura_LSQFitSinusoidFloatBatchDevice",
                     "sakura_ComputeStatisticsFloatBatch", "sakura_ComputeStatisticsFloatBatchDevice"):
            getattr(L, name).restype = C.c_int
        i32c = C.c_int
        for suffix, t in (("Float", f32), ("Int", i32c)):
            for incl in ("Inclusive", "Exclusive"):
                fn = getattr(L, f"sakura_SetTrueIfInRanges{incl}{suffix}")
                fn.argtypes = [sz, vp, sz, vp, vp, vp]
                fn.restype = C.c_int
            for cmp_ in ("GreaterThan", "GreaterThanOrEquals", "LessThan", "LessThanOrEquals"):
                fn = getattr(L, f"sakura_SetTrueIf{cmp_}{suffix}")
                fn.argtypes = [sz, vp, t, vp]
                fn.restype = C.c_int
            getattr(L, f"sakura_SetTrueIfInRanges{suffix}Device").argtypes = [C.c_bool, sz, vp, sz, vp, vp,
                                                                             vp, vp, vp]
            getattr(L, f"sakura_SetTrueIfCompares{suffix}Device").argtypes = [C.c_int, sz, vp, t, vp, vp, vp]
        for name in ("sakura_SetFalseIfNanOrInfFloat", "sakura_Uint8ToBool", "sakura_Uint32ToBool",
                     "sakura_InvertBool"):
            getattr(L, name).argtypes = [sz, vp, vp]
            getattr(L, name).restype = C.c_int
            getattr(L, name + "Device").argtypes = [sz, vp, vp, vp, vp]
            getattr(L, name + "Device").restype = C.c_int
        for name in ("sakura_SetTrueIfInRangesFloatDevice", "sakura_SetTrueIfInRangesIntDevice",
                     "sakura_SetTrueIfComparesFloatDevice", "sakura_SetTrueIfComparesIntDevice"):
            getattr(L, name).restype = C.c_int
        for op in ("And", "ConverseNonImplication", "Implication", "Or", "Xor"):
            getattr(L, f"sakura_OperateBitwise{op}Uint8").argtypes = [C.c_uint8, sz, vp, vp, vp]
            getattr(L, f"sakura_OperateBitwise{op}Uint32").argtypes = [C.c_uint32, sz, vp, vp, vp]
        L.sakura_OperateBitwiseNotUint8.argtypes = [sz, vp, vp, vp]
        L.sakura_OperateBitwiseNotUint32.argtypes = [sz, vp, vp, vp]
        L.sakura_OperateBitwiseUint8Device.argtypes = [C.c_int, C.c_uint8, sz, vp, vp, vp, vp]
        L.sakura_OperateBitwiseUint32Device.argtypes = [C.c_int, C.c_uint32, sz, vp, vp, vp, vp]
        for name in ("sakura_InterpolateXAxisFloat", "sakura_InterpolateYAxisFloat"):
            getattr(L, name).argtypes = [C.c_int, C.c_uint8, sz, vp, sz, vp, vp, sz, vp, vp, vp]
            getattr(L, name).restype = C.c_int
            getattr(L, name + "Device").argtypes = [C.c_int, C.c_uint8, sz, vp, sz, vp, vp, sz, vp, vp, vp, vp]
            getattr(L, name + "Device").restype = C.c_int
        L.sakura_b200_GetKernelLaunchCount.restype = C.c_uint64
        L.sakura_b200_GetLastHandOverCount.argtypes = [vp]
        L.sakura_b200_GetLastHandOverCount.restype = C.c_int64
        L.sakura_b200_GetLastKernelName.restype = C.c_char_p
        L.sakura_b200_GetLastErrorMessage.restype = C.c_char_p
        L.sakura_b200_GetVersionString.restype = C.c_char_p
        L.sakura_b200_SetHostPipelineTuning.argtypes = [C.c_int64, C.c_int64]
        L.sakura_b200_SetHostPipelineTuning.restype = None

    # ------------------------------------------------------------------ contexts
    def create_polynomial_context(self, order: int, num_data: int,
                                  lsqfit_type: int = LSQFIT_TYPE_POLYNOMIAL):
        ctx = C.c_void_p()
        st = self.lib.sakura_CreateLSQFitContextPolynomialFloat(lsqfit_type, order, num_data,
                                                                C.byref(ctx))
        return st, (ctx if st == STATUS_OK else None)

    def create_cubicspline_context(self, npiece: int, num_data: int):
        ctx = C.c_void_p()
        st = self.lib.sakura_CreateLSQFitContextCubicSplineFloat(npiece, num_data, C.byref(ctx))
        return st, (ctx if st == STATUS_OK else None)

    def create_sinusoid_context(self, nwave: int, num_data: int):
        ctx = C.c_void_p()
        st = self.lib.sakura_CreateLSQFitContextSinusoidFloat(nwave, num_data, C.byref(ctx))
        return st, (ctx if st == STATUS_OK else None)

    def destroy_context(self, ctx) -> int:
        return self.lib.sakura_DestroyLSQFitContextFloat(ctx)

    def get_number_of_coefficients(self, ctx, order: int):
        n = C.c_size_t(0)
        st = self.lib.sakura_GetNumberOfCoefficientsFloat(ctx, order, C.byref(n))
        return st, n.value

    # ------------------------------------------------------------------ launch accounting
    def kernel_launch_count(self) -> int:
        return int(self.lib.sakura_b200_GetKernelLaunchCount())

    def last_kernel_name(self) -> str:
        return self.lib.sakura_b200_GetLastKernelName().decode()

    def last_handover_count(self, ctx) -> int:
        """Rows the screening kernel handed to the per-channel kernel in the last fit on `ctx`."""
        return int(self.lib.sakura_b200_GetLastHandOverCount(ctx))

    def set_host_pipeline_tuning(self, chunk_mib: int = -1, packed_kib: int = -1) -> None:
        """sakura_b200_SetHostPipelineTuning; -1 keeps a setting."""
        self.lib.sakura_b200_SetHostPipelineTuning(chunk_mib, packed_kib)

    def last_error(self) -> str:
        return self.lib.sakura_b200_GetLastErrorMessage().decode()

    # ------------------------------------------------------------------ batched fits (host arrays)
    def _batch_call(self, kind: str, devices, args):
        """`devices` None: sakura_LSQFit<kind>FloatBatch on the current device; "all" or a sequence of
        CUDA device ordinals: sakura_LSQFit<kind>FloatBatchMultiDevice (rows split over the devices)."""
        if devices is None:
            return getattr(self.lib, f"sakura_LSQFit{kind}FloatBatch")(*args)
        fn = getattr(self.lib, f"sakura_LSQFit{kind}FloatBatchMultiDevice")
        if isinstance(devices, str):
            if devices != "all":
                raise ValueError("devices must be None, 'all' or a sequence of device ordinals")
            return fn(*args, 0, None)
        ids = (C.c_int * len(devices))(*[int(d) for d in devices])
        return fn(*args, len(devices), C.cast(ids, C.c_void_p))

    @staticmethod
    def _prep_rows(data, mask):
        data = np.asarray(data)
        mask = np.asarray(mask)
        if data.ndim == 1:
            data = data[None, :]
            mask = mask[None, :]
        if data.dtype != np.float32 or not data.flags.c_contiguous or not _is_aligned(data):
            data = aligned_copy(data, np.float32)
        if mask.dtype != np.bool_ or not mask.flags.c_contiguous or not _is_aligned(mask):
            mask = aligned_copy(mask, np.bool_)
        return data, mask

    def lsqfit_polynomial_batch(self, ctx, order: int, data, mask, clip_threshold_sigma: float,
                                num_fitting_max: int, num_coeff: Optional[int] = None,
                                want_coeff: bool = True, want_best_fit: bool = False,
                                want_residual: bool = True, out: Optional[dict] = None,
                                devices=None) -> dict:
        """Rows of `data`/`mask` ([R, n]) through sakura_LSQFitPolynomialFloatBatch[MultiDevice]."""
        data, mask = self._prep_rows(data, mask)
        R, n = data.shape
        K = order + 1 if num_coeff is None else num_coeff
        res = out if out is not None else {}
        res.setdefault("coeff", aligned_empty((R, K), np.float64) if want_coeff else None)
        res.setdefault("best_fit", aligned_empty((R, n), np.float32) if want_best_fit else None)
        res.setdefault("residual", aligned_empty((R, n), np.float32) if want_residual else None)
        res.setdefault("final_mask", aligned_empty((R, n), np.bool_))
        res.setdefault("rms", aligned_empty((R,), np.float32))
        res["status"] = aligned_empty((R,), np.int32)
        res["lsqfit_status"] = aligned_empty((R,), np.int32)
        if out is None:
            for k in ("coeff", "best_fit", "residual", "rms"):
                if res[k] is not None:
                    res[k][...] = np.nan
            res["final_mask"][...] = False
        res["call_status"] = self._batch_call("Polynomial", devices, (
            ctx, order, R, n, _ptr(data), n, _ptr(mask), n, clip_threshold_sigma, num_fitting_max, K,
            _ptr(res["coeff"]), _ptr(res["best_fit"]), _ptr(res["residual"]), _ptr(res["final_mask"]),
            _ptr(res["rms"]), _ptr(res["status"]), _ptr(res["lsqfit_status"])))
        return res

    def lsqfit_cubicspline_batch(self, ctx, num_pieces: int, data, mask, clip_threshold_sigma: float,
                                 num_fitting_max: int, want_coeff: bool = True,
                                 want_best_fit: bool = False, want_residual: bool = True,
                                 devices=None) -> dict:
        data, mask = self._prep_rows(data, mask)
        R, n = data.shape
        res = {
            "coeff": aligned_empty((R, num_pieces, 4), np.float64) if want_coeff else None,
            "best_fit": aligned_empty((R, n), np.float32) if want_best_fit else None,
            "residual": aligned_empty((R, n), np.float32) if want_residual else None,
            "final_mask": aligned_empty((R, n), np.bool_),
            "rms": aligned_empty((R,), np.float32),
            "boundary": aligned_empty((R, num_pieces + 1), np.uint64),
            "status": aligned_empty((R,), np.int32),
            "lsqfit_status": aligned_empty((R,), np.int32),
        }
        for k in ("coeff", "best_fit", "residual", "rms"):
            if res[k] is not None:
                res[k][...] = np.nan
        res["final_mask"][...] = False
        res["boundary"][...] = np.iinfo(np.uint64).max
        res["call_status"] = self._batch_call("CubicSpline", devices, (
            ctx, num_pieces, R, n, _ptr(data), n, _ptr(mask), n, clip_threshold_sigma, num_fitting_max,
            _ptr(res["coeff"]), _ptr(res["best_fit"]), _ptr(res["residual"]), _ptr(res["final_mask"]),
            _ptr(res["rms"]), _ptr(res["boundary"]), _ptr(res["status"]), _ptr(res["lsqfit_status"])))
        return res

    def lsqfit_sinusoid_batch(self, ctx, nwave: Sequence[int], data, mask, clip_threshold_sigma: float,
                              num_fitting_max: int, num_coeff: Optional[int] = None,
                              want_coeff: bool = True, want_best_fit: bool = False,
                              want_residual: bool = True, devices=None) -> dict:
        data, mask = self._prep_rows(data, mask)
        R, n = data.shape
        nw = aligned_copy(np.asarray(nwave, dtype=np.uint64))
        if num_coeff is None:
            num_coeff = 2 * len(nw) - 1 if nw[0] == 0 else 2 * len(nw)
        res = {
            "coeff": aligned_empty((R, num_coeff), np.float64) if want_coeff else None,
            "best_fit": aligned_empty((R, n), np.float32) if want_best_fit else None,
            "residual": aligned_empty((R, n), np.float32) if want_residual else None,
            "final_mask": aligned_empty((R, n), np.bool_),
            "rms": aligned_empty((R,), np.float32),
            "status": aligned_empty((R,), np.int32),
            "lsqfit_status": aligned_empty((R,), np.int32),
        }
        for k in ("coeff", "best_fit", "residual", "rms"):
            if res[k] is not None:
                res[k][...] = np.nan
        res["final_mask"][...] = False
        res["call_status"] = self._batch_call("Sinusoid", devices, (
            ctx, len(nw), _ptr(nw), R, n, _ptr(data), n, _ptr(mask), n, clip_threshold_sigma,
            num_fitting_max, num_coeff, _ptr(res["coeff"]), _ptr(res["best_fit"]),
            _ptr(res["residual"]), _ptr(res["final_mask"]), _ptr(res["rms"]), _ptr(res["status"]),
            _ptr(res["lsqfit_status"])))
        return res

    # ------------------------------------------------------------------ calibration
    def calibrate(self, scaling, data, reference, in_place: bool = False):
        """One array through sakura_CalibrateDataWith{Array,Const}ScalingFloat (host pointers);
        `scaling` is a float or an array like data. Returns (status, result)."""
        data = aligned_copy(np.asarray(data, dtype=np.float32).ravel())
        ref = aligned_copy(np.asarray(reference, dtype=np.float32).ravel())
        n = data.size
        result = data if in_place else aligned_empty(max(n, 1), np.float32)[:n]
        if np.isscalar(scaling):
            st = self.lib.sakura_CalibrateDataWithConstScalingFloat(float(scaling), n, _ptr(data),
                                                                    _ptr(ref), _ptr(result))
        else:
            sc = aligned_copy(np.asarray(scaling, dtype=np.float32).ravel())
            st = self.lib.sakura_CalibrateDataWithArrayScalingFloat(n, _ptr(sc), _ptr(data),
                                                                    _ptr(ref), _ptr(result))
        return st, result

    # ------------------------------------------------------------------ boolean filters
    _RANGE_OPS = {"ranges_inclusive": "Inclusive", "ranges_exclusive": "Exclusive"}
    _COMPARE_OPS = {"gt": "GreaterThan", "ge": "GreaterThanOrEquals", "lt": "LessThan",
                    "le": "LessThanOrEquals"}

    def bool_filter(self, op: str, data, lower=None, upper=None, threshold=0):
        """Host arrays through the reference-named entry point selected by `op` ("ranges_inclusive",
        "ranges_exclusive", "gt", "ge", "lt", "le", "finite", "nonzero", "invert") and the dtype of
        `data`. Returns (result bool[n], status)."""
        data = np.asarray(data)
        if not _is_aligned(data) or not data.flags.c_contiguous:
            data = aligned_copy(data)
        n = data.size
        if n == 0:
            data = aligned_empty(8, data.dtype)
        suffix = {np.dtype(np.float32): "Float", np.dtype(np.int32): "Int"}.get(data.dtype)
        result = aligned_empty((max(n, 1),), np.bool_)
        result[...] = False
        if op in self._RANGE_OPS:
            lo = aligned_copy(np.asarray(lower, dtype=data.dtype).reshape(-1))
            hi = aligned_copy(np.asarray(upper, dtype=data.dtype).reshape(-1))
            nc = lo.size
            if nc == 0:
                lo, hi = aligned_empty(8, data.dtype), aligned_empty(8, data.dtype)
            fn = getattr(self.lib, f"sakura_SetTrueIfInRanges{self._RANGE_OPS[op]}{suffix}")
            st = fn(n, _ptr(data), nc, _ptr(lo), _ptr(hi), _ptr(result))
        elif op in self._COMPARE_OPS:
            fn = getattr(self.lib, f"sakura_SetTrueIf{self._COMPARE_OPS[op]}{suffix}")
            st = fn(n, _ptr(data), threshold, _ptr(result))
        elif op == "finite":
            st = self.lib.sakura_SetFalseIfNanOrInfFloat(n, _ptr(data), _ptr(result))
        elif op == "nonzero":
            fn = self.lib.sakura_Uint32ToBool if data.dtype == np.uint32 else self.lib.sakura_Uint8ToBool
            st = fn(n, _ptr(data), _ptr(result))
        elif op == "invert":
            st = self.lib.sakura_InvertBool(n, _ptr(data), _ptr(result))
        else:
            raise ValueError(op)
        return result[:n], int(st)

    # ------------------------------------------------------------------ bit operations
    _BIT_NAMES = {"and": "And", "converse_nonimplication": "ConverseNonImplication",
                  "implication": "Implication", "not": "Not", "or": "Or", "xor": "Xor"}

    def bit_operation(self, op: str, bit_mask: int, data, edit_mask, in_place: bool = False):
        """sakura_OperateBitwise<Op>{Uint8,Uint32} on host arrays; returns (result, status)."""
        data = np.asarray(data)
        suffix = "Uint32" if data.dtype == np.uint32 else "Uint8"
        n = data.size
        d = aligned_copy(data) if n else aligned_empty(8, data.dtype)
        m = aligned_copy(np.asarray(edit_mask).view(np.uint8)) if n else aligned_empty(8, np.uint8)
        out = d if in_place else aligned_empty((max(n, 1),), data.dtype)
        fn = getattr(self.lib, f"sakura_OperateBitwise{self._BIT_NAMES[op]}{suffix}")
        if op == "not":
            st = fn(n, _ptr(d), _ptr(m), _ptr(out))
        else:
            st = fn(int(bit_mask), n, _ptr(d), _ptr(m), _ptr(out))
        return out[:n], int(st)

    # ------------------------------------------------------------------ interpolation
    def interpolate(self, axis: str, method: int, base_position, base_data, base_mask,
                    interpolated_position, polynomial_order: int = 0, prefill: float = np.nan):
        """sakura_Interpolate{X,Y}AxisFloat on host arrays. axis "y": base_data [num_base, num_array]
        -> [num_interpolated, num_array]; axis "x": [num_array, num_base] -> [num_array,
        num_interpolated]. Returns (data, mask, status); arrays without a valid base point keep
        `prefill`."""
        bd = aligned_copy(np.asarray(base_data, dtype=np.float32))
        bm = aligned_copy(np.asarray(base_mask, dtype=np.bool_))
        bp = aligned_copy(np.asarray(base_position, dtype=np.float64))
        ip = aligned_copy(np.asarray(interpolated_position, dtype=np.float64))
        nb, ni = bp.size, ip.size
        na = bd.shape[1] if axis == "y" else bd.shape[0]
        shape = (ni, na) if axis == "y" else (na, ni)
        out = aligned_empty(shape, np.float32)
        out[...] = prefill
        om = aligned_empty(shape, np.bool_)
        om[...] = False
        fn = self.lib.sakura_InterpolateYAxisFloat if axis == "y" else self.lib.sakura_InterpolateXAxisFloat
        st = fn(method, polynomial_order, nb, _ptr(bp), na, _ptr(bd), _ptr(bm), ni, _ptr(ip), _ptr(out),
                _ptr(om))
        return out, om, int(st)

    # ------------------------------------------------------------------ statistics
    def compute_statistics_batch(self, data, mask) -> np.ndarray:
        data, mask = self._prep_rows(data, mask)
        R, n = data.shape
        result = aligned_empty((R,), STAT_DTYPE)
        st = self.lib.sakura_ComputeStatisticsFloatBatch(R, n, _ptr(data), n, _ptr(mask), n,
                                                         _ptr(result))
        if st != STATUS_OK:
            raise RuntimeError(f"sakura_ComputeStatisticsFloatBatch -> {st}: {self.last_error()}")
        return result

    def compute_statistics(self, data, mask, accurate: bool = True):
        """Single array through sakura_Compute[Accurate]StatisticsFloat; returns (status, record)."""
        data = np.asarray(data)
        mask = np.asarray(mask)
        if data.dtype != np.float32 or not _is_aligned(data) or not data.flags.c_contiguous:
            data = aligned_copy(data, np.float32)
        if mask.dtype != np.bool_ or not _is_aligned(mask) or not mask.flags.c_contiguous:
            mask = aligned_copy(mask, np.bool_)
        result = aligned_empty((1,), STAT_DTYPE)
        fn = (self.lib.sakura_ComputeAccurateStatisticsFloat if accurate
              else self.lib.sakura_ComputeStatisticsFloat)
        n = data.size
        if n == 0:  # numpy gives empty arrays arbitrary addresses; hand over aligned ones
            data = aligned_empty(8, np.float32)
            mask = aligned_empty(8, np.bool_)
        st = fn(n, _ptr(data), _ptr(mask), _ptr(result))
        return st, result[0]
