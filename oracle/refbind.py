"""TEST INFRASTRUCTURE ONLY (oracle). ctypes access to the CPU checkers:

* ``oracle/_ref/libsakura_ref.so`` -- the reference's own object code for the hot path
  (built by ``oracle/Makefile`` from /root/reference; kind "reference"), and
* ``oracle/libsakura_oracle.so``   -- the plain-C restatement (kind "port").

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs
may import this module.  The shipped package (sakura_b200/) never does.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from typing import Optional, Sequence

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REF_SO = os.path.join(HERE, "_ref", "libsakura_ref.so")
REF_NOFMA_SO = os.path.join(HERE, "_ref", "libsakura_ref_nofma.so")  # same sources, no fused multiply-adds
PORT_SO = os.path.join(HERE, "libsakura_oracle.so")

STAT_DTYPE = np.dtype([("count", np.uint64), ("sum", np.float64), ("square_sum", np.float64),
                       ("min", np.float32), ("max", np.float32),
                       ("index_of_min", np.int64), ("index_of_max", np.int64)], align=True)


def aligned_empty(shape, dtype, align: int = 64) -> np.ndarray:
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


def _ptr(a):
    return None if a is None else C.c_void_p(a.ctypes.data)


def build(verbose: bool = False) -> None:
    """Runs oracle/Makefile (reference build only where /root/reference exists)."""
    res = subprocess.run(["make", "-C", HERE, "all"], capture_output=True, text=True)
    if verbose or res.returncode != 0:
        print(res.stdout + res.stderr)
    if res.returncode != 0:
        raise RuntimeError("oracle build failed")


class Oracle:
    """Row-batched access to a CPU checker library exposing the ``ref_*`` driver API
    (oracle/ref_driver.cc for the reference, the same names in oracle/sakura_oracle.c)."""

    def __init__(self, kind: Optional[str] = None):
        if kind is None:
            kind = "reference" if os.path.exists(REF_SO) else "port"
        self.kind = kind
        path = {"reference": REF_SO, "reference_nofma": REF_NOFMA_SO}.get(kind, PORT_SO)
        if not os.path.exists(path):
            raise OSError(f"oracle library {path} is missing; run `make -C oracle`")
        self.path = path
        self.lib = C.CDLL(path)
        L = self.lib
        vp, sz, u16, f32, i32 = C.c_void_p, C.c_size_t, C.c_uint16, C.c_float, C.c_int
        L.ref_init.restype = i32
        L.ref_max_threads.restype = i32
        L.ref_lsqfit_polynomial_batch.argtypes = [i32, u16, u16, sz, sz, vp, sz, vp, sz, f32, u16, sz,
                                                  vp, vp, vp, vp, vp, vp, vp, i32]
        L.ref_lsqfit_cubicspline_batch.argtypes = [u16, sz, sz, sz, vp, sz, vp, sz, f32, u16, vp, vp,
                                                   vp, vp, vp, vp, vp, vp, i32]
        L.ref_lsqfit_sinusoid_batch.argtypes = [u16, sz, vp, sz, sz, vp, sz, vp, sz, f32, u16, sz, vp,
                                                vp, vp, vp, vp, vp, vp, i32]
        L.ref_statistics_batch.argtypes = [i32, sz, sz, vp, sz, vp, sz, vp, vp, i32]
        L.ref_calibrate_batch.argtypes = [sz, sz, vp, sz, f32, vp, sz, vp, sz, vp, sz, vp]
        L.ref_calibrate_batch.restype = i32
        L.ref_bool_filter.argtypes = [i32, i32, sz, vp, sz, vp, vp, C.c_uint32, vp]
        L.ref_bool_filter.restype = i32
        L.ref_bit_operation.argtypes = [i32, i32, C.c_uint32, sz, vp, vp, vp]
        L.ref_bit_operation.restype = i32
        L.ref_interpolate.argtypes = [i32, i32, i32, sz, vp, sz, vp, vp, sz, vp, vp, vp]
        L.ref_interpolate.restype = i32
        for name in ("ref_lsqfit_polynomial_batch", "ref_lsqfit_cubicspline_batch",
                     "ref_lsqfit_sinusoid_batch", "ref_statistics_batch"):
            getattr(L, name).restype = i32
        if L.ref_init() != 0:
            raise RuntimeError("oracle initialisation failed")

    def max_threads(self) -> int:
        return int(self.lib.ref_max_threads())

    @staticmethod
    def _rows(data, mask):
        data = np.ascontiguousarray(data, dtype=np.float32)
        mask = np.ascontiguousarray(mask, dtype=np.bool_)
        if data.ndim == 1:
            data, mask = data[None, :], mask[None, :]
        return data, mask

    @staticmethod
    def _outputs(R, n, coeff_shape, want_coeff, want_best_fit, want_residual):
        res = {
            "coeff": aligned_empty(coeff_shape, np.float64) if want_coeff else None,
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
        return res

    def lsqfit_polynomial_batch(self, order: int, data, mask, clip_threshold_sigma: float,
                                num_fitting_max: int, num_coeff: Optional[int] = None,
                                lsqfit_type: int = 0, context_order: Optional[int] = None,
                                want_coeff: bool = True, want_best_fit: bool = False,
                                want_residual: bool = True, nthreads: int = 0) -> dict:
        data, mask = self._rows(data, mask)
        R, n = data.shape
        K = order + 1 if num_coeff is None else num_coeff
        res = self._outputs(R, n, (R, K), want_coeff, want_best_fit, want_residual)
        rc = self.lib.ref_lsqfit_polynomial_batch(
            lsqfit_type, order if context_order is None else context_order, order, R, n,
            _ptr(data), n, _ptr(mask), n, clip_threshold_sigma, num_fitting_max, K,
            _ptr(res["coeff"]), _ptr(res["best_fit"]), _ptr(res["residual"]), _ptr(res["final_mask"]),
            _ptr(res["rms"]), _ptr(res["status"]), _ptr(res["lsqfit_status"]), nthreads)
        if rc != 0:
            raise RuntimeError("oracle: context creation failed")
        # the same fit of one row with another iteration count (tests/parity.py: threshold of the
        # iteration at which a clip decision diverged)
        res["refit"] = lambda row, nfit: self.lsqfit_polynomial_batch(
            order, data[row], mask[row], clip_threshold_sigma, nfit, num_coeff=num_coeff,
            lsqfit_type=lsqfit_type, context_order=context_order, want_coeff=False)
        res["num_fitting_max"] = num_fitting_max
        return res

    def lsqfit_cubicspline_batch(self, num_pieces: int, data, mask, clip_threshold_sigma: float,
                                 num_fitting_max: int, context_npiece: Optional[int] = None,
                                 want_coeff: bool = True, want_best_fit: bool = False,
                                 want_residual: bool = True, nthreads: int = 0) -> dict:
        data, mask = self._rows(data, mask)
        R, n = data.shape
        res = self._outputs(R, n, (R, num_pieces, 4), want_coeff, want_best_fit, want_residual)
        res["boundary"] = aligned_empty((R, num_pieces + 1), np.uint64)
        res["boundary"][...] = np.iinfo(np.uint64).max
        rc = self.lib.ref_lsqfit_cubicspline_batch(
            num_pieces if context_npiece is None else context_npiece, num_pieces, R, n, _ptr(data), n,
            _ptr(mask), n, clip_threshold_sigma, num_fitting_max, _ptr(res["coeff"]),
            _ptr(res["best_fit"]), _ptr(res["residual"]), _ptr(res["final_mask"]), _ptr(res["rms"]),
            _ptr(res["boundary"]), _ptr(res["status"]), _ptr(res["lsqfit_status"]), nthreads)
        if rc != 0:
            raise RuntimeError("oracle: context creation failed")
        res["refit"] = lambda row, nfit: self.lsqfit_cubicspline_batch(
            num_pieces, data[row], mask[row], clip_threshold_sigma, nfit, context_npiece=context_npiece,
            want_coeff=False)
        res["num_fitting_max"] = num_fitting_max
        return res

    def lsqfit_sinusoid_batch(self, nwave: Sequence[int], data, mask, clip_threshold_sigma: float,
                              num_fitting_max: int, num_coeff: Optional[int] = None,
                              context_nwave: Optional[int] = None, want_coeff: bool = True,
                              want_best_fit: bool = False, want_residual: bool = True,
                              nthreads: int = 0) -> dict:
        data, mask = self._rows(data, mask)
        R, n = data.shape
        nw = aligned_copy(np.asarray(nwave, dtype=np.uint64))
        if num_coeff is None:
            num_coeff = 2 * len(nw) - 1 if nw[0] == 0 else 2 * len(nw)
        res = self._outputs(R, n, (R, num_coeff), want_coeff, want_best_fit, want_residual)
        rc = self.lib.ref_lsqfit_sinusoid_batch(
            int(nw[-1]) if context_nwave is None else context_nwave, len(nw), _ptr(nw), R, n,
            _ptr(data), n, _ptr(mask), n, clip_threshold_sigma, num_fitting_max, num_coeff,
            _ptr(res["coeff"]), _ptr(res["best_fit"]), _ptr(res["residual"]), _ptr(res["final_mask"]),
            _ptr(res["rms"]), _ptr(res["status"]), _ptr(res["lsqfit_status"]), nthreads)
        if rc != 0:
            raise RuntimeError("oracle: context creation failed")
        res["refit"] = lambda row, nfit: self.lsqfit_sinusoid_batch(
            nwave, data[row], mask[row], clip_threshold_sigma, nfit, num_coeff=num_coeff,
            context_nwave=context_nwave, want_coeff=False)
        res["num_fitting_max"] = num_fitting_max
        return res

    def calibrate_batch(self, scaling, data, reference, in_place: bool = False):
        """sakura_CalibrateDataWith{Array,Const}ScalingFloat row by row. `scaling` is a float
        (const scaling) or an array shaped like data, or one row shared by all rows; `reference`
        likewise. Returns (result [R, n] float32, status [R] int32); with in_place the result
        overwrites (an aligned copy of) data, as the reference allows."""
        data = np.atleast_2d(np.asarray(data, dtype=np.float32))
        R, n = data.shape
        data = aligned_copy(data)
        ref = aligned_copy(np.atleast_2d(np.asarray(reference, dtype=np.float32)))
        ref_stride = 0 if ref.shape[0] == 1 and R > 1 else n
        if np.isscalar(scaling):
            sc, sc_ptr, sc_stride, const = None, None, 0, float(scaling)
        else:
            sc = aligned_copy(np.atleast_2d(np.asarray(scaling, dtype=np.float32)))
            sc_ptr, sc_stride, const = _ptr(sc), (0 if sc.shape[0] == 1 and R > 1 else n), 0.0
        result = data if in_place else aligned_empty((R, n), np.float32)
        status = aligned_empty((R,), np.int32)
        self.lib.ref_calibrate_batch(R, n, sc_ptr, sc_stride, const, _ptr(data), n, _ptr(ref),
                                     ref_stride, _ptr(result), n, _ptr(status))
        return result, status

    FILTER_OPS = {"ranges_inclusive": 0, "ranges_exclusive": 1, "gt": 2, "ge": 3, "lt": 4, "le": 5,
                  "finite": 6, "nonzero": 7, "invert": 8}

    def bool_filter(self, op: str, data, lower=None, upper=None, threshold=0):
        """One call of the reference's bool_filter_collection.cc function selected by `op` and the
        dtype of `data` (float32, int32, uint8, uint32, bool). Returns (result bool[n], status)."""
        data = np.asarray(data)
        type_id = {np.dtype(np.float32): 0, np.dtype(np.int32): 1, np.dtype(np.uint8): 2,
                   np.dtype(np.bool_): 2, np.dtype(np.uint32): 3}[data.dtype]
        n = data.size
        # (numpy gives empty arrays arbitrary addresses: hand over aligned dummies instead)
        data = aligned_copy(data) if n else aligned_empty(8, data.dtype)
        lo = hi = None
        nc = 0
        if lower is not None:
            nc = np.asarray(lower).size
            lo = aligned_copy(np.asarray(lower, dtype=data.dtype)) if nc else aligned_empty(8, data.dtype)
            hi = aligned_copy(np.asarray(upper, dtype=data.dtype)) if nc else aligned_empty(8, data.dtype)
        thr = np.asarray([threshold], dtype=data.dtype if type_id < 2 else np.uint32)
        bits = int(thr.view(np.uint32)[0])
        result = aligned_empty((max(n, 1),), np.bool_)[:n]
        result[...] = False
        st = self.lib.ref_bool_filter(self.FILTER_OPS[op], type_id, n, _ptr(data), nc, _ptr(lo),
                                      _ptr(hi), bits, _ptr(result))
        return result, int(st)

    BIT_OPS = {"and": 0, "converse_nonimplication": 1, "implication": 2, "not": 3, "or": 4, "xor": 5}

    def bit_operation(self, op: str, bit_mask: int, data, edit_mask):
        """sakura_OperateBitwise<Op>{Uint8,Uint32} selected by `op` and the dtype of `data`.
        Returns (result, status)."""
        data = np.asarray(data)
        wide = 1 if data.dtype == np.uint32 else 0
        n = data.size
        d = aligned_copy(data) if n else aligned_empty(8, data.dtype)
        m = aligned_copy(np.asarray(edit_mask).view(np.uint8)) if n else aligned_empty(8, np.uint8)
        out = aligned_empty((max(n, 1),), data.dtype)[:n]
        out[...] = 0
        st = self.lib.ref_bit_operation(self.BIT_OPS[op], wide, int(bit_mask) & 0xffffffff, n, _ptr(d), _ptr(m),
                                        _ptr(out) if n else _ptr(aligned_empty(8, data.dtype)))
        return out, int(st)

    def interpolate(self, axis: str, method: int, base_position, base_data, base_mask,
                    interpolated_position, polynomial_order: int = 0, prefill: float = np.nan):
        """sakura_Interpolate{X,Y}AxisFloat. axis "y": base_data [num_base, num_array] -> result
        [num_interpolated, num_array]; axis "x": [num_array, num_base] -> [num_array, num_interpolated].
        Returns (data, mask, status); arrays without a valid base point keep `prefill`."""
        bd = np.asarray(base_data, dtype=np.float32)
        bm = np.asarray(base_mask, dtype=np.bool_)
        bp = aligned_copy(np.asarray(base_position, dtype=np.float64))
        ip = aligned_copy(np.asarray(interpolated_position, dtype=np.float64))
        nb, ni = bp.size, ip.size
        na = bd.shape[1] if axis == "y" else bd.shape[0]
        shape = (ni, na) if axis == "y" else (na, ni)
        out = aligned_empty(shape, np.float32)
        out[...] = prefill
        om = aligned_empty(shape, np.bool_)
        om[...] = False
        bd, bm = aligned_copy(bd), aligned_copy(bm)  # (named: they must outlive the call)
        st = self.lib.ref_interpolate(1 if axis == "y" else 0, method, polynomial_order, nb, _ptr(bp), na,
                                      _ptr(bd), _ptr(bm), ni, _ptr(ip), _ptr(out), _ptr(om))
        return out, om, int(st)

    def statistics_batch(self, data, mask, accurate: bool = True, nthreads: int = 0):
        data, mask = self._rows(data, mask)
        R, n = data.shape
        data = aligned_copy(data)
        mask = aligned_copy(mask)
        result = aligned_empty((R,), STAT_DTYPE)
        status = aligned_empty((R,), np.int32)
        self.lib.ref_statistics_batch(1 if accurate else 0, R, n, _ptr(data), n, _ptr(mask), n,
                                      _ptr(result), _ptr(status), nthreads)
        return result, status
