"""sakura_Subtract{Polynomial,CubicSpline,Sinusoid}Float and sakura_GetNumberOfCoefficientsFloat of the
GPU library against the REFERENCE'S OWN functions of the same names (oracle/_ref/libsakura_ref.so is
the reference's object code: baseline.cc:1646, 1902-2017), called with identical coefficients on
identical data. The subtraction is `data - (float) model` with the model an FP64 sum over the basis
table: equal to the reference within one float ulp of the model (the gate of tests/parity.py)."""
import ctypes as C

import numpy as np
import pytest

from sakura_b200 import synth
from sakura_b200.capi import (STATUS_OK, STATUS_INVALID_ARGUMENT, LSQFIT_TYPE_POLYNOMIAL,
                              LSQFIT_TYPE_CHEBYSHEV, aligned_copy, aligned_empty)

pytestmark = pytest.mark.gpu

vp, sz, u16 = C.c_void_p, C.c_size_t, C.c_uint16


class RefLib:
    """The reference's C API through ctypes (contexts are the reference's own)."""

    def __init__(self, oracle_ref):
        L = oracle_ref.lib
        L.sakura_CreateLSQFitContextPolynomialFloat.argtypes = [C.c_int, u16, sz, C.POINTER(vp)]
        L.sakura_CreateLSQFitContextCubicSplineFloat.argtypes = [u16, sz, C.POINTER(vp)]
        L.sakura_CreateLSQFitContextSinusoidFloat.argtypes = [u16, sz, C.POINTER(vp)]
        L.sakura_DestroyLSQFitContextFloat.argtypes = [vp]
        L.sakura_GetNumberOfCoefficientsFloat.argtypes = [vp, u16, C.POINTER(sz)]
        L.sakura_SubtractPolynomialFloat.argtypes = [vp, sz, vp, sz, vp, vp]
        L.sakura_SubtractCubicSplineFloat.argtypes = [vp, sz, vp, sz, vp, vp, vp]
        L.sakura_SubtractSinusoidFloat.argtypes = [vp, sz, vp, sz, vp, sz, vp, vp]
        for f in ("sakura_CreateLSQFitContextPolynomialFloat", "sakura_CreateLSQFitContextCubicSplineFloat",
                  "sakura_CreateLSQFitContextSinusoidFloat", "sakura_DestroyLSQFitContextFloat",
                  "sakura_GetNumberOfCoefficientsFloat", "sakura_SubtractPolynomialFloat",
                  "sakura_SubtractCubicSplineFloat", "sakura_SubtractSinusoidFloat"):
            getattr(L, f).restype = C.c_int
        self.L = L

    def create(self, kind, param, n, lsqfit_type=LSQFIT_TYPE_POLYNOMIAL):
        ctx = vp()
        if kind == "poly":
            st = self.L.sakura_CreateLSQFitContextPolynomialFloat(lsqfit_type, param, n, C.byref(ctx))
        elif kind == "spline":
            st = self.L.sakura_CreateLSQFitContextCubicSplineFloat(param, n, C.byref(ctx))
        else:
            st = self.L.sakura_CreateLSQFitContextSinusoidFloat(param, n, C.byref(ctx))
        return st, ctx


@pytest.fixture(scope="module")
def ref(oracle_ref):
    return RefLib(oracle_ref)


def create_gpu(lib, kind, param, n, lsqfit_type=LSQFIT_TYPE_POLYNOMIAL):
    if kind == "poly":
        return lib.create_polynomial_context(param, n, lsqfit_type)
    if kind == "spline":
        return lib.create_cubicspline_context(param, n)
    return lib.create_sinusoid_context(param, n)


def within_one_ulp_of_model(got, want, data):
    model = np.abs(data.astype(np.float64) - want.astype(np.float64))
    tol = 1.2e-7 * model + 1e-30
    return np.all(np.abs(got.astype(np.float64) - want.astype(np.float64)) <= tol)


@pytest.mark.parametrize("kind,param,lsqfit_type", [("poly", 0, 0), ("poly", 5, 0), ("poly", 9, 0), ("poly", 5, 1),
                                                    ("poly", 12, 1), ("spline", 1, 0), ("spline", 7, 0),
                                                    ("sinusoid", 0, 0), ("sinusoid", 6, 0)])
def test_get_number_of_coefficients(lib, ref, kind, param, lsqfit_type):
    """Same status and same count as the reference for every order up to and beyond the context's
    (baseline.cc:1646-1685)."""
    n = 512
    st_r, ctx_r = ref.create(kind, param, n, lsqfit_type)
    st_g, ctx_g = create_gpu(lib, kind, param, n, lsqfit_type)
    assert st_r == st_g == STATUS_OK
    for order in list(range(0, param + 3)) + [100, 65535]:
        nr, ng = sz(12345), sz(12345)
        sr = ref.L.sakura_GetNumberOfCoefficientsFloat(ctx_r, order, C.byref(nr))
        sg = lib.lib.sakura_GetNumberOfCoefficientsFloat(ctx_g, order, C.byref(ng))
        assert sr == sg, (kind, param, order, sr, sg)
        if sr == STATUS_OK:
            assert nr.value == ng.value, (kind, param, order, nr.value, ng.value)
    # NULL arguments
    nr = sz(0)
    assert lib.lib.sakura_GetNumberOfCoefficientsFloat(None, 1, C.byref(nr)) \
        == ref.L.sakura_GetNumberOfCoefficientsFloat(None, 1, C.byref(nr)) == STATUS_INVALID_ARGUMENT
    assert lib.lib.sakura_GetNumberOfCoefficientsFloat(ctx_g, 0, None) \
        == ref.L.sakura_GetNumberOfCoefficientsFloat(ctx_r, 0, None) == STATUS_INVALID_ARGUMENT
    ref.L.sakura_DestroyLSQFitContextFloat(ctx_r)
    lib.destroy_context(ctx_g)


@pytest.mark.parametrize("n,order,lsqfit_type", [(1024, 5, 0), (4096, 5, 0), (333, 2, 0), (2048, 7, 0),
                                                 (1024, 5, 1), (1024, 11, 1), (64, 0, 0)])
def test_subtract_polynomial(lib, ref, oracle_ref, n, order, lsqfit_type):
    data, mask = synth.make_spectra(6, n, seed=40 + order, ripple=True)
    o = oracle_ref.lsqfit_polynomial_batch(order, data, mask, 3.0, 3, lsqfit_type=lsqfit_type)
    st_r, ctx_r = ref.create("poly", order, n, lsqfit_type)
    st_g, ctx_g = lib.create_polynomial_context(order, n, lsqfit_type)
    assert st_r == st_g == STATUS_OK
    for r in range(6):
        row = aligned_copy(data[r])
        for ncoeff in sorted({order + 1, max(1, order)}):  # also fewer coefficients than the context has
            co = aligned_copy(o["coeff"][r][:ncoeff].copy())
            out_r = aligned_empty(n, np.float32)
            out_g = aligned_empty(n, np.float32)
            sr = ref.L.sakura_SubtractPolynomialFloat(ctx_r, n, row.ctypes.data, ncoeff, co.ctypes.data,
                                                      out_r.ctypes.data)
            sg = lib.lib.sakura_SubtractPolynomialFloat(ctx_g, n, row.ctypes.data, ncoeff, co.ctypes.data,
                                                        out_g.ctypes.data)
            assert sr == sg == STATUS_OK, (sr, sg, lib.last_error())
            assert within_one_ulp_of_model(out_g, out_r, data[r]), float(np.abs(out_g - out_r).max())
            if ncoeff == order + 1:
                # ... and the reference's subtraction reproduces its own fit's residual
                assert within_one_ulp_of_model(out_r, o["residual"][r], data[r])
    # argument rules, side by side
    row = aligned_copy(data[0])
    co = aligned_copy(np.zeros(order + 2))
    out = aligned_empty(n, np.float32)
    for args in ((n, order + 2), (n - 1 if n > 64 else n + 1, order + 1), (n, 0)):
        sr = ref.L.sakura_SubtractPolynomialFloat(ctx_r, args[0], row.ctypes.data, args[1], co.ctypes.data,
                                                  out.ctypes.data)
        sg = lib.lib.sakura_SubtractPolynomialFloat(ctx_g, args[0], row.ctypes.data, args[1], co.ctypes.data,
                                                    out.ctypes.data)
        assert sr == sg == STATUS_INVALID_ARGUMENT, (args, sr, sg)
    ref.L.sakura_DestroyLSQFitContextFloat(ctx_r)
    lib.destroy_context(ctx_g)


@pytest.mark.parametrize("n,npiece", [(1024, 5), (4096, 20), (300, 1), (2048, 3)])
def test_subtract_cubic_spline(lib, ref, oracle_ref, n, npiece):
    data, mask = synth.make_spectra(5, n, seed=60 + npiece, ripple=True)
    o = oracle_ref.lsqfit_cubicspline_batch(npiece, data, mask, 3.0, 3)
    st_r, ctx_r = ref.create("spline", npiece, n)
    st_g, ctx_g = lib.create_cubicspline_context(npiece, n)
    assert st_r == st_g == STATUS_OK
    for r in range(5):
        row = aligned_copy(data[r])
        co = aligned_copy(o["coeff"][r])
        bd = aligned_copy(o["boundary"][r])
        out_r = aligned_empty(n, np.float32)
        out_g = aligned_empty(n, np.float32)
        sr = ref.L.sakura_SubtractCubicSplineFloat(ctx_r, n, row.ctypes.data, npiece, co.ctypes.data,
                                                   bd.ctypes.data, out_r.ctypes.data)
        sg = lib.lib.sakura_SubtractCubicSplineFloat(ctx_g, n, row.ctypes.data, npiece, co.ctypes.data,
                                                     bd.ctypes.data, out_g.ctypes.data)
        assert sr == sg == STATUS_OK, (sr, sg, lib.last_error())
        # identical coefficients, identical per-piece cubic: same float model up to one ulp
        assert within_one_ulp_of_model(out_g, out_r, data[r]), float(np.abs(out_g - out_r).max())
    ref.L.sakura_DestroyLSQFitContextFloat(ctx_r)
    lib.destroy_context(ctx_g)


@pytest.mark.parametrize("n,nwave", [(1024, [0, 1, 2, 3, 4]), (2048, [1, 3]), (512, [0]), (4096, list(range(0, 11)))])
def test_subtract_sinusoid(lib, ref, oracle_ref, n, nwave):
    data, mask = synth.make_spectra(5, n, seed=70 + len(nwave), ripple=True)
    o = oracle_ref.lsqfit_sinusoid_batch(nwave, data, mask, 3.0, 3)
    K = o["coeff"].shape[1]
    st_r, ctx_r = ref.create("sinusoid", max(nwave), n)
    st_g, ctx_g = lib.create_sinusoid_context(max(nwave), n)
    assert st_r == st_g == STATUS_OK
    nw = aligned_copy(np.asarray(nwave, dtype=np.uint64))
    for r in range(5):
        row = aligned_copy(data[r])
        co = aligned_copy(o["coeff"][r])
        out_r = aligned_empty(n, np.float32)
        out_g = aligned_empty(n, np.float32)
        sr = ref.L.sakura_SubtractSinusoidFloat(ctx_r, n, row.ctypes.data, len(nwave), nw.ctypes.data, K,
                                                co.ctypes.data, out_r.ctypes.data)
        sg = lib.lib.sakura_SubtractSinusoidFloat(ctx_g, n, row.ctypes.data, len(nwave), nw.ctypes.data, K,
                                                  co.ctypes.data, out_g.ctypes.data)
        assert sr == sg == STATUS_OK, (sr, sg, lib.last_error())
        assert within_one_ulp_of_model(out_g, out_r, data[r]), float(np.abs(out_g - out_r).max())
    ref.L.sakura_DestroyLSQFitContextFloat(ctx_r)
    lib.destroy_context(ctx_g)
