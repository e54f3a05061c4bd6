"""CPU-side checks of libsakura_b200.so: it loads without a GPU, exports every symbol that
include/sakura_b200.h declares, and the argument validation of the sakura.h contract
(libsakura/test/lsq.cc:374-868 NULL / not-aligned / out-of-range cases -> kInvalidArgument,
status set to kNG first) runs before any CUDA call. No compute is attempted here."""
import ctypes as C
import os
import re
import subprocess

import numpy as np
import pytest

from sakura_b200.capi import (ALIGNMENT, LSQFIT_NG, LSQFIT_TYPE_CHEBYSHEV, LSQFIT_TYPE_POLYNOMIAL,
                              STATUS_INVALID_ARGUMENT, STATUS_OK, aligned_copy, aligned_empty)

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "sakura_b200.h")
IA = STATUS_INVALID_ARGUMENT
OK = 0


def declared_symbols():
    text = open(HEADER).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(sakura_(?:b200_)?[A-Za-z0-9]+)\s*\(", text)))


def test_header_is_valid_c(tmp_path):
    src = tmp_path / "c_interface.c"
    src.write_text('#include "sakura_b200.h"\nint main(void){return (int)sakura_Status_kOK;}\n')
    subprocess.run(["gcc", "-std=c99", "-Wall", "-Werror", "-I", os.path.join(ROOT, "include"),
                    "-fsyntax-only", str(src)], check=True)


def test_exports_every_declared_symbol(libpath):
    names = declared_symbols()
    assert len(names) >= 35
    out = subprocess.run(["nm", "-D", "--defined-only", libpath], check=True, capture_output=True,
                         text=True).stdout
    exported = {line.split()[-1] for line in out.splitlines() if line.strip()}
    missing = [n for n in names if n not in exported]
    assert not missing, missing
    lib = C.CDLL(libpath)
    for n in names:
        assert getattr(lib, n) is not None


def test_alignment_helpers(hostlib):
    L = hostlib.lib
    assert L.sakura_GetAlignment() == ALIGNMENT == 32
    buf = aligned_empty(256, np.uint8)
    assert L.sakura_IsAligned(buf.ctypes.data)
    assert not L.sakura_IsAligned(buf.ctypes.data + 1)
    L.sakura_AlignAny.restype = C.c_void_p
    L.sakura_AlignAny.argtypes = [C.c_size_t, C.c_void_p, C.c_size_t]
    p = L.sakura_AlignAny(200, buf.ctypes.data + 1, 100)
    assert p == buf.ctypes.data + 32
    assert L.sakura_AlignAny(100, buf.ctypes.data + 1, 100) is None  # arena too small
    assert L.sakura_AlignAny(100, None, 10) is None
    L.sakura_AlignFloat.restype = C.c_void_p
    L.sakura_AlignFloat.argtypes = [C.c_size_t, C.c_void_p, C.c_size_t]
    assert L.sakura_AlignFloat(50, buf.ctypes.data + 4, 40) == buf.ctypes.data + 32
    assert b"sm_100a" in L.sakura_b200_GetVersionString()
    assert hostlib.kernel_launch_count() == 0


def test_context_creation_contract(hostlib):
    # lsq.cc:374-437: order/num_data bounds, NULL context
    st, ctx = hostlib.create_polynomial_context(3, 15)
    assert st == STATUS_OK
    assert hostlib.get_number_of_coefficients(ctx, 2) == (STATUS_OK, 3)
    assert hostlib.get_number_of_coefficients(ctx, 4)[0] == IA
    assert hostlib.destroy_context(ctx) == STATUS_OK
    assert hostlib.create_polynomial_context(3, 4)[0] == STATUS_OK  # num_dataLB
    assert hostlib.create_polynomial_context(3, 3)[0] == IA  # num_dataTL
    assert hostlib.create_polynomial_context(3, 15, lsqfit_type=LSQFIT_TYPE_CHEBYSHEV)[0] == STATUS_OK
    assert hostlib.create_polynomial_context(3, 15, lsqfit_type=7)[0] == IA
    assert hostlib.lib.sakura_CreateLSQFitContextPolynomialFloat(0, 3, 15, None) == IA
    assert hostlib.create_cubicspline_context(0, 15)[0] == IA
    assert hostlib.create_cubicspline_context(12, 15)[0] == STATUS_OK
    assert hostlib.create_cubicspline_context(13, 15)[0] == IA
    st, ctx = hostlib.create_cubicspline_context(3, 15)
    assert hostlib.get_number_of_coefficients(ctx, 3) == (STATUS_OK, 12)
    assert hostlib.get_number_of_coefficients(ctx, 0)[0] == IA
    hostlib.destroy_context(ctx)
    assert hostlib.create_sinusoid_context(6, 15)[0] == STATUS_OK  # 2*6+2 <= 15
    assert hostlib.create_sinusoid_context(7, 15)[0] == IA
    st, ctx = hostlib.create_sinusoid_context(3, 15)
    assert hostlib.get_number_of_coefficients(ctx, 1)[0] == IA  # not for sinusoid contexts
    hostlib.destroy_context(ctx)
    assert hostlib.lib.sakura_DestroyLSQFitContextFloat(None) == IA


class PolyArgs:
    """Default parameter set of lsq.cc:1284-1303 with one thing broken at a time."""

    def __init__(self, lib, n=15, order=3):
        self.lib = lib
        self.n, self.order = n, order
        st, self.ctx = lib.create_polynomial_context(order, n)
        assert st == STATUS_OK
        self.data = aligned_copy(np.full(n, 10.0, dtype=np.float32))
        self.mask = aligned_copy(np.ones(n, dtype=np.bool_))
        self.coeff = aligned_empty(order + 1, np.float64)
        self.best_fit = aligned_empty(n, np.float32)
        self.residual = aligned_empty(n, np.float32)
        self.final_mask = aligned_empty(n, np.bool_)
        self.rms = C.c_float(-1.0)
        self.lsq = C.c_int(-1)

    def call(self, **kw):
        a = dict(ctx=self.ctx, order=self.order, n=self.n, data=self.data.ctypes.data,
                 mask=self.mask.ctypes.data, clip=3.0, nfit=2, num_coeff=self.order + 1,
                 coeff=self.coeff.ctypes.data, best_fit=self.best_fit.ctypes.data,
                 residual=self.residual.ctypes.data, final_mask=self.final_mask.ctypes.data,
                 rms=C.byref(self.rms), lsq=C.byref(self.lsq))
        a.update(kw)
        self.lsq.value = -1
        return self.lib.lib.sakura_LSQFitPolynomialFloat(
            a["ctx"], a["order"], a["n"], a["data"], a["mask"], a["clip"], a["nfit"], a["num_coeff"],
            a["coeff"], a["best_fit"], a["residual"], a["final_mask"], a["rms"], a["lsq"])


def test_polynomial_fit_invalid_arguments(hostlib):
    p = PolyArgs(hostlib)
    cases = dict(
        contextNULL=dict(ctx=None), orderTG=dict(order=4), num_dataTL=dict(n=14), num_dataTG=dict(n=16),
        dataNULL=dict(data=None), dataNA=dict(data=p.data.ctypes.data + 4),
        maskNULL=dict(mask=None), maskNA=dict(mask=p.mask.ctypes.data + 1),
        clipNeg=dict(clip=-1.0), clipZero=dict(clip=0.0),
        num_coeffTL=dict(num_coeff=3), num_coeffTG=dict(num_coeff=5),
        coeffNA=dict(coeff=p.coeff.ctypes.data + 8), best_fitNA=dict(best_fit=p.best_fit.ctypes.data + 4),
        residualNA=dict(residual=p.residual.ctypes.data + 4),
        final_maskNULL=dict(final_mask=None), final_maskNA=dict(final_mask=p.final_mask.ctypes.data + 1),
        rmsNULL=dict(rms=None))
    for name, kw in cases.items():
        assert p.call(**kw) == IA, name
        assert p.lsq.value == LSQFIT_NG, name  # baseline.cc:1696: set before the checks
    assert p.call(lsq=None) == IA
    # wrong context kinds (contextOR cspline / sinusoid)
    for make in (lambda: hostlib.create_cubicspline_context(3, 15), lambda: hostlib.create_sinusoid_context(3, 15)):
        st, other = make()
        assert p.call(ctx=other) == IA
        hostlib.destroy_context(other)
    assert hostlib.kernel_launch_count() == 0  # nothing above reached the GPU


def test_spline_and_sinusoid_invalid_arguments(hostlib):
    n = 15
    data = aligned_copy(np.full(n, 10.0, dtype=np.float32))
    mask = aligned_copy(np.ones(n, dtype=np.bool_))
    fm = aligned_empty(n, np.bool_)
    rms, lsq = C.c_float(), C.c_int()
    st, cs = hostlib.create_cubicspline_context(3, n)
    coeff = aligned_empty((3, 4), np.float64)
    bnd = aligned_empty(4, np.uint64)
    f = hostlib.lib.sakura_LSQFitCubicSplineFloat

    def spline(ctx=cs, npiece=3, nn=n, d=data.ctypes.data, m=mask.ctypes.data, clip=3.0,
               c=coeff.ctypes.data, final=fm.ctypes.data, b=bnd.ctypes.data, r=C.byref(rms)):
        return f(ctx, npiece, nn, d, m, clip, 2, c, None, None, final, r, b, C.byref(lsq))

    assert spline(npiece=0) == IA and spline(npiece=4) == IA and spline(nn=14) == IA
    assert spline(d=None) == IA and spline(m=mask.ctypes.data + 1) == IA and spline(clip=0.0) == IA
    assert spline(b=None) == IA and spline(b=bnd.ctypes.data + 8) == IA and spline(final=None) == IA
    assert spline(r=None) == IA and spline(c=coeff.ctypes.data + 8) == IA
    st, cp = hostlib.create_polynomial_context(3, n)
    assert spline(ctx=cp) == IA
    hostlib.destroy_context(cs)

    st, ss = hostlib.create_sinusoid_context(3, n)
    nw = aligned_copy(np.array([0, 1, 2], dtype=np.uint64))
    cf = aligned_empty(7, np.float64)
    g = hostlib.lib.sakura_LSQFitSinusoidFloat

    def sinus(ctx=ss, k=3, w=nw.ctypes.data, nn=n, ncoef=5, clip=3.0):
        return g(ctx, k, w, nn, data.ctypes.data, mask.ctypes.data, clip, 2, ncoef, cf.ctypes.data, None,
                 None, fm.ctypes.data, C.byref(rms), C.byref(lsq))

    assert sinus(k=0) == IA and sinus(w=None) == IA and sinus(nn=16) == IA
    assert sinus(ncoef=4) == IA and sinus(ncoef=8) == IA and sinus(clip=-2.0) == IA
    bad = aligned_copy(np.array([0, 2, 2], dtype=np.uint64))  # not strictly ascending
    assert sinus(w=bad.ctypes.data) == IA
    big = aligned_copy(np.array([0, 1, 4], dtype=np.uint64))  # beyond the context's nwave
    assert sinus(w=big.ctypes.data) == IA
    assert sinus(ctx=cp) == IA
    hostlib.destroy_context(ss)
    hostlib.destroy_context(cp)


def test_statistics_and_blocks_invalid_arguments(hostlib):
    L = hostlib.lib
    d = aligned_copy(np.zeros(8, dtype=np.float32))
    m = aligned_copy(np.ones(8, dtype=np.bool_))
    res = aligned_empty(1, np.dtype("V48"))
    for fn in (L.sakura_ComputeStatisticsFloat, L.sakura_ComputeAccurateStatisticsFloat):
        assert fn(8, None, m.ctypes.data, res.ctypes.data) == IA
        assert fn(8, d.ctypes.data, None, res.ctypes.data) == IA
        assert fn(8, d.ctypes.data + 4, m.ctypes.data, res.ctypes.data) == IA
        assert fn(8, d.ctypes.data, m.ctypes.data, None) == IA
        assert fn((1 << 31), d.ctypes.data, m.ctypes.data, res.ctypes.data) == IA  # statistics.cc:855
    basis = aligned_copy(np.ones((8, 2)))
    use = aligned_copy(np.arange(2, dtype=np.uint64))
    G = aligned_empty(4, np.float64)
    b = aligned_empty(2, np.float64)
    get = L.sakura_GetLSQCoefficientsDouble
    ok = (8, d.ctypes.data, m.ctypes.data, 2, basis.ctypes.data, 2, use.ctypes.data, G.ctypes.data, b.ctypes.data)
    for i, badv in ((0, 0), (1, None), (2, None), (3, 0), (3, 9), (4, None), (5, 0), (5, 3), (6, None), (7, None), (8, None)):
        a = list(ok)
        a[i] = badv
        assert get(*a) == IA, i
    sol = L.sakura_SolveSimultaneousEquationsByLUDouble
    assert sol(2, None, b.ctypes.data, G.ctypes.data) == IA
    assert sol(2, G.ctypes.data, b.ctypes.data, b.ctypes.data) == IA  # in_vector == out
    assert hostlib.kernel_launch_count() == 0


def test_mask_chain_invalid_arguments(hostlib):
    """Mask builders, bit operations, interpolation: the reference's argument rules are applied
    before anything touches the GPU (bool_filter_collection.cc:535-600, bit_operation.cc:41-52,
    interpolation.cc:1288-1340); zero-length calls are kOK without a launch."""
    L = hostlib.lib
    f = aligned_copy(np.zeros(16, dtype=np.float32))
    i = aligned_copy(np.zeros(16, dtype=np.int32))
    u8 = aligned_copy(np.zeros(16, dtype=np.uint8))
    r = aligned_empty(16, np.bool_)
    lo = aligned_copy(np.zeros(2, dtype=np.float32))
    hi = aligned_copy(np.ones(2, dtype=np.float32))
    p = lambda a, off=0: C.c_void_p(a.ctypes.data + off)
    for fn in (L.sakura_SetTrueIfInRangesInclusiveFloat, L.sakura_SetTrueIfInRangesExclusiveFloat):
        assert fn(16, None, 2, p(lo), p(hi), p(r)) == IA
        assert fn(16, p(f, 4), 2, p(lo), p(hi), p(r)) == IA
        assert fn(16, p(f), 2, None, p(hi), p(r)) == IA
        assert fn(16, p(f), 0, None, None, p(r)) == IA
        assert fn(16, p(f), 2, p(lo), p(hi, 4), p(r)) == IA
        assert fn(16, p(f), 2, p(lo), p(hi), None) == IA
        assert fn(0, p(f), 2, p(lo), p(hi), p(r)) == OK
    assert L.sakura_SetTrueIfLessThanInt(16, p(i, 4), 0, p(r)) == IA
    assert L.sakura_SetTrueIfGreaterThanOrEqualsFloat(16, p(f), 0.0, None) == IA
    assert L.sakura_SetFalseIfNanOrInfFloat(16, None, p(r)) == IA
    assert L.sakura_Uint8ToBool(16, p(u8, 1), p(r)) == IA
    assert L.sakura_Uint32ToBool(16, None, p(r)) == IA
    assert L.sakura_InvertBool(16, p(r), p(r, 1)) == IA
    assert L.sakura_InvertBool(0, p(r), p(r)) == OK
    assert L.sakura_OperateBitwiseOrUint8(128, 16, None, p(r), p(u8)) == IA
    assert L.sakura_OperateBitwiseAndUint32(1, 4, p(i), p(r, 1), p(i)) == IA
    assert L.sakura_OperateBitwiseNotUint8(16, p(u8), p(r), None) == IA
    assert L.sakura_OperateBitwiseXorUint8(1, 0, p(u8), p(r), p(u8)) == OK
    xb = aligned_copy(np.array([0.0, 1.0]))
    yb = aligned_copy(np.zeros((2, 4), dtype=np.float32))
    mb = aligned_copy(np.ones((2, 4), dtype=np.bool_))
    xi = aligned_copy(np.array([0.5]))
    out = aligned_empty((1, 4), np.float32)
    om = aligned_empty((1, 4), np.bool_)
    for fn in (L.sakura_InterpolateXAxisFloat, L.sakura_InterpolateYAxisFloat):
        assert fn(1, 0, 0, p(xb), 4, p(yb), p(mb), 1, p(xi), p(out), p(om)) == IA
        assert fn(1, 0, 2, p(xb), 4, p(yb), p(mb), 0, p(xi), p(out), p(om)) == OK
        assert fn(1, 0, 2, p(xb), 0, p(yb), p(mb), 1, p(xi), p(out), p(om)) == OK
        assert fn(1, 0, 2, p(xb, 8), 4, p(yb), p(mb), 1, p(xi), p(out), p(om)) == IA
        assert fn(1, 0, 2, p(xb), 4, None, p(mb), 1, p(xi), p(out), p(om)) == IA
        assert fn(3, 0, 2, p(xb), 4, p(yb), p(mb), 1, p(xi), p(out), p(om)) == IA   # spline: not implemented
        assert fn(7, 0, 2, p(xb), 4, p(yb), p(mb), 1, p(xi), p(out), p(om)) == IA   # no such method
    assert hostlib.kernel_launch_count() == 0
