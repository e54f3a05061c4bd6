// Device helpers shared by the polynomial fit kernels (lsqfit_poly.cu, lsqfit_poly_screen.cu).
#ifndef SAKURA_B200_POLY_DEVICE_CUH_
#define SAKURA_B200_POLY_DEVICE_CUH_

#include <float.h>
#include <stdint.h>
#include <cuda_runtime.h>

namespace sakura_b200 {

constexpr int kMaxPolyK = 8;
constexpr int kMaxMoments = 2 * kMaxPolyK - 1;

struct PolyConst {
	double h; // 1/(n-1)
	double hpow[kMaxMoments]; // h^k
	double s_all[kMaxMoments]; // sum_{i=0}^{n-1} (i h)^k, correctly rounded
	double factor[kMaxPolyK]; // (n-1)^k by repeated multiplication (baseline.cc:1313-1318)
	// Chebyshev coefficients from power-basis ones: a_j = sum_{k>=j} cheb[j][k] c_k, where
	// sum_k c_k x^k = sum_j a_j T_j(2x - 1), x = i/(n-1) (baseline.cc:299-323 basis)
	double cheb[kMaxPolyK][kMaxPolyK];
	// 32 Chebyshev nodes mapped to x in [0,1], and the Ehlich-Zeller factors 1/cos(k pi/64):
	// max over [0,1] of a polynomial of degree k <= ez[k] * max over the nodes (screening kernel)
	double node[32];
	double ez[kMaxPolyK];
	// u^j = sum_{k<=j} binom[j][k] C(u + k, k) (integers): turns the running sums of running sums the
	// screening kernel accumulates into power moments
	double binom[kMaxPolyK][kMaxPolyK];
};

/** Per-channel-count constants, built once on the host (lsqfit_poly.cu). */
PolyConst const &GetPolyConst(int n);

/** value if `keep` is non-zero, else +0.0f (a select on the bit pattern: a NaN/Inf under the mask
 *  must not leak, so this cannot be a multiplication by 0/1). */
__device__ __forceinline__ float select_or_zero(float value, uint32_t keep) {
	return __int_as_float(keep != 0u ? __float_as_int(value) : 0);
}

/** Coefficients of p(zb + u) in powers of u from those of p(z) in powers of z
 *  (repeated synthetic division: K(K-1)/2 FMAs). */
template<int K>
__device__ __forceinline__ void TaylorShift(double const *c, double zb, double *d) {
#pragma unroll
	for (int k = 0; k < K; ++k) {
		d[k] = c[k];
	}
#pragma unroll
	for (int i = 0; i < K - 1; ++i) {
#pragma unroll
		for (int j = K - 2; j >= i; --j) {
			d[j] = fma(zb, d[j + 1], d[j]);
		}
	}
}

__device__ __forceinline__ double u32_to_double(uint32_t v) {
	// exact, without the (slow) conversion pipe
	return __hiloint2double(0x43300000, static_cast<int>(v)) - 4503599627370496.0;
}

// Reciprocal of a positive, normal double: hardware estimate + one cubic step (4 instructions
// instead of the ~25 of an IEEE division; measured DDIV throughput on B200 is 2.6 /clk/SM).
// Accurate to ~1 ulp, which is all the elimination below needs.
__device__ __forceinline__ double fast_rcp(double x) {
	double r;
	asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x));
	// one third-order step: r (1 + e + e^2) = (1 - e^3) / x, e ~ 2^-23 for the hardware estimate
	double const e = fma(-x, r, 1.0);
	return fma(r, fma(e, e, e), r);
}

// ---- solve G c = T, G[j][k] = S[j+k], entirely in registers (every lane the same work) -------
// LDL^T / Gaussian elimination in natural order on the upper triangle. G is a Gram matrix:
// symmetric positive definite unless the row is degenerate. Returns false when a pivot is not
// safely positive; the caller then uses the full-pivot rank-revealing solver.
template<int K>
__device__ __forceinline__ bool SolveLDLT(double const *S, double const *T, double *c_out) {
	double a[K][K];
	double c[K];
	double rcp[K];
#pragma unroll
	for (int i = 0; i < K; ++i) {
#pragma unroll
		for (int j = i; j < K; ++j) {
			a[i][j] = S[i + j];
		}
		c[i] = T[i];
	}
	double maxpivot = 0.0;
	bool ok = true;
#pragma unroll
	for (int k = 0; k < K; ++k) {
		double const pivot = a[k][k];
		double const apivot = fabs(pivot);
		maxpivot = (apivot > maxpivot) ? apivot : maxpivot; // (a NaN pivot fails the test below)
		ok = ok && (pivot > maxpivot * (DBL_EPSILON * K));
		rcp[k] = fast_rcp(pivot);
#pragma unroll
		for (int i = k + 1; i < K; ++i) {
			double const l = a[k][i] * rcp[k];
#pragma unroll
			for (int j = i; j < K; ++j) {
				a[i][j] = fma(-l, a[k][j], a[i][j]);
			}
			c[i] = fma(-l, c[k], c[i]);
		}
	}
#pragma unroll
	for (int j = K - 1; j >= 0; --j) {
		c[j] *= rcp[j];
#pragma unroll
		for (int i = 0; i < j; ++i) {
			c[i] = fma(-a[i][j], c[j], c[i]);
		}
	}
#pragma unroll
	for (int i = 0; i < K; ++i) {
		c_out[i] = c[i];
	}
	return ok;
}

} // namespace sakura_b200

#endif
