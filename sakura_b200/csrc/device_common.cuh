// Shared device-side helpers for the sakura_b200 kernels (sm_100a).
// All floating-point code in this library is compiled with --fmad=false: every
// fused multiply-add is written explicitly (fma / __fmaf_rn) exactly where the
// reference's compiled code fuses (see DESIGN.md "Arithmetic contract").
#ifndef SAKURA_B200_DEVICE_COMMON_CUH_
#define SAKURA_B200_DEVICE_COMMON_CUH_

#include <cuda_runtime.h>
#include <stdint.h>
#include <math.h>

namespace sakura_b200 {

constexpr int kWarp = 32;

enum FitType : int {
	kFitPolynomial = 0, kFitChebyshev = 1, kFitCubicSpline = 2, kFitSinusoid = 3
};

// values of sakura_Status / sakura_LSQFitStatus used on the device
constexpr int kStatusOK = 0;
constexpr int kStatusNG = 1;
constexpr int kLsqOK = 0;
constexpr int kLsqNG = 1;
constexpr int kLsqNotEnoughData = 2;

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
	for (int o = 16; o > 0; o >>= 1) {
		v += __shfl_xor_sync(0xffffffffu, v, o);
	}
	return v;
}

__device__ __forceinline__ int warp_sum(int v) {
#pragma unroll
	for (int o = 16; o > 0; o >>= 1) {
		v += __shfl_xor_sync(0xffffffffu, v, o);
	}
	return v;
}

__device__ __forceinline__ void dmma8x8x4(double &c0, double &c1, double a, double b) {
	// (not volatile: a pure function of its operands, so that independent products can be interleaved)
	asm("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
			: "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}

/** Sum of `v` over the 32 lanes, delivered to every lane: two FP64 tensor-core products with a matrix
 *  of ones (B[k][n] comes from lane k + 4 n, so the first product adds groups of four lanes, the second
 *  the eight group sums) instead of five shuffle rounds. Fixed order: deterministic. */
__device__ __forceinline__ double warp_total(double v) {
	double c0 = 0.0, c1 = 0.0;
	dmma8x8x4(c0, c1, 1.0, v);
	double const h = c0 + c1;
	double d0 = 0.0, d1 = 0.0;
	dmma8x8x4(d0, d1, 1.0, h);
	return d0;
}

/**
 * Deterministic block-wide sum of (count, s1, s2). `scratch` must hold
 * 3 * (blockDim.x/32) doubles. Every thread receives the totals.
 * Order: xor-butterfly inside each warp, then warps added in ascending order.
 */
__device__ __forceinline__ void block_sum3(int &cnt, double &s1, double &s2, double *scratch) {
	int const lane = threadIdx.x & 31;
	int const warp = threadIdx.x >> 5;
	int const nwarp = (blockDim.x + 31) >> 5;
	cnt = warp_sum(cnt);
	s1 = warp_sum(s1);
	s2 = warp_sum(s2);
	__syncthreads(); // scratch may still be read from a previous call
	if (lane == 0) {
		scratch[warp] = static_cast<double>(cnt);
		scratch[nwarp + warp] = s1;
		scratch[2 * nwarp + warp] = s2;
	}
	__syncthreads();
	double c = 0.0, a = 0.0, b = 0.0;
	for (int w = 0; w < nwarp; ++w) {
		c += scratch[w];
		a += scratch[nwarp + w];
		b += scratch[2 * nwarp + w];
	}
	cnt = static_cast<int>(c);
	s1 = a;
	s2 = b;
}

__device__ __forceinline__ int block_sum_int(int v, int *scratch) {
	int const lane = threadIdx.x & 31;
	int const warp = threadIdx.x >> 5;
	int const nwarp = (blockDim.x + 31) >> 5;
	v = warp_sum(v);
	__syncthreads();
	if (lane == 0) {
		scratch[warp] = v;
	}
	__syncthreads();
	int t = 0;
	for (int w = 0; w < nwarp; ++w) {
		t += scratch[w];
	}
	return t;
}

/**
 * mean / rms / clip thresholds exactly as the reference computes them
 * (baseline.cc:1200-1207, as compiled: var = fnmadd(mean, mean, sq/count);
 * thresholds rounded to float).
 */
struct ClipThresholds {
	double mean;
	double rms;
	float lower;
	float upper;
};

__device__ __forceinline__ ClipThresholds make_thresholds(int count, double sum, double square_sum,
		float clip_threshold_sigma) {
	ClipThresholds t;
	double const cnt = static_cast<double>(count);
	t.mean = sum / cnt;
	double const q = square_sum / cnt;
	t.rms = sqrt(fabs(fma(-t.mean, t.mean, q)));
	float const thr = static_cast<float>(static_cast<double>(clip_threshold_sigma) * t.rms);
	t.lower = static_cast<float>(t.mean - static_cast<double>(thr));
	t.upper = static_cast<float>(t.mean + static_cast<double>(thr));
	return t;
}

/** The reference's clip predicate, in float (baseline.cc:1088). */
__device__ __forceinline__ bool is_clipped(float r, float lower, float upper) {
	float const a = __fsub_rn(r, lower);
	float const b = __fsub_rn(upper, r);
	return __fmul_rn(a, b) < 0.0f;
}

} // namespace sakura_b200

#endif
