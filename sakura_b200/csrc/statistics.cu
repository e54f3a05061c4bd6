// Batched masked statistics: count, sum, square_sum (FP64), min, max (+ one index of each).
// Device restatement of sakura_ComputeStatisticsFloat / sakura_ComputeAccurateStatisticsFloat
// (libsakura/src/statistics.cc:338-402 StatsBlock, :595-700, :822-841):
//   * count  = number of valid elements (exact)
//   * sum    = sum of (double)x over valid x           (statistics.cc:393-394)
//   * square_sum = sum of fma(x, x, .) in double         (statistics.cc:395-396)
//   * min/max over valid, non-NaN elements; NaN if none  (statistics.cc:182-188, 369-390)
//   * index_of_min/max = an index whose value equals min/max (the API leaves the choice
//     among ties undefined, sakura.h:250); this kernel returns the smallest such index.
// Summation is a fixed tree (per-thread stride chains -> warp butterfly -> warps in order
// -> chunks in order), so results are deterministic and at least as accurate as the
// reference's blocked pairwise traverse (statistics.cc:126-167).
#include "statistics.cuh"
#include "device_common.cuh"

namespace sakura_b200 {

namespace {

constexpr int kStatThreads = 256;

struct Acc {
	long long count;
	double sum;
	double sq;
	double sum_lo; // low parts of the double-double accumulators
	double sq_lo;
	float mn;
	float mx;
	long long imn;
	long long imx;
};

// (hi, lo) += (v, v_lo) with the rounding error of the high-part addition kept in lo
// (Knuth two-sum): sums stay accurate to ~1e-32 relative whatever the order.
__device__ __forceinline__ void two_sum_acc(double &hi, double &lo, double v, double v_lo) {
	double const s = hi + v;
	double const bb = s - hi;
	double const err = (hi - (s - bb)) + (v - bb);
	hi = s;
	lo += err + v_lo;
}

__device__ __forceinline__ void acc_clear(Acc &a) {
	a.count = 0;
	a.sum = 0.0;
	a.sq = 0.0;
	a.sum_lo = 0.0;
	a.sq_lo = 0.0;
	a.mn = __int_as_float(0x7fc00000);
	a.mx = __int_as_float(0x7fc00000);
	a.imn = -1;
	a.imx = -1;
}

__device__ __forceinline__ void acc_merge(Acc &a, Acc const &b) {
	a.count += b.count;
	two_sum_acc(a.sum, a.sum_lo, b.sum, b.sum_lo);
	two_sum_acc(a.sq, a.sq_lo, b.sq, b.sq_lo);
	if (b.imn >= 0 && (a.imn < 0 || b.mn < a.mn || (b.mn == a.mn && b.imn < a.imn))) {
		a.mn = b.mn;
		a.imn = b.imn;
	}
	if (b.imx >= 0 && (a.imx < 0 || b.mx > a.mx || (b.mx == a.mx && b.imx < a.imx))) {
		a.mx = b.mx;
		a.imx = b.imx;
	}
}

__device__ __forceinline__ Acc acc_shfl_xor(Acc const &a, int o) {
	Acc b;
	b.count = __shfl_xor_sync(0xffffffffu, a.count, o);
	b.sum = __shfl_xor_sync(0xffffffffu, a.sum, o);
	b.sq = __shfl_xor_sync(0xffffffffu, a.sq, o);
	b.sum_lo = __shfl_xor_sync(0xffffffffu, a.sum_lo, o);
	b.sq_lo = __shfl_xor_sync(0xffffffffu, a.sq_lo, o);
	b.mn = __shfl_xor_sync(0xffffffffu, a.mn, o);
	b.mx = __shfl_xor_sync(0xffffffffu, a.mx, o);
	b.imn = __shfl_xor_sync(0xffffffffu, a.imn, o);
	b.imx = __shfl_xor_sync(0xffffffffu, a.imx, o);
	return b;
}

__device__ __forceinline__ void store_result(StatResult *r, Acc const &a) {
	r->count = static_cast<unsigned long long>(a.count);
	r->sum = a.sum + a.sum_lo;
	r->square_sum = a.sq + a.sq_lo;
	r->min = a.mn;
	r->max = a.mx;
	r->index_of_min = a.imn;
	r->index_of_max = a.imx;
}

// One CTA reduces chunk `c` of row `r`: elements [c*chunk_len, min(n,(c+1)*chunk_len)).
template<bool COMPENSATED>
__global__ void __launch_bounds__(kStatThreads)
StatsKernel(size_t num_rows, size_t n, size_t chunks, size_t chunk_len, float const *data,
		size_t data_stride, uint8_t const *mask, size_t mask_stride, StatResult *result,
		Acc *partial) {
	__shared__ Acc warp_acc[kStatThreads / 32];
	size_t const total = num_rows * chunks;
	for (size_t w = blockIdx.x; w < total; w += gridDim.x) {
		size_t const row = w / chunks;
		size_t const c = w - row * chunks;
		float const *d = data + row * data_stride;
		uint8_t const *m = mask + row * mask_stride;
		size_t const lo = c * chunk_len;
		size_t const hi = (lo + chunk_len < n) ? lo + chunk_len : n;
		Acc a;
		acc_clear(a);
		for (size_t i = lo + threadIdx.x; i < hi; i += blockDim.x) {
			if (m[i]) {
				float const v = d[i];
				double const vd = static_cast<double>(v);
				a.count += 1;
				if (COMPENSATED) {
					two_sum_acc(a.sum, a.sum_lo, vd, 0.0);
					two_sum_acc(a.sq, a.sq_lo, vd * vd, 0.0); // exact: 24-bit squared
				} else {
					a.sum += vd;
					a.sq = fma(vd, vd, a.sq);
				}
				if (v == v) {
					if (a.imn < 0 || v < a.mn) {
						a.mn = v;
						a.imn = static_cast<long long>(i);
					}
					if (a.imx < 0 || v > a.mx) {
						a.mx = v;
						a.imx = static_cast<long long>(i);
					}
				}
			}
		}
#pragma unroll
		for (int o = 16; o > 0; o >>= 1) {
			Acc const b = acc_shfl_xor(a, o);
			acc_merge(a, b); // fp add is commutative: both lanes get identical bits
		}
		__syncthreads();
		if ((threadIdx.x & 31) == 0) {
			warp_acc[threadIdx.x >> 5] = a;
		}
		__syncthreads();
		if (threadIdx.x == 0) {
			Acc t = warp_acc[0];
			for (int k = 1; k < kStatThreads / 32; ++k) {
				acc_merge(t, warp_acc[k]);
			}
			if (chunks == 1) {
				store_result(&result[row], t);
			} else {
				partial[w] = t;
			}
		}
	}
}

__global__ void StatsFinalizeKernel(size_t num_rows, size_t chunks, Acc const *partial,
		StatResult *result) {
	size_t const row = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x;
	if (row >= num_rows) {
		return;
	}
	Acc t = partial[row * chunks];
	for (size_t c = 1; c < chunks; ++c) {
		acc_merge(t, partial[row * chunks + c]);
	}
	store_result(&result[row], t);
}

} // namespace

size_t StatsPartialBytes(size_t num_rows, size_t n, size_t *chunks_out, size_t *chunk_len_out) {
	// Long rows (the reference's 81.92M-element accuracy case) are split into chunks so
	// that the whole GPU works on a single row; short rows use one CTA each.
	size_t chunk_len = n;
	size_t chunks = 1;
	size_t const kLong = static_cast<size_t>(kStatThreads) * 256; // 65536 elements
	if (n > kLong && num_rows < 1024) {
		chunk_len = kLong;
		chunks = (n + chunk_len - 1) / chunk_len;
	}
	*chunks_out = chunks;
	*chunk_len_out = chunk_len;
	return chunks == 1 ? 0 : num_rows * chunks * sizeof(Acc);
}

cudaError_t LaunchStats(size_t num_rows, size_t n, float const *data, size_t data_stride,
		uint8_t const *mask, size_t mask_stride, StatResult *result, void *partial_ws,
		int num_sms, cudaStream_t stream, int *launches) {
	size_t chunks, chunk_len;
	(void) StatsPartialBytes(num_rows, n, &chunks, &chunk_len);
	size_t const total = num_rows * chunks;
	if (total == 0) {
		return cudaSuccess;
	}
	size_t grid = static_cast<size_t>(num_sms) * 8;
	if (grid > total) {
		grid = total;
	}
	// long rows (the 1e8-element accuracy case of the reference's tests): double-double sums;
	// short rows: plain FP64 per thread (<= a few hundred addends), double-double in the tree
	if (chunks > 1) {
		StatsKernel<true><<<static_cast<unsigned>(grid), kStatThreads, 0, stream>>>(num_rows, n,
				chunks, chunk_len, data, data_stride, mask, mask_stride, result,
				static_cast<Acc *>(partial_ws));
	} else {
		StatsKernel<false><<<static_cast<unsigned>(grid), kStatThreads, 0, stream>>>(num_rows, n,
				chunks, chunk_len, data, data_stride, mask, mask_stride, result,
				static_cast<Acc *>(partial_ws));
	}
	*launches += 1;
	cudaError_t err = cudaGetLastError();
	if (err != cudaSuccess || chunks == 1) {
		return err;
	}
	unsigned const fb = 128;
	StatsFinalizeKernel<<<static_cast<unsigned>((num_rows + fb - 1) / fb), fb, 0, stream>>>(
			num_rows, chunks, static_cast<Acc const *>(partial_ws), result);
	*launches += 1;
	return cudaGetLastError();
}

} // namespace sakura_b200
