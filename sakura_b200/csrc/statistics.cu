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
		Acc *partial, bool vec4) {
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
		auto add = [&](float v, unsigned valid, size_t i) {
			if (valid != 0u) {
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
		};
		size_t scalar_from = lo;
		if (vec4 && lo % 4 == 0) {
			// 128-bit data / 32-bit mask loads, two in flight per thread
			float4 const *d4 = reinterpret_cast<float4 const *>(d);
			unsigned const *m4 = reinterpret_cast<unsigned const *>(m);
			size_t const j_end = hi / 4;
			for (size_t j = lo / 4 + threadIdx.x; j < j_end; j += 2 * blockDim.x) {
				size_t const j2 = j + blockDim.x;
				float4 const v0 = __ldg(d4 + j);
				unsigned const m0 = __ldg(m4 + j);
				float4 v1 = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
				unsigned m1 = 0u;
				if (j2 < j_end) {
					v1 = __ldg(d4 + j2);
					m1 = __ldg(m4 + j2);
				}
				add(v0.x, m0 & 0xffu, 4 * j);
				add(v0.y, m0 & 0xff00u, 4 * j + 1);
				add(v0.z, m0 & 0xff0000u, 4 * j + 2);
				add(v0.w, m0 & 0xff000000u, 4 * j + 3);
				add(v1.x, m1 & 0xffu, 4 * j2);
				add(v1.y, m1 & 0xff00u, 4 * j2 + 1);
				add(v1.z, m1 & 0xff0000u, 4 * j2 + 2);
				add(v1.w, m1 & 0xff000000u, 4 * j2 + 3);
			}
			scalar_from = 4 * j_end;
		}
		for (size_t i = scalar_from + threadIdx.x; i < hi; i += blockDim.x) {
			add(d[i], m[i], i);
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

// ---- rows up to kWarpRowMax elements: one warp per row, 128-bit loads, no block barrier ------
// Per lane: plain FP64 sums of <= kWarpRowMax / 32 addends (a flagged element contributes +0.0 by
// a select on its bits, so a NaN/Inf under the mask cannot leak), float min/max with the first
// index of each; lanes are merged with the same double-double butterfly as above.
constexpr size_t kWarpRowMax = 32768;

struct LaneAcc {
	int count;
	double sum;
	double sq;
	float mn; // NaN until the first valid non-NaN element
	float mx;
	int imn;
	int imx;
};

__device__ __forceinline__ void lane_add(LaneAcc &a, float v, unsigned valid, int i) {
	float const vz = __int_as_float(valid != 0u ? __float_as_int(v) : 0);
	double const vd = static_cast<double>(vz);
	a.sum += vd;
	a.sq = fma(vd, vd, a.sq);
	bool const ok = valid != 0u && v == v;
	bool const lt = ok && !(v >= a.mn); // also true while a.mn is still NaN
	bool const gt = ok && !(v <= a.mx);
	a.mn = lt ? v : a.mn;
	a.imn = lt ? i : a.imn;
	a.mx = gt ? v : a.mx;
	a.imx = gt ? i : a.imx;
}

__global__ void __launch_bounds__(kStatThreads)
StatsWarpKernel(size_t num_rows, int n, float const *data, size_t data_stride, uint8_t const *mask,
		size_t mask_stride, StatResult *result) {
	int const lane = threadIdx.x & 31;
	size_t const warps = static_cast<size_t>(gridDim.x) * (kStatThreads / 32);
	int const n4 = n / 4;
	for (size_t row = static_cast<size_t>(blockIdx.x) * (kStatThreads / 32) + (threadIdx.x >> 5);
			row < num_rows; row += warps) {
		float4 const *d4 = reinterpret_cast<float4 const *>(data + row * data_stride);
		unsigned const *m4 = reinterpret_cast<unsigned const *>(mask + row * mask_stride);
		LaneAcc a;
		a.count = 0;
		a.sum = 0.0;
		a.sq = 0.0;
		a.mn = __int_as_float(0x7fc00000);
		a.mx = __int_as_float(0x7fc00000);
		a.imn = -1;
		a.imx = -1;
		for (int j0 = 0; j0 < n4; j0 += 128) {
			float4 v[4];
			unsigned m[4];
#pragma unroll
			for (int u = 0; u < 4; ++u) {
				int const j = j0 + 32 * u + lane;
				m[u] = 0u;
				v[u] = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
				if (j < n4) {
					v[u] = __ldg(d4 + j);
					m[u] = __ldg(m4 + j);
				}
			}
#pragma unroll
			for (int u = 0; u < 4; ++u) {
				int const i = 4 * (j0 + 32 * u + lane);
				a.count += __popc(__vcmpne4(m[u], 0u) & 0x01010101u);
				lane_add(a, v[u].x, m[u] & 0xffu, i);
				lane_add(a, v[u].y, m[u] & 0xff00u, i + 1);
				lane_add(a, v[u].z, m[u] & 0xff0000u, i + 2);
				lane_add(a, v[u].w, m[u] & 0xff000000u, i + 3);
			}
		}
		Acc t;
		t.count = a.count;
		t.sum = a.sum;
		t.sq = a.sq;
		t.sum_lo = 0.0;
		t.sq_lo = 0.0;
		t.mn = a.mn;
		t.mx = a.mx;
		t.imn = a.imn;
		t.imx = a.imx;
#pragma unroll
		for (int o = 16; o > 0; o >>= 1) {
			Acc const b = acc_shfl_xor(t, o);
			acc_merge(t, b);
		}
		if (lane == 0) {
			store_result(&result[row], t);
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
	if (chunks == 1 && n <= kWarpRowMax && n % 4 == 0 && data_stride % 4 == 0 && mask_stride % 4 == 0
			&& (reinterpret_cast<uintptr_t>(data) % 16) == 0
			&& (reinterpret_cast<uintptr_t>(mask) % 4) == 0) {
		size_t const per_cta = kStatThreads / 32;
		size_t wgrid = (num_rows + per_cta - 1) / per_cta;
		size_t const cap = static_cast<size_t>(num_sms) * 8;
		if (wgrid > cap) {
			wgrid = cap;
		}
		StatsWarpKernel<<<static_cast<unsigned>(wgrid), kStatThreads, 0, stream>>>(num_rows,
				static_cast<int>(n), data, data_stride, mask, mask_stride, result);
		*launches += 1;
		return cudaGetLastError();
	}
	size_t grid = static_cast<size_t>(num_sms) * 8;
	if (grid > total) {
		grid = total;
	}
	bool const vec4 = data_stride % 4 == 0 && mask_stride % 4 == 0 && chunk_len % 4 == 0
			&& (reinterpret_cast<uintptr_t>(data) % 16) == 0
			&& (reinterpret_cast<uintptr_t>(mask) % 4) == 0;
	// long rows (the 1e8-element accuracy case of the reference's tests): double-double sums;
	// short rows: plain FP64 per thread (<= a few hundred addends), double-double in the tree
	if (chunks > 1) {
		StatsKernel<true><<<static_cast<unsigned>(grid), kStatThreads, 0, stream>>>(num_rows, n,
				chunks, chunk_len, data, data_stride, mask, mask_stride, result,
				static_cast<Acc *>(partial_ws), vec4);
	} else {
		StatsKernel<false><<<static_cast<unsigned>(grid), kStatThreads, 0, stream>>>(num_rows, n,
				chunks, chunk_len, data, data_stride, mask, mask_stride, result,
				static_cast<Acc *>(partial_ws), vec4);
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
