// Device building blocks shared by the fit kernels and the standalone
// normal-equation / solver entry points.
#ifndef SAKURA_B200_LSQ_DEVICE_CUH_
#define SAKURA_B200_LSQ_DEVICE_CUH_

#include "device_common.cuh"
#include <float.h>

namespace sakura_b200 {

struct SmemView {
	double *A; // K*K, row-major; becomes the LU factors in place
	double *bvec; // K
	double *cvec; // K   solution of the normal equations
	double *cfull; // 4*npiece  per-piece cubic coefficients (spline)
	double *part; // blockDim.x  partial sums / reduction scratch
	unsigned long long *bound; // npiece+1
	int *row_tr; // K
	int *col_tr; // K
	int *iscr; // 64 ints of scratch
};

// ---- normal equations ---------------------------------------------------------
// G[j][k] = sum_{i:mask} phi_j(i) phi_k(i)   (full K x K, as the reference stores it)
// b[k]    = sum_{i:mask} y_i phi_k(i)
// Each sum is an fma chain like the reference's; channels are split over thread
// groups when there are fewer entries than threads, partials added in group order.
template<typename IndexT>
__device__ void BuildNormalEquations(int n, int K, double const *table, int stride,
		IndexT const *use_idx, uint8_t const *mask, float const *data, SmemView const &s,
		bool sequential = false) {
	int const tid = threadIdx.x;
	int const nt = blockDim.x;
	int const ne = K * K + K;
	if (ne >= nt || sequential) {
		// one thread per entry, channels in ascending order: the reference's own
		// summation order (numeric_operation.cc:209-219, 388-395)
		for (int e = tid; e < ne; e += nt) {
			double acc = 0.0;
			if (e < K * K) {
				int const cj = use_idx[e / K];
				int const ck = use_idx[e % K];
				for (int i = 0; i < n; ++i) {
					if (mask[i]) {
						double const *row = table + static_cast<size_t>(i) * stride;
						acc = fma(row[cj], row[ck], acc);
					}
				}
				s.A[e] = acc;
			} else {
				int const ck = use_idx[e - K * K];
				for (int i = 0; i < n; ++i) {
					if (mask[i]) {
						double const *row = table + static_cast<size_t>(i) * stride;
						acc = fma(static_cast<double>(data[i]), row[ck], acc);
					}
				}
				s.bvec[e - K * K] = acc;
			}
		}
		__syncthreads();
		return;
	}
	int const groups = nt / ne;
	int const g = tid / ne;
	int const e = tid - g * ne;
	if (g < groups) {
		double acc = 0.0;
		if (e < K * K) {
			int const cj = use_idx[e / K];
			int const ck = use_idx[e % K];
			for (int i = g; i < n; i += groups) {
				if (mask[i]) {
					double const *row = table + static_cast<size_t>(i) * stride;
					acc = fma(row[cj], row[ck], acc);
				}
			}
		} else {
			int const ck = use_idx[e - K * K];
			for (int i = g; i < n; i += groups) {
				if (mask[i]) {
					double const *row = table + static_cast<size_t>(i) * stride;
					acc = fma(static_cast<double>(data[i]), row[ck], acc);
				}
			}
		}
		s.part[tid] = acc;
	}
	__syncthreads();
	if (tid < ne) {
		double acc = 0.0;
		for (int gg = 0; gg < groups; ++gg) {
			acc += s.part[gg * ne + tid];
		}
		if (tid < K * K) {
			s.A[tid] = acc;
		} else {
			s.bvec[tid - K * K] = acc;
		}
	}
	__syncthreads();
}

// ---- full-pivot LU solve ------------------------------------------------------
// Same algorithm as Eigen's FullPivLU (the reference's solver,
// numeric_operation.cc:609-622): pivot = first strict maximum of |a_ij| over the
// trailing block scanned column by column; rank threshold eps*K*max|pivot|;
// truncated-rank solution, never an error.
static __device__ void SolveFullPivLU(int K, SmemView const &s) {
	int const tid = threadIdx.x;
	int const nt = blockDim.x;
	int const lane = tid & 31;
	int const warp = tid >> 5;
	int const nwarp = nt >> 5;
	double *A = s.A;
	double *red_v = s.part; // nwarp doubles
	int *red_i = s.iscr; // nwarp ints, then [32]=nonzero_pivots, [33]=stop flag
	double maxpivot = 0.0;
	int nonzero_pivots = K;
	for (int k = 0; k < K; ++k) {
		int const m = K - k;
		double best_v = -1.0;
		int best_i = 0x7fffffff;
		for (int idx = tid; idx < m * m; idx += nt) {
			int const j = k + idx / m; // column (outer)
			int const i = k + idx % m; // row (inner)
			double const v = fabs(A[i * K + j]);
			if (v > best_v) {
				best_v = v;
				best_i = idx;
			}
		}
#pragma unroll
		for (int o = 16; o > 0; o >>= 1) {
			double const ov = __shfl_xor_sync(0xffffffffu, best_v, o);
			int const oi = __shfl_xor_sync(0xffffffffu, best_i, o);
			if (ov > best_v || (ov == best_v && oi < best_i)) {
				best_v = ov;
				best_i = oi;
			}
		}
		if (lane == 0) {
			red_v[warp] = best_v;
			red_i[warp] = best_i;
		}
		__syncthreads();
		best_v = red_v[0];
		best_i = red_i[0];
		for (int w = 1; w < nwarp; ++w) {
			double const ov = red_v[w];
			int const oi = red_i[w];
			if (ov > best_v || (ov == best_v && oi < best_i)) {
				best_v = ov;
				best_i = oi;
			}
		}
		__syncthreads(); // red_* reused next round
		if (best_v == 0.0) {
			nonzero_pivots = k;
			for (int i = k + tid; i < K; i += nt) {
				s.row_tr[i] = i;
				s.col_tr[i] = i;
			}
			break;
		}
		if (best_v > maxpivot) {
			maxpivot = best_v;
		}
		int const pcol = k + best_i / m;
		int const prow = k + best_i % m;
		if (tid == 0) {
			s.row_tr[k] = prow;
			s.col_tr[k] = pcol;
		}
		if (prow != k) {
			for (int j = tid; j < K; j += nt) {
				double const t = A[k * K + j];
				A[k * K + j] = A[prow * K + j];
				A[prow * K + j] = t;
			}
			__syncthreads();
		}
		if (pcol != k) {
			for (int i = tid; i < K; i += nt) {
				double const t = A[i * K + k];
				A[i * K + k] = A[i * K + pcol];
				A[i * K + pcol] = t;
			}
			__syncthreads();
		}
		if (k < K - 1) {
			double const pivot = A[k * K + k];
			for (int i = k + 1 + tid; i < K; i += nt) {
				A[i * K + k] = A[i * K + k] / pivot;
			}
			__syncthreads();
			int const mm = m - 1;
			for (int idx = tid; idx < mm * mm; idx += nt) {
				int const i = k + 1 + idx / mm;
				int const j = k + 1 + idx % mm;
				A[i * K + j] = fma(-A[i * K + k], A[k * K + j], A[i * K + j]);
			}
			__syncthreads();
		}
	}
	__syncthreads();
	// rank
	double const threshold = fabs(maxpivot) * (DBL_EPSILON * static_cast<double>(K));
	int rank = 0;
	for (int i = 0; i < nonzero_pivots; ++i) {
		rank += (fabs(A[i * K + i]) > threshold) ? 1 : 0;
	}
	double *c = s.bvec; // solved in place
	if (warp == 0) {
		if (lane == 0) {
			for (int k = 0; k < K; ++k) {
				int const r = s.row_tr[k];
				if (r != k) {
					double const t = c[k];
					c[k] = c[r];
					c[r] = t;
				}
			}
		}
		__syncwarp();
		for (int j = 0; j < K; ++j) {
			double const cj = c[j];
			for (int i = j + 1 + lane; i < K; i += 32) {
				c[i] = fma(-A[i * K + j], cj, c[i]);
			}
			__syncwarp();
		}
		for (int j = rank - 1; j >= 0; --j) {
			if (lane == 0) {
				c[j] = c[j] / A[j * K + j];
			}
			__syncwarp();
			double const cj = c[j];
			for (int i = lane; i < j; i += 32) {
				c[i] = fma(-A[i * K + j], cj, c[i]);
			}
			__syncwarp();
		}
		// x = Q c ; Q from the column transpositions applied in order
		if (lane == 0) {
			int *q = s.row_tr; // row transpositions no longer needed
			for (int i = 0; i < K; ++i) {
				q[i] = i;
			}
			for (int k = 0; k < K; ++k) {
				int const t = q[k];
				q[k] = q[s.col_tr[k]];
				q[s.col_tr[k]] = t;
			}
			for (int i = 0; i < K; ++i) {
				s.cvec[i] = 0.0;
			}
			for (int i = 0; i < rank; ++i) {
				s.cvec[q[i]] = c[i];
			}
		}
	}
	__syncthreads();
}


} // namespace sakura_b200

#endif
