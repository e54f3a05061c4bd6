// General batched LSQ-fit kernel: polynomial / Chebyshev / cubic spline / sinusoid,
// any number of fitted bases K <= kMaxBases, any channel count.
//
// One CTA per spectrum. It restates, per row, the reference's DoLSQFit loop
// (libsakura/src/baseline.cc:1115-1224) with the wrappers LSQFit (:1268-1329) and
// LSQFitCubicSpline (:1342-1409):
//   count unmasked -> [spline: boundaries (:808-844) + per-row basis (:911-980)] ->
//   repeat { normal equations (numeric_operation.cc:192-225, 370-396) ->
//            full-pivot LU solve (numeric_operation.cc:609-622, Eigen FullPivLU) ->
//            model + residual (baseline.cc:661-758, 629-642) ->
//            accurate statistics (statistics.cc:822-841) -> thresholds (baseline.cc:1200-1207) ->
//            clip (baseline.cc:1063-1098) } -> outputs.
// The normal equations are rebuilt from the current mask in every iteration (the
// reference's incremental downdate, numeric_operation.cc:283-314, differs from a
// rebuild only by rounding; baseline.cc:1166 switches between the two itself).
//
// This kernel is the correctness path for every shape the API accepts; the
// polynomial hot path has its own register-resident kernel (lsqfit_poly.cu).
#include "lsqfit_general.cuh"
#include "device_common.cuh"
#include "lsq_device.cuh"

#include <float.h>

namespace sakura_b200 {

namespace {

__device__ __forceinline__ SmemView carve(unsigned char *base, int K, int npiece, int threads) {
	SmemView s;
	double *d = reinterpret_cast<double *>(base);
	s.A = d;
	d += static_cast<size_t>(K) * K;
	s.bvec = d;
	d += K;
	s.cvec = d;
	d += K;
	s.cfull = d;
	d += 4 * (npiece > 0 ? npiece : 1);
	s.part = d;
	d += threads;
	s.bound = reinterpret_cast<unsigned long long *>(d);
	d += (npiece + 2);
	int *ip = reinterpret_cast<int *>(d);
	s.row_tr = ip;
	ip += K;
	s.col_tr = ip;
	ip += K;
	s.iscr = ip;
	return s;
}

// ---- cubic spline helpers -------------------------------------------------------

__device__ __forceinline__ double nonneg(double v) {
	return (0.0 < v) ? v : 0.0; // std::max(0.0, v), baseline.cc:924
}

// baseline.cc:808-844: boundary[idx] (1 <= idx < npiece) is the unmasked channel
// whose rank among the unmasked channels equals ceil(N*idx/npiece).
__device__ void SplineBoundaries(int n, int npiece, int num_unmasked, uint8_t const *mask,
		SmemView const &s) {
	int const tid = threadIdx.x;
	int const nt = blockDim.x;
	int const chunk = (n + nt - 1) / nt;
	int const lo = min(n, tid * chunk);
	int const hi = min(n, lo + chunk);
	int cnt = 0;
	for (int i = lo; i < hi; ++i) {
		cnt += mask[i] ? 1 : 0;
	}
	int *counts = reinterpret_cast<int *>(s.part); // nt ints
	__syncthreads();
	counts[tid] = cnt;
	__syncthreads();
	int start = 0;
	for (int t = 0; t < tid; ++t) {
		start += counts[t];
	}
	if (tid == 0) {
		s.bound[0] = 0ull;
		s.bound[npiece] = static_cast<unsigned long long>(n);
	}
	long long const N = num_unmasked;
	for (int idx = 1; idx < npiece; ++idx) {
		long long const target = (N * idx + npiece - 1) / npiece;
		if (target >= start && target < start + cnt) {
			int r = start;
			for (int i = lo; i < hi; ++i) {
				if (mask[i]) {
					if (r == target) {
						s.bound[idx] = static_cast<unsigned long long>(i);
						break;
					}
					++r;
				}
			}
		}
	}
	__syncthreads();
}

// baseline.cc:952-980 + 911-940 + 251-264: per-row basis [n][stride], K = npiece+3 columns.
__device__ void SplineBasisTable(int n, int npiece, int stride, double *table, SmemView const &s) {
	double const max_x = static_cast<double>(n - 1);
	int const num_boundary = npiece + 1;
	for (int i = threadIdx.x; i < n; i += blockDim.x) {
		double *row = table + static_cast<size_t>(i) * stride;
		double const i_d = static_cast<double>(i);
		double const x = i_d / max_x;
		double val = 1.0;
		row[0] = val;
		val *= x;
		row[1] = val;
		val *= x;
		row[2] = val;
		val *= x;
		double x3 = val;
		if (num_boundary > 2) {
			double v = (i_d - static_cast<double>(s.bound[1])) / max_x;
			double cube_prev = (v * v) * v;
			x3 -= nonneg(cube_prev);
			for (int j = 2; j < num_boundary; ++j) {
				double w = (i_d - static_cast<double>(s.bound[j])) / max_x;
				double const cube_cur = (w * w) * w;
				row[2 + j] = nonneg(cube_prev - nonneg(cube_cur));
				cube_prev = cube_cur;
			}
		}
		row[3] = x3;
	}
	__syncthreads();
}

// baseline.cc:868-893 as compiled (fused where GCC fuses)
__device__ void SplineFullCoefficients(int n, int npiece, SmemView const &s) {
	if (threadIdx.x == 0) {
		double const *raw = s.cvec;
		double *cf = s.cfull;
		for (int i = 0; i < 4; ++i) {
			cf[i] = raw[i];
		}
		double const max_x = static_cast<double>(n - 1);
		for (int i = 1; i < npiece; ++i) {
			int const j = i + 3;
			double const c = raw[j] - cf[4 * (i - 1) + 3];
			double const b = static_cast<double>(s.bound[i]) / max_x;
			double const b3 = b * 3.0;
			double const bb = b * b;
			double const b3b = b3 * b;
			double const bbb = bb * b;
			cf[4 * i + 0] = fma(-bbb, c, cf[4 * (i - 1) + 0]);
			cf[4 * i + 1] = fma(b3b, c, cf[4 * (i - 1) + 1]);
			cf[4 * i + 2] = fma(-b3, c, cf[4 * (i - 1) + 2]);
			cf[4 * i + 3] = raw[j];
		}
	}
	__syncthreads();
}

__device__ __forceinline__ int PieceOf(int i, int npiece, unsigned long long const *bound) {
	int lo = 0, hi = npiece; // bound[lo] <= i < bound[hi]
	while (hi - lo > 1) {
		int const mid = (lo + hi) >> 1;
		if (static_cast<unsigned long long>(i) >= bound[mid]) {
			lo = mid;
		} else {
			hi = mid;
		}
	}
	return lo;
}

// baseline.cc:724-758: products first, then left-to-right adds, no fma
__device__ __forceinline__ double SplineModel(int i, int n, int npiece, SmemView const &s) {
	double const x = static_cast<double>(i) / static_cast<double>(n - 1);
	double const x2 = x * x;
	double const x3 = x2 * x;
	double const *c = s.cfull + 4 * PieceOf(i, npiece, s.bound);
	double const p0 = c[0] * 1.0;
	double const p1 = c[1] * x;
	double const p2 = c[2] * x2;
	double const p3 = c[3] * x3;
	return ((p0 + p1) + p2) + p3;
}

__global__ void __launch_bounds__(kGeneralThreads)
GeneralFitKernel(GeneralFitParams const p) {
	extern __shared__ __align__(16) unsigned char smem_raw[];
	SmemView const s = carve(smem_raw, p.K, p.npiece, blockDim.x);
	int const tid = threadIdx.x;
	int const nt = blockDim.x;
	int const n = p.n;
	int const K = p.K;
	bool const spline = (p.fit_type == kFitCubicSpline);

	// Optional indirection: fit only the rows listed in row_index[0 .. *row_count) (used for the
	// rows a specialised kernel hands over, e.g. nearly singular normal equations).
	size_t const total_rows = p.row_count ? static_cast<size_t>(*p.row_count) : p.num_spectra;
	for (size_t it = blockIdx.x; it < total_rows; it += gridDim.x) {
		size_t const row = p.row_index ? static_cast<size_t>(p.row_index[it]) : it;
		float const *data = p.data + row * p.data_stride;
		uint8_t const *mask = p.mask + row * p.mask_stride;
		uint8_t *fmask = p.final_mask + row * p.mask_stride;
		// Results are staged in per-CTA scratch and published only when the row
		// succeeds, like the reference's context scratch (baseline.cc:1321-1328).
		float *resid = p.ws_residual + static_cast<size_t>(blockIdx.x) * n;
		float *bestfit = p.best_fit ? p.ws_best_fit + static_cast<size_t>(blockIdx.x) * n : nullptr;
		double const *table = p.basis;
		int stride = p.nb_ctx;
		__syncthreads(); // previous row's shared state fully consumed

		// baseline.cc:1696 / 1771 / 1845: status starts as NG
		int cnt = 0;
		for (int i = tid; i < n; i += nt) {
			cnt += mask[i] ? 1 : 0;
		}
		int const num_valid = block_sum_int(cnt, s.iscr);

		if (spline) {
			if (num_valid < p.npiece) { // baseline.cc:822-824 -> Status_kNG, lsqfit_status stays kNG
				if (tid == 0) {
					p.status[row] = kStatusNG;
					p.lsqfit_status[row] = kLsqNG;
				}
				continue;
			}
			SplineBoundaries(n, p.npiece, num_valid, mask, s);
			if (p.boundary != nullptr) {
				for (int i = tid; i <= p.npiece; i += nt) {
					p.boundary[row * (p.npiece + 1) + i] = s.bound[i];
				}
			}
			if (tid == 0 && p.flags != nullptr) {
				p.flags[row] = 4;
			}
		}
		if (p.num_fitting_max == 0) { // baseline.cc:1295-1296, 1371-1372
			if (tid == 0) {
				p.status[row] = kStatusOK;
				p.lsqfit_status[row] = kLsqOK;
			}
			continue;
		}
		if (num_valid < K) { // baseline.cc:1147-1150
			if (tid == 0) {
				p.status[row] = kStatusNG;
				p.lsqfit_status[row] = kLsqNotEnoughData;
			}
			continue;
		}
		if (spline) {
			double *tbl = p.ws_spline_basis + static_cast<size_t>(blockIdx.x) * n * p.nb_row;
			SplineBasisTable(n, p.npiece, p.nb_row, tbl, s);
			table = tbl;
			stride = p.nb_row;
		}
		// baseline.cc:1154-1156
		if (fmask != mask) {
			for (int i = tid; i < n; i += nt) {
				fmask[i] = mask[i];
			}
		}
		__syncthreads();

		int num_unmasked = n; // baseline.cc:1152 (sic)
		double rms_d = 0.0;
		bool failed = false;
		for (int iter = 1; iter <= p.num_fitting_max; ++iter) {
			if (num_unmasked < K) { // baseline.cc:1162-1165
				failed = true;
				break;
			}
			BuildNormalEquations(n, K, table, stride, p.use_idx, fmask, data, s);
			SolveFullPivLU(K, s);
			if (spline) {
				SplineFullCoefficients(n, p.npiece, s);
			}
			// model, residual (all channels), statistics over final_mask
			int c = 0;
			double sum = 0.0, sq = 0.0;
			for (int i = tid; i < n; i += nt) {
				double m;
				if (spline) {
					m = SplineModel(i, n, p.npiece, s);
				} else {
					double const *trow = table + static_cast<size_t>(i) * stride;
					m = 0.0;
					for (int j = 0; j < K; ++j) {
						m = fma(s.cvec[j], trow[p.use_idx[j]], m);
					}
				}
				float const mf = static_cast<float>(m);
				float const r = __fsub_rn(data[i], mf);
				resid[i] = r;
				if (bestfit != nullptr) {
					bestfit[i] = mf;
				}
				if (fmask[i]) {
					double const rd = static_cast<double>(r);
					++c;
					sum += rd;
					sq = fma(rd, rd, sq);
				}
			}
			block_sum3(c, sum, sq, s.part);
			ClipThresholds const th = make_thresholds(c, sum, sq, p.clip_threshold_sigma);
			rms_d = th.rms;
			if (iter < p.num_fitting_max) {
				__syncthreads(); // residual writes visible block-wide
				int nclip = 0;
				for (int i = tid; i < n; i += nt) {
					// last channel of every piece is never examined (baseline.cc:1081-1096)
					bool skip = (i == n - 1);
					if (spline && !skip) {
						int const pc = PieceOf(i, p.npiece, s.bound);
						skip = (static_cast<unsigned long long>(i) + 1ull == s.bound[pc + 1]);
					}
					if (!skip && fmask[i]) {
						if (is_clipped(resid[i], th.lower, th.upper)) {
							fmask[i] = 0;
							++nclip;
						}
					}
				}
				int const num_clipped = block_sum_int(nclip, s.iscr);
				if (num_clipped == 0) {
					break;
				}
				num_unmasked = c - num_clipped;
				__syncthreads();
			}
		}
		if (failed) {
			if (tid == 0) {
				p.status[row] = kStatusNG;
				p.lsqfit_status[row] = kLsqNotEnoughData;
				if (p.flags != nullptr) {
					p.flags[row] = (spline ? 4 : 0) | 2; // final_mask already clipped in place
				}
			}
			continue;
		}
		// outputs (baseline.cc:1310-1328, 1389-1408)
		__syncthreads();
		if (p.residual != nullptr) {
			float *out = p.residual + row * p.data_stride;
			for (int i = tid; i < n; i += nt) {
				out[i] = resid[i];
			}
		}
		if (p.best_fit != nullptr) {
			float *out = p.best_fit + row * p.data_stride;
			for (int i = tid; i < n; i += nt) {
				out[i] = bestfit[i];
			}
		}
		if (p.coeff != nullptr) {
			double *co = p.coeff + row * p.coeff_stride;
			if (tid == 0) {
				double const max_x = static_cast<double>(n - 1);
				if (spline) {
					for (int i = 0; i < p.npiece; ++i) {
						double factor = 1.0;
						for (int j = 0; j < 4; ++j) {
							co[4 * i + j] = s.cfull[4 * i + j] / factor;
							factor *= max_x;
						}
					}
				} else if (p.fit_type == kFitPolynomial) {
					double factor = 1.0;
					for (int j = 0; j < K; ++j) {
						co[j] = s.cvec[j] / factor;
						factor *= max_x;
					}
				} else {
					for (int j = 0; j < K; ++j) {
						co[j] = s.cvec[j];
					}
				}
			}
		}
		if (tid == 0) {
			p.rms[row] = static_cast<float>(rms_d);
			p.status[row] = kStatusOK;
			p.lsqfit_status[row] = kLsqOK;
			if (p.flags != nullptr) {
				p.flags[row] = (spline ? 4 : 0) | 3;
			}
		}
	}
}

__global__ void __launch_bounds__(kGeneralThreads)
SubtractKernel(SubtractParams const p) {
	__shared__ double cf[4 * kMaxBases];
	__shared__ unsigned long long bound[kMaxBases + 2];
	int const n = p.n;
	bool const spline = (p.fit_type == kFitCubicSpline);
	for (size_t row = blockIdx.x; row < p.num_spectra; row += gridDim.x) {
		__syncthreads();
		int const ncoef = spline ? 4 * p.npiece : p.K;
		for (int i = threadIdx.x; i < ncoef; i += blockDim.x) {
			cf[i] = p.coeff[row * p.coeff_stride + i];
		}
		if (spline) {
			for (int i = threadIdx.x; i <= p.npiece; i += blockDim.x) {
				bound[i] = p.boundary[row * (p.npiece + 1) + i];
			}
		}
		__syncthreads();
		float const *data = p.data + row * p.data_stride;
		float *out = p.out + row * p.data_stride;
		for (int i = threadIdx.x; i < n; i += blockDim.x) {
			double m;
			if (spline) {
				double const x = static_cast<double>(i) / static_cast<double>(n - 1);
				double const x2 = x * x;
				double const x3 = x2 * x;
				int lo = 0, hi = p.npiece;
				while (hi - lo > 1) {
					int const mid = (lo + hi) >> 1;
					if (static_cast<unsigned long long>(i) >= bound[mid]) {
						lo = mid;
					} else {
						hi = mid;
					}
				}
				double const *c = cf + 4 * lo;
				m = ((c[0] * 1.0 + c[1] * x) + c[2] * x2) + c[3] * x3;
			} else {
				double const *trow = p.basis + static_cast<size_t>(i) * p.nb_ctx;
				m = 0.0;
				for (int j = 0; j < p.K; ++j) {
					m = fma(cf[j], trow[p.use_idx[j]], m);
				}
			}
			out[i] = __fsub_rn(data[i], static_cast<float>(m));
		}
	}
}

} // namespace

size_t GeneralFitSharedBytes(int K, int npiece, int threads) {
	size_t doubles = static_cast<size_t>(K) * K + 2 * static_cast<size_t>(K)
			+ 4 * static_cast<size_t>(npiece > 0 ? npiece : 1) + threads + (npiece + 2);
	size_t ints = 2 * static_cast<size_t>(K) + 64;
	return doubles * sizeof(double) + ints * sizeof(int) + 16;
}

cudaError_t LaunchGeneralFit(GeneralFitParams const &p, int grid, cudaStream_t stream) {
	size_t const smem = GeneralFitSharedBytes(p.K, p.npiece, kGeneralThreads);
	cudaError_t err = cudaFuncSetAttribute(GeneralFitKernel,
			cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem));
	if (err != cudaSuccess) {
		return err;
	}
	GeneralFitKernel<<<grid, kGeneralThreads, smem, stream>>>(p);
	return cudaGetLastError();
}

cudaError_t LaunchSubtract(SubtractParams const &p, int grid, cudaStream_t stream) {
	SubtractKernel<<<grid, kGeneralThreads, 0, stream>>>(p);
	return cudaGetLastError();
}

} // namespace sakura_b200
