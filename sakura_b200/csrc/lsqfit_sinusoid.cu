// Batched sinusoid baseline fit (sakura_LSQFitSinusoidFloat semantics, baseline.cc:1835-1893 ->
// LSQFit :1268 -> DoLSQFit :1115) for the high-basis-count configs (BASELINE C4: 8192 channels,
// wave numbers 0..20 = 41 bases).
//
// One 256-thread CTA fits a tile of 32 spectra together, so that the FP64 basis table is read
// once per tile instead of once per spectrum, and the two dense contractions run on the FP64
// tensor cores (mma.sync m8n8k4 = DMMA):
//   data moments   b[r][k]   = sum_i  (mask*y)[r][i] * table[i][k]      (rows x channels x bases)
//   model          m[r][i]   = sum_k  coeff[r][k]    * table[i][k]      (rows x bases x channels)
// The Gram matrix needs no per-channel outer products at all: with phi = {1, sin(w t), cos(w t)}
//   sin a sin b = (cos(a-b) - cos(a+b))/2,  cos a cos b = (cos(a-b) + cos(a+b))/2,
//   sin a cos b = (sin(a+b) + sin(a-b))/2
// every entry is a combination of the masked sums C_m = sum_valid cos(m t_i), S_m = sum_valid
// sin(m t_i), m <= 2*max wave. Those are (all-channel constants) - (sum over the flagged
// channels), and a clip pass only subtracts the clipped channels (downdate, like the reference's
// numeric_operation.cc:283-314). The K x K solve is an LDL^T per row in shared memory; a row whose
// pivots are not safely positive is handed to the general kernel (full-pivot rank-revealing LU).
//
// Thresholds, the float clip predicate, "last channel never clipped", early exit, not-enough-data
// and which outputs a failing row leaves untouched follow baseline.cc:1142-1224 exactly as in the
// other kernels; results are staged per tile and published only for rows that succeed.
#include "lsqfit_sinusoid.cuh"
#include "device_common.cuh"

#include <float.h>

namespace sakura_b200 {

namespace {

constexpr int RB = kSinRowsPerTile;
constexpr int NWARP = kSinThreads / 32;

enum RowState : int {
	kActive = 0, kConverged = 1, kFailedAfterClip = 2, kFailedBefore = 3, kFallback = 4, kNoRow = 5
};

__device__ __forceinline__ void dmma(double &c0, double &c1, double a, double b) {
	asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
			: "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}

__device__ __forceinline__ double fast_rcp_pos(double x) {
	double r;
	asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x));
	double e = fma(-x, r, 1.0);
	r = fma(r, e, r);
	e = fma(-x, r, 1.0);
	r = fma(r, e, r);
	return r;
}

struct Smem {
	double *coef; // [RB][Kp]
	double *bvec; // [RB][Kp]
	double *trig; // [RB][NEp]
	double *gw; // [NWARP][gw_stride]
	double *stat; // [NWARP][RB][3]
	double *rmsd; // [RB]
	float *lo; // [RB]
	float *hi; // [RB]
	int *state; // [RB]
	int *cnt; // [RB] valid channels (statistics count of the latest fit)
	int gw_stride;
	int NEp;
};

__host__ __device__ inline int GwStride(int Kp) {
	int const a = Kp * (Kp + 1);
	int const b = RB * Kp;
	return a > b ? a : b;
}

__host__ __device__ inline int PadNE(int next) {
	return (next + 3) / 4 * 4;
}

__device__ __forceinline__ Smem Carve(unsigned char *base, int Kp, int next) {
	Smem s;
	s.NEp = PadNE(next);
	s.gw_stride = GwStride(Kp);
	double *d = reinterpret_cast<double *>(base);
	s.coef = d;
	d += RB * Kp;
	s.bvec = d;
	d += RB * Kp;
	s.trig = d;
	d += RB * s.NEp;
	s.gw = d;
	d += NWARP * s.gw_stride;
	s.stat = d;
	d += NWARP * RB * 3;
	s.rmsd = d;
	d += RB;
	s.lo = reinterpret_cast<float *>(d);
	s.hi = s.lo + RB;
	s.state = reinterpret_cast<int *>(s.hi + RB);
	s.cnt = s.state + RB;
	return s;
}

// masked trigonometric sums needed by basis pair (j, k): see the header comment
__device__ __forceinline__ double GramEntry(double const *trig, int kj, int wj, int kk, int wk) {
	auto C = [&](int m) {return trig[m == 0 ? 0 : 2 * m];};
	auto S = [&](int m) {return m == 0 ? 0.0 : (m > 0 ? trig[2 * m - 1] : -trig[-2 * m - 1]);};
	if (kj == 0 && kk == 0) {
		return C(0);
	}
	if (kj == 0) {
		return kk == 1 ? S(wk) : C(wk);
	}
	if (kk == 0) {
		return kj == 1 ? S(wj) : C(wj);
	}
	int const d = wj > wk ? wj - wk : wk - wj;
	if (kj == 1 && kk == 1) {
		return 0.5 * (C(d) - C(wj + wk));
	}
	if (kj == 2 && kk == 2) {
		return 0.5 * (C(d) + C(wj + wk));
	}
	// one sine (wave a), one cosine (wave b): (S_{a+b} + S_{a-b}) / 2
	int const a = (kj == 1) ? wj : wk;
	int const b = (kj == 1) ? wk : wj;
	return 0.5 * (S(a + b) + S(a - b));
}

// LDL^T of the K x K Gram matrix of one row in this warp's workspace, then the solve.
// Returns false (warp-uniform) when a pivot is not safely positive.
__device__ bool WarpSolveRow(SinusoidFitParams const &p, double const *trig, double const *bvec,
		double *A, double *x) {
	int const lane = threadIdx.x & 31;
	int const K = p.K;
	int const Kp = p.Kp + 1; // leading dimension, odd number of doubles: no bank conflicts
	// lower triangle of G
	for (int e = lane; e < K * K; e += 32) {
		int const i = e / K;
		int const j = e - i * K;
		if (j <= i) {
			A[i * Kp + j] = GramEntry(trig, p.kind[i], p.wave[i], p.kind[j], p.wave[j]);
		}
	}
	__syncwarp();
	double maxd = 0.0;
	bool ok = true;
	for (int k = 0; k < K; ++k) {
		double const d = A[k * Kp + k];
		maxd = fmax(maxd, fabs(d));
		ok = ok && (d > maxd * (DBL_EPSILON * K));
		double const rcp = fast_rcp_pos(d);
		// A[i][j] -= A[i][k] * A[j][k] / d for k < j <= i; row i owned by one lane
		for (int i = k + 1 + lane; i < K; i += 32) {
			double const l = A[i * Kp + k] * rcp;
			for (int j = k + 1; j <= i; ++j) {
				A[i * Kp + j] = fma(-l, A[j * Kp + k], A[i * Kp + j]);
			}
		}
		__syncwarp();
		// scale the column afterwards (the update above needs the unscaled entries of other rows)
		for (int i = k + 1 + lane; i < K; i += 32) {
			A[i * Kp + k] *= rcp;
		}
		__syncwarp();
	}
	// L z = b (column oriented), z /= D, L^T x = z
	for (int i = lane; i < K; i += 32) {
		x[i] = bvec[i];
	}
	__syncwarp();
	for (int k = 0; k < K; ++k) {
		double const xk = x[k];
		for (int i = k + 1 + lane; i < K; i += 32) {
			x[i] = fma(-A[i * Kp + k], xk, x[i]);
		}
		__syncwarp();
	}
	for (int i = lane; i < K; i += 32) {
		x[i] = x[i] / A[i * Kp + i];
	}
	__syncwarp();
	for (int k = K - 1; k >= 0; --k) {
		double const xk = x[k];
		for (int i = lane; i < k; i += 32) {
			x[i] = fma(-A[k * Kp + i], xk, x[i]);
		}
		__syncwarp();
	}
	return ok;
}

template<int NT8>
__global__ void __launch_bounds__(kSinThreads, 1)
SinusoidFitKernel(SinusoidFitParams const p) {
	extern __shared__ __align__(16) unsigned char smem_raw[];
	Smem const s = Carve(smem_raw, p.Kp, p.next);
	int const tid = threadIdx.x;
	int const lane = tid & 31;
	int const warp = tid >> 5;
	int const n = p.n;
	int const K = p.K;
	int const Kp = p.Kp;
	int const next = p.next;
	size_t const num_tiles = (p.num_spectra + RB - 1) / RB;
	float *wres = p.ws_residual + static_cast<size_t>(blockIdx.x) * RB * n;
	float *wbest = p.ws_best_fit ? p.ws_best_fit + static_cast<size_t>(blockIdx.x) * RB * n : nullptr;
	uint8_t *wmask = p.ws_mask + static_cast<size_t>(blockIdx.x) * RB * n;
	float const clip = p.clip_threshold_sigma;

	for (size_t tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
		size_t const row0 = tile * RB;
		__syncthreads();
		// ---- P0: working mask, valid counts, masked trigonometric sums (one warp per row) -------
		for (int r = warp; r < RB; r += NWARP) {
			size_t const row = row0 + r;
			if (row >= p.num_spectra) {
				if (lane == 0) {
					s.state[r] = kNoRow;
				}
				continue;
			}
			uint8_t const *gm = p.mask + row * p.mask_stride;
			uint8_t *wm = wmask + static_cast<size_t>(r) * n;
			int cnt = 0;
			for (int i = lane * 8; i < n; i += 256) {
				uint2 v = *reinterpret_cast<uint2 const *>(gm + i);
				v.x = __vcmpne4(v.x, 0u) & 0x01010101u;
				v.y = __vcmpne4(v.y, 0u) & 0x01010101u;
				*reinterpret_cast<uint2 *>(wm + i) = v;
				cnt += __popc(v.x) + __popc(v.y);
			}
			cnt = warp_sum(cnt);
			__syncwarp();
			int st = kActive;
			if (p.num_fitting_max == 0) {
				st = kConverged; // baseline.cc:1295-1296: kOK, nothing written
			} else if (cnt < K) {
				st = kFailedBefore; // baseline.cc:1147-1150
			}
			if (lane == 0) {
				s.state[r] = st;
				s.cnt[r] = cnt;
			}
			if (st != kActive) {
				continue;
			}
			bool const complement = (2 * cnt >= n);
			double acc[3] = { 0.0, 0.0, 0.0 }; // next <= 96 columns: three per lane
			for (int base = 0; base < n; base += 32) {
				bool const valid = wm[base + lane] != 0;
				unsigned b = __ballot_sync(0xffffffffu, complement ? !valid : valid);
				while (b != 0u) {
					int const j = __ffs(b) - 1;
					b &= b - 1u;
					double const *er = p.ext + static_cast<size_t>(base + j) * next;
#pragma unroll
					for (int q = 0; q < 3; ++q) {
						int const e = lane + 32 * q;
						if (e < next) {
							acc[q] += er[e];
						}
					}
				}
			}
#pragma unroll
			for (int q = 0; q < 3; ++q) {
				int const e = lane + 32 * q;
				if (e < next) {
					s.trig[r * s.NEp + e] = complement ? (p.ext_all[e] - acc[q]) : acc[q];
				}
			}
		}
		__syncthreads();

		// ---- P1: data moments b = (mask*y) x table on the tensor cores ---------------------------
		{
			double c[4][NT8][2];
#pragma unroll
			for (int mt = 0; mt < 4; ++mt) {
#pragma unroll
				for (int nt = 0; nt < NT8; ++nt) {
					c[mt][nt][0] = 0.0;
					c[mt][nt][1] = 0.0;
				}
			}
			int const per_warp = n / NWARP; // n % 8 == 0 and NWARP == 8: a multiple of... see host check
			int const c_begin = warp * per_warp;
			int const arow = lane >> 2; // fragment row / column-of-B index
			int const acol = lane & 3;
			bool act1[4];
#pragma unroll
			for (int mt = 0; mt < 4; ++mt) {
				act1[mt] = (s.state[8 * mt + arow] == kActive);
			}
			for (int c0 = c_begin; c0 < c_begin + per_warp; c0 += 4) {
				int const ch = c0 + acol;
				double a[4];
#pragma unroll
				for (int mt = 0; mt < 4; ++mt) {
					int const r = 8 * mt + arow;
					size_t const row = row0 + r;
					double v = 0.0;
					if (act1[mt]) {
						float const y = p.data[row * p.data_stride + ch];
						uint8_t const m = wmask[static_cast<size_t>(r) * n + ch];
						v = static_cast<double>(__int_as_float(m ? __float_as_int(y) : 0));
					}
					a[mt] = v;
				}
				double const *trow = p.table + static_cast<size_t>(ch) * Kp + arow;
#pragma unroll
				for (int nt = 0; nt < NT8; ++nt) {
					double const b = trow[8 * nt];
#pragma unroll
					for (int mt = 0; mt < 4; ++mt) {
						dmma(c[mt][nt][0], c[mt][nt][1], a[mt], b);
					}
				}
			}
			double *mine = s.gw + warp * s.gw_stride; // [RB][Kp]
#pragma unroll
			for (int mt = 0; mt < 4; ++mt) {
#pragma unroll
				for (int nt = 0; nt < NT8; ++nt) {
					int const r = 8 * mt + (lane >> 2);
					int const k = 8 * nt + 2 * (lane & 3);
					mine[r * Kp + k] = c[mt][nt][0];
					mine[r * Kp + k + 1] = c[mt][nt][1];
				}
			}
			__syncthreads();
			for (int e = tid; e < RB * Kp; e += kSinThreads) {
				double tot = 0.0;
#pragma unroll
				for (int w = 0; w < NWARP; ++w) {
					tot += s.gw[w * s.gw_stride + e];
				}
				s.bvec[e] = tot;
			}
			__syncthreads();
		}

		// ---- clip iterations -------------------------------------------------------------------------
		for (int iter = 1; iter <= p.num_fitting_max; ++iter) {
			// (a) solve, one warp per row
			bool any_active = false;
			for (int r = warp; r < RB; r += NWARP) {
				if (s.state[r] != kActive) {
					continue;
				}
				any_active = true;
				bool const ok = WarpSolveRow(p, s.trig + r * s.NEp, s.bvec + r * Kp,
						s.gw + warp * s.gw_stride, s.coef + r * Kp);
				if (!ok) {
					if (lane == 0) {
						s.state[r] = kFallback;
						int const slot = atomicAdd(p.fallback_count, 1);
						p.fallback_rows[slot] = static_cast<int32_t>(row0 + r);
					}
				}
				for (int k = K + lane; k < Kp; k += 32) {
					s.coef[r * Kp + k] = 0.0;
				}
			}
			if (__syncthreads_or(any_active ? 1 : 0) == 0) {
				break;
			}
			// (b) model + residual + statistics: coeff x table on the tensor cores
			{
				int cnt[4] = { 0, 0, 0, 0 };
				double sum[4] = { 0.0, 0.0, 0.0, 0.0 };
				double sq[4] = { 0.0, 0.0, 0.0, 0.0 };
				int const frow = lane >> 2;
				int const fk = lane & 3;
				bool act[4];
#pragma unroll
				for (int mt = 0; mt < 4; ++mt) {
					act[mt] = (s.state[8 * mt + frow] == kActive);
				}
				for (int c0 = warp * 8; c0 < n; c0 += 8 * NWARP) {
					double c[4][2];
#pragma unroll
					for (int mt = 0; mt < 4; ++mt) {
						c[mt][0] = 0.0;
						c[mt][1] = 0.0;
					}
					double const *tb = p.table + static_cast<size_t>(c0 + frow) * Kp + fk;
					for (int ks = 0; ks < Kp; ks += 4) {
						double const b = tb[ks];
#pragma unroll
						for (int mt = 0; mt < 4; ++mt) {
							double const a = s.coef[(8 * mt + frow) * Kp + ks + fk];
							dmma(c[mt][0], c[mt][1], a, b);
						}
					}
					int const ch = c0 + 2 * fk;
#pragma unroll
					for (int mt = 0; mt < 4; ++mt) {
						if (act[mt]) {
							int const r = 8 * mt + frow;
							size_t const row = row0 + r;
							float2 const y = *reinterpret_cast<float2 const *>(p.data + row * p.data_stride + ch);
							uchar2 const m = *reinterpret_cast<uchar2 const *>(wmask + static_cast<size_t>(r) * n + ch);
							float const m0 = static_cast<float>(c[mt][0]);
							float const m1 = static_cast<float>(c[mt][1]);
							float const r0 = __fsub_rn(y.x, m0);
							float const r1 = __fsub_rn(y.y, m1);
							*reinterpret_cast<float2 *>(wres + static_cast<size_t>(r) * n + ch) = make_float2(r0, r1);
							if (wbest != nullptr) {
								*reinterpret_cast<float2 *>(wbest + static_cast<size_t>(r) * n + ch) = make_float2(m0, m1);
							}
							if (m.x) {
								double const rd = static_cast<double>(r0);
								++cnt[mt];
								sum[mt] += rd;
								sq[mt] = fma(rd, rd, sq[mt]);
							}
							if (m.y) {
								double const rd = static_cast<double>(r1);
								++cnt[mt];
								sum[mt] += rd;
								sq[mt] = fma(rd, rd, sq[mt]);
							}
						}
					}
				}
#pragma unroll
				for (int mt = 0; mt < 4; ++mt) {
#pragma unroll
					for (int o = 1; o <= 2; o <<= 1) {
						cnt[mt] += __shfl_xor_sync(0xffffffffu, cnt[mt], o);
						sum[mt] += __shfl_xor_sync(0xffffffffu, sum[mt], o);
						sq[mt] += __shfl_xor_sync(0xffffffffu, sq[mt], o);
					}
					if (fk == 0) {
						double *st = s.stat + (warp * RB + 8 * mt + frow) * 3;
						st[0] = static_cast<double>(cnt[mt]);
						st[1] = sum[mt];
						st[2] = sq[mt];
					}
				}
			}
			__syncthreads();
			if (tid < RB && s.state[tid] == kActive) {
				double cd = 0.0, sum = 0.0, sq = 0.0;
				for (int w = 0; w < NWARP; ++w) {
					double const *st = s.stat + (w * RB + tid) * 3;
					cd += st[0];
					sum += st[1];
					sq += st[2];
				}
				int const c = static_cast<int>(cd);
				ClipThresholds const th = make_thresholds(c, sum, sq, clip);
				s.rmsd[tid] = th.rms;
				s.lo[tid] = th.lower;
				s.hi[tid] = th.upper;
				s.cnt[tid] = c;
			}
			__syncthreads();
			if (iter == p.num_fitting_max) {
				break;
			}
			// (c) clip + downdate, one warp per row, channels in ascending order
			for (int r = warp; r < RB; r += NWARP) {
				if (s.state[r] != kActive) {
					continue;
				}
				size_t const row = row0 + r;
				float const lo = s.lo[r];
				float const hi = s.hi[r];
				uint8_t *wm = wmask + static_cast<size_t>(r) * n;
				float const *rr = wres + static_cast<size_t>(r) * n;
				float const *yy = p.data + row * p.data_stride;
				double *trig = s.trig + r * s.NEp;
				double *bv = s.bvec + r * Kp;
				int num_clipped = 0;
				for (int base = 0; base < n; base += 32) {
					int const i = base + lane;
					bool clipped = false;
					if (i < n - 1 && wm[i] != 0) { // the last channel is never examined
						clipped = is_clipped(rr[i], lo, hi);
					}
					unsigned b = __ballot_sync(0xffffffffu, clipped);
					if (clipped) {
						wm[i] = 0;
					}
					num_clipped += __popc(b);
					while (b != 0u) {
						int const j = __ffs(b) - 1;
						b &= b - 1u;
						int const ch = base + j;
						double const *er = p.ext + static_cast<size_t>(ch) * next;
						for (int e = lane; e < next; e += 32) {
							trig[e] -= er[e];
						}
						double const yd = static_cast<double>(yy[ch]);
						double const *tr = p.table + static_cast<size_t>(ch) * Kp;
						for (int k = lane; k < K; k += 32) {
							bv[k] = fma(-yd, tr[k], bv[k]);
						}
					}
				}
				__syncwarp();
				if (lane == 0) {
					if (num_clipped == 0) {
						s.state[r] = kConverged; // baseline.cc:1211-1213
					} else if (s.cnt[r] - num_clipped < K) {
						s.state[r] = kFailedAfterClip; // baseline.cc:1162-1165 at the next iteration
					}
				}
			}
			__syncthreads();
		}
		__syncthreads();

		// ---- outputs ---------------------------------------------------------------------------------
		for (int r = warp; r < RB; r += NWARP) {
			int const st = s.state[r];
			size_t const row = row0 + r;
			if (st == kNoRow || st == kFallback) {
				continue;
			}
			if (st == kFailedBefore) {
				if (lane == 0) {
					p.status[row] = kStatusNG;
					p.lsqfit_status[row] = kLsqNotEnoughData;
				}
				continue;
			}
			if (p.num_fitting_max == 0) {
				if (lane == 0) {
					p.status[row] = kStatusOK;
					p.lsqfit_status[row] = kLsqOK;
				}
				continue;
			}
			uint8_t const *wm = wmask + static_cast<size_t>(r) * n;
			uint8_t *fm = p.final_mask + row * p.mask_stride;
			for (int i = lane * 8; i < n; i += 256) {
				*reinterpret_cast<uint2 *>(fm + i) = *reinterpret_cast<uint2 const *>(wm + i);
			}
			if (st == kFailedAfterClip) {
				if (lane == 0) {
					p.status[row] = kStatusNG;
					p.lsqfit_status[row] = kLsqNotEnoughData;
					if (p.flags != nullptr) {
						p.flags[row] = 2;
					}
				}
				continue;
			}
			if (p.residual != nullptr) {
				float4 const *src = reinterpret_cast<float4 const *>(wres + static_cast<size_t>(r) * n);
				float4 *dst = reinterpret_cast<float4 *>(p.residual + row * p.data_stride);
				for (int i = lane; i < n / 4; i += 32) {
					dst[i] = src[i];
				}
			}
			if (p.best_fit != nullptr) {
				float4 const *src = reinterpret_cast<float4 const *>(wbest + static_cast<size_t>(r) * n);
				float4 *dst = reinterpret_cast<float4 *>(p.best_fit + row * p.data_stride);
				for (int i = lane; i < n / 4; i += 32) {
					dst[i] = src[i];
				}
			}
			if (p.coeff != nullptr) {
				for (int k = lane; k < K; k += 32) {
					p.coeff[row * K + k] = s.coef[r * Kp + k]; // no rescale for sinusoids
				}
			}
			if (lane == 0) {
				p.rms[row] = static_cast<float>(s.rmsd[r]);
				p.status[row] = kStatusOK;
				p.lsqfit_status[row] = kLsqOK;
				if (p.flags != nullptr) {
					p.flags[row] = 3;
				}
			}
		}
	}
}

__global__ void GatherTableKernel(double const *ctx_table, int nb_ctx, int n, int K, int Kp,
		int16_t const *use_idx, double *table) {
	size_t const total = static_cast<size_t>(n) * Kp;
	for (size_t e = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x; e < total;
			e += static_cast<size_t>(gridDim.x) * blockDim.x) {
		size_t const i = e / Kp;
		int const k = static_cast<int>(e - i * Kp);
		table[e] = (k < K) ? ctx_table[i * nb_ctx + use_idx[k]] : 0.0;
	}
}

} // namespace

size_t SinusoidSharedBytes(int Kp, int next) {
	size_t doubles = static_cast<size_t>(2 * RB * Kp) + static_cast<size_t>(RB) * PadNE(next)
			+ static_cast<size_t>(NWARP) * GwStride(Kp) + static_cast<size_t>(NWARP) * RB * 3 + RB;
	return doubles * sizeof(double) + RB * (2 * sizeof(float) + 2 * sizeof(int)) + 32;
}

cudaError_t LaunchGatherTable(double const *ctx_table, int nb_ctx, int n, int K, int Kp,
		int16_t const *use_idx_dev, double *table, cudaStream_t stream) {
	GatherTableKernel<<<256, 256, 0, stream>>>(ctx_table, nb_ctx, n, K, Kp, use_idx_dev, table);
	return cudaGetLastError();
}

template<int NT8>
static cudaError_t LaunchNT(SinusoidFitParams const &p, int grid, cudaStream_t stream) {
	size_t const smem = SinusoidSharedBytes(p.Kp, p.next);
	cudaError_t err = cudaFuncSetAttribute(SinusoidFitKernel<NT8>,
			cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem));
	if (err != cudaSuccess) {
		return err;
	}
	SinusoidFitKernel<NT8><<<grid, kSinThreads, smem, stream>>>(p);
	return cudaGetLastError();
}

cudaError_t LaunchSinusoidFit(SinusoidFitParams const &p, int grid, cudaStream_t stream) {
	switch (p.Kp / 8) {
	case 1:
		return LaunchNT<1>(p, grid, stream);
	case 2:
		return LaunchNT<2>(p, grid, stream);
	case 3:
		return LaunchNT<3>(p, grid, stream);
	case 4:
		return LaunchNT<4>(p, grid, stream);
	case 5:
		return LaunchNT<5>(p, grid, stream);
	case 6:
		return LaunchNT<6>(p, grid, stream);
	default:
		return cudaErrorInvalidValue;
	}
}

} // namespace sakura_b200
