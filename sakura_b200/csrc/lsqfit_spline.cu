// Batched cubic-spline baseline fit (sakura_LSQFitCubicSplineFloat semantics, baseline.cc:1761-1833
// -> LSQFitCubicSpline :1342 -> DoLSQFit :1115), one 256-thread CTA per spectrum, the spectrum and
// its working mask resident in shared memory for the whole fit.
//
// The reference builds a per-spectrum basis table [n][npiece+3] (baseline.cc:911-980: 1, x, x^2,
// x^3 - (x-b_1)^3_+, then (x-b_{j-1})^3_+ - (x-b_j)^3_+) and accumulates all (npiece+3)^2 products
// per channel. Every one of those bases is a cubic polynomial inside a piece, phi_k(x) =
// sum_a M_q[k][a] u^a with u = x - x(b_q) the piece-local coordinate, so the normal equations are
//   G = sum_q M_q Hank(h_q) M_q^T,   b = sum_q M_q t_q,
//   h_q[m] = sum_{valid i in piece q} u_i^m (m = 0..6),   t_q[a] = sum_valid y_i u_i^a (a = 0..3):
// 11 moments per channel instead of 552 products, and a clip pass only subtracts the clipped
// channels from the moments of their piece (downdate, as numeric_operation.cc:283-314 does for
// the matrix). All expansions are about the piece origin with non-negative shifts, so nothing
// cancels. The solve is the shared blocked LDL^T; a row whose pivots are not safely positive is
// handed to the general kernel (full-pivot LU restating Eigen's). The model is evaluated in the
// piece-local form; the per-piece global cubic coefficients the API returns are derived from the
// solution exactly as baseline.cc:868-893 does.
//
// Boundaries (baseline.cc:808-844), "the last channel of each piece is never examined"
// (:1081-1096), thresholds, early exit, not-enough-data and which outputs a failing row leaves
// untouched follow the general kernel (lsqfit_general.cu), which remains the restatement of the
// reference's arithmetic order.
#include "lsqfit_spline.cuh"
#include "device_common.cuh"
#include "ldlt_device.cuh"

namespace sakura_b200 {

namespace {

constexpr int NTHR = kSplineThreads;
constexpr int NW = NTHR / 32;
constexpr unsigned kFull = 0xffffffffu;

// exact for 0 <= v < 2^31, without the (slow) conversion pipe
__device__ __forceinline__ double int_to_double(int v) {
	return __hiloint2double(0x43300000, v) - 4503599627370496.0;
}

__host__ __device__ inline int PaddedBases(int npiece) {
	return (npiece + 3 + 7) / 8 * 8;
}

struct Smem {
	double *M; // [npiece][Kp][4] local cubic of every basis in every piece
	double *W; // [npiece][Kp][4] M_q Hank(h_q)
	double *H; // [npiece][8] local power sums (7 used)
	double *T; // [npiece][4] local data moments
	double *loc; // [npiece][4] local cubic of the fitted model
	double *A; // bordered normal equations
	double *cvec; // [Kp] solution
	double *cfull; // [npiece][4] global cubic per piece
	double *part; // [3 * NW]
	float *y; // [n]
	float *r; // [n]
	int *bound; // [npiece + 1]
	int *iscr; // [NW + 4]
	uint8_t *wm; // [n] working mask, 0/1
};

__host__ __device__ inline size_t SmemDoubles(int npiece) {
	int const Kp = PaddedBases(npiece);
	return 2 * static_cast<size_t>(npiece) * Kp * 4
			+ static_cast<size_t>(npiece) * (8 + 4 + 4 + 4) + LdltDoubles(Kp) + Kp + 3 * NW;
}

__device__ __forceinline__ Smem Carve(unsigned char *base, int n, int npiece) {
	int const Kp = PaddedBases(npiece);
	Smem s;
	double *d = reinterpret_cast<double *>(base);
	s.M = d;
	d += npiece * Kp * 4;
	s.W = d;
	d += npiece * Kp * 4;
	s.H = d;
	d += npiece * 8;
	s.T = d;
	d += npiece * 4;
	s.loc = d;
	d += npiece * 4;
	s.cfull = d;
	d += npiece * 4;
	s.A = d;
	d += LdltDoubles(Kp);
	s.cvec = d;
	d += Kp;
	s.part = d;
	d += 3 * NW;
	s.y = reinterpret_cast<float *>(d);
	s.r = s.y + n;
	s.bound = reinterpret_cast<int *>(s.r + n);
	s.iscr = s.bound + (npiece + 1 + 3) / 4 * 4;
	s.wm = reinterpret_cast<uint8_t *>(s.iscr + NW + 4);
	return s;
}

// One warp walks the channels [b0, b1) of a piece: the 4-aligned interior four channels per lane
// (128-bit shared-memory accesses, vec(i) with i % 4 == 0), the <= 3 + 3 channels at the two ends
// one per lane (one(i)). The order in which channels are visited does not matter to the callers.
template<class Vec, class One>
__device__ __forceinline__ void ForPiece(int b0, int b1, int lane, Vec vec, One one) {
	int const a0 = min((b0 + 3) & ~3, b1);
	int const a1 = max(b1 & ~3, a0);
	if (lane < a0 - b0) {
		one(b0 + lane);
	} else if (lane >= 16 && lane - 16 < b1 - a1) {
		one(a1 + lane - 16);
	}
	for (int i = a0 + 4 * lane; i < a1; i += 128) {
		vec(i);
	}
}

// local power sums h[m] += u^m (m <= 6) and data moments t[a] += y u^a (a <= 3) of one channel
__device__ __forceinline__ void AddMoments(double (&h)[7], double (&t)[4], double u, double y) {
	double const u2 = u * u;
	double const u3 = u2 * u;
	h[0] += 1.0;
	h[1] += u;
	h[2] += u2;
	h[3] += u3;
	h[4] = fma(u2, u2, h[4]);
	h[5] = fma(u2, u3, h[5]);
	h[6] = fma(u3, u3, h[6]);
	t[0] += y;
	t[1] = fma(y, u, t[1]);
	t[2] = fma(y, u2, t[2]);
	t[3] = fma(y, u3, t[3]);
}

// baseline.cc:808-844: boundary[idx] (1 <= idx < npiece) is the unmasked channel whose rank among
// the unmasked channels equals ceil(N*idx/npiece).
__device__ void Boundaries(int n, int npiece, int num_valid, Smem const &s, int tid) {
	int const lane = tid & 31;
	int const warp = tid >> 5;
	int const chunk = ((n / 4 + NTHR - 1) / NTHR) * 4;
	int const lo = min(n, tid * chunk);
	int const hi = min(n, lo + chunk);
	int cnt = 0;
	for (int i = lo; i < hi; i += 4) {
		cnt += __popc(*reinterpret_cast<unsigned const *>(s.wm + i));
	}
	int incl = cnt;
#pragma unroll
	for (int o = 1; o < 32; o <<= 1) {
		int const v = __shfl_up_sync(kFull, incl, o);
		incl += (lane >= o) ? v : 0;
	}
	__syncthreads(); // iscr may still be read from a previous use
	if (lane == 31) {
		s.iscr[warp] = incl;
	}
	__syncthreads();
	int start = incl - cnt;
	for (int w = 0; w < warp; ++w) {
		start += s.iscr[w];
	}
	if (tid == 0) {
		s.bound[0] = 0;
		s.bound[npiece] = n;
	}
	// rank targets once per row (64-bit divisions), then every thread only compares
	int *tgt = reinterpret_cast<int *>(s.loc);
	if (tid >= 1 && tid < npiece) {
		long long const N = num_valid;
		tgt[tid] = static_cast<int>((N * tid + npiece - 1) / npiece);
	}
	__syncthreads();
	for (int idx = 1; idx < npiece; ++idx) {
		int const target = tgt[idx];
		if (target >= start && target < start + cnt) {
			int rem = target - start;
			for (int i = lo; i < hi; i += 4) {
				unsigned const w = *reinterpret_cast<unsigned const *>(s.wm + i);
				int const c = __popc(w);
				if (rem < c) {
					int at = i;
					unsigned v = w;
					while (true) {
						if (v & 0xffu) {
							if (rem == 0) {
								break;
							}
							--rem;
						}
						v >>= 8;
						++at;
					}
					s.bound[idx] = at;
					break;
				}
				rem -= c;
			}
		}
	}
	__syncthreads();
}

// baseline.cc:868-893 as compiled (fused where GCC fuses): global cubic coefficients per piece
__device__ void FullCoefficients(int n, int npiece, Smem const &s) {
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

template<int NT8>
__global__ void __launch_bounds__(NTHR, 2)
SplineFitKernel(SplineFitParams const p) {
	constexpr int Kp = 8 * NT8;
	constexpr int LD = LdltLd(Kp);
	constexpr int kTri = Kp * (Kp + 1) / 2;
	extern __shared__ __align__(16) unsigned char smem_raw[];
	int const n = p.n;
	int const npiece = p.npiece;
	int const K = npiece + 3;
	Smem const s = Carve(smem_raw, n, npiece);
	int const tid = threadIdx.x;
	int const lane = tid & 31;
	int const warp = tid >> 5;
	double const max_x = static_cast<double>(n - 1);
	double const inv = 1.0 / max_x;

	for (size_t row = blockIdx.x; row < p.num_spectra; row += gridDim.x) {
		float const *data = p.data + row * p.data_stride;
		uint8_t const *mask = p.mask + row * p.mask_stride;
		__syncthreads(); // previous row's shared state fully consumed
		int cnt = 0;
		for (int i = tid * 4; i < n; i += NTHR * 4) {
			unsigned v = *reinterpret_cast<unsigned const *>(mask + i);
			v = __vcmpne4(v, 0u) & 0x01010101u;
			*reinterpret_cast<unsigned *>(s.wm + i) = v;
			cnt += __popc(v);
			*reinterpret_cast<float4 *>(s.y + i) = *reinterpret_cast<float4 const *>(data + i);
		}
		int const num_valid = block_sum_int(cnt, s.iscr);
		if (num_valid < npiece) { // baseline.cc:822-824 -> Status_kNG, lsqfit_status stays kNG
			if (tid == 0) {
				p.status[row] = kStatusNG;
				p.lsqfit_status[row] = kLsqNG;
			}
			continue;
		}
		Boundaries(n, npiece, num_valid, s, tid);
		if (p.boundary != nullptr) {
			for (int i = tid; i <= npiece; i += NTHR) {
				p.boundary[row * (npiece + 1) + i] = static_cast<unsigned long long>(s.bound[i]);
			}
		}
		if (tid == 0 && p.flags != nullptr) {
			p.flags[row] = 4;
		}
		if (p.num_fitting_max == 0) { // baseline.cc:1371-1372
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
		// ---- piece tables ----------------------------------------------------------------------------
		for (int e = tid; e < npiece * Kp; e += NTHR) {
			int const q = e / Kp;
			int const k = e - q * Kp;
			double m0 = 0.0, m1 = 0.0, m2 = 0.0, m3 = 0.0;
			int const bq = s.bound[q];
			if (k < 3) {
				double const xq = int_to_double(bq) / max_x;
				if (k == 0) {
					m0 = 1.0;
				} else if (k == 1) {
					m0 = xq;
					m1 = 1.0;
				} else {
					m0 = xq * xq;
					m1 = 2.0 * xq;
					m2 = 1.0;
				}
			} else if (k < K) {
				// basis k = (x - b_{j-1})^3_+ - (x - b_j)^3_+, j = k - 2 (b_0 = 0; the last one has b_j = n)
				int const j = k - 2;
				if (q == j - 1) {
					m3 = 1.0;
				} else if (q >= j) {
					int const bl = s.bound[j - 1];
					int const bh = s.bound[j];
					double const f = int_to_double(bq - bl) / max_x;
					double const g = int_to_double(bq - bh) / max_x;
					double const d = int_to_double(bh - bl) / max_x;
					m0 = d * fma(f, f, fma(f, g, g * g));
					m1 = 3.0 * d * (f + g);
					m2 = 3.0 * d;
				}
			}
			double *m = s.M + static_cast<size_t>(e) * 4;
			m[0] = m0;
			m[1] = m1;
			m[2] = m2;
			m[3] = m3;
		}
		// ---- local moments, one warp per piece ---------------------------------------------------------
		for (int q = warp; q < npiece; q += NW) {
			int const b0 = s.bound[q];
			int const b1 = s.bound[q + 1];
			double h[7] = { 0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0 };
			double t[4] = { 0.0, 0.0, 0.0, 0.0 };
			auto one = [&](int i, float yf, bool valid) {
				if (valid) {
					AddMoments(h, t, int_to_double(i - b0) * inv, static_cast<double>(yf));
				}
			};
			ForPiece(b0, b1, lane, [&](int i) {
				float4 const y = *reinterpret_cast<float4 const *>(s.y + i);
				unsigned const m = *reinterpret_cast<unsigned const *>(s.wm + i);
				one(i, y.x, (m & 0xffu) != 0u);
				one(i + 1, y.y, (m & 0xff00u) != 0u);
				one(i + 2, y.z, (m & 0xff0000u) != 0u);
				one(i + 3, y.w, (m & 0xff000000u) != 0u);
			}, [&](int i) {
				one(i, s.y[i], s.wm[i] != 0);
			});
#pragma unroll
			for (int m = 0; m < 7; ++m) {
				h[m] = warp_sum(h[m]);
			}
#pragma unroll
			for (int a = 0; a < 4; ++a) {
				t[a] = warp_sum(t[a]);
			}
			if (lane == 0) {
#pragma unroll
				for (int m = 0; m < 7; ++m) {
					s.H[q * 8 + m] = h[m];
				}
#pragma unroll
				for (int a = 0; a < 4; ++a) {
					s.T[q * 4 + a] = t[a];
				}
			}
		}
		__syncthreads();

		int num_unmasked = n; // baseline.cc:1152 (sic)
		double rms_d = 0.0;
		bool failed = false;
		bool fallback = false;
		for (int iter = 1; iter <= p.num_fitting_max; ++iter) {
			if (num_unmasked < K) { // baseline.cc:1162-1165
				failed = true;
				break;
			}
			// ---- normal equations from the moments --------------------------------------------------
			for (int e = tid; e < npiece * Kp; e += NTHR) {
				int const q = e / Kp;
				double const *m = s.M + static_cast<size_t>(e) * 4;
				double const *h = s.H + q * 8;
				double *w = s.W + static_cast<size_t>(e) * 4;
#pragma unroll
				for (int b = 0; b < 4; ++b) {
					w[b] = fma(m[3], h[b + 3], fma(m[2], h[b + 2], fma(m[1], h[b + 1], m[0] * h[b])));
				}
			}
			__syncthreads();
			for (int e = tid; e < kTri + Kp; e += NTHR) {
				if (e < kTri) {
					int k = static_cast<int>((sqrtf(8.0f * e + 1.0f) - 1.0f) * 0.5f);
					while ((k + 1) * (k + 2) / 2 <= e) {
						++k;
					}
					while (k * (k + 1) / 2 > e) {
						--k;
					}
					int const l = e - k * (k + 1) / 2; // l <= k
					double v = 0.0;
					if (k >= K) {
						v = (k == l) ? 1.0 : 0.0;
					} else {
						for (int q = max(0, k - 3); q < npiece; ++q) {
							double const *w = s.W + static_cast<size_t>(q * Kp + k) * 4;
							double const *m = s.M + static_cast<size_t>(q * Kp + l) * 4;
							v = fma(w[0], m[0], v);
							v = fma(w[1], m[1], v);
							v = fma(w[2], m[2], v);
							v = fma(w[3], m[3], v);
						}
					}
					s.A[k * LD + l] = v;
				} else {
					int const k = e - kTri;
					double v = 0.0;
					if (k < K) {
						for (int q = max(0, k - 3); q < npiece; ++q) {
							double const *m = s.M + static_cast<size_t>(q * Kp + k) * 4;
							double const *t = s.T + q * 4;
							v = fma(m[0], t[0], v);
							v = fma(m[1], t[1], v);
							v = fma(m[2], t[2], v);
							v = fma(m[3], t[3], v);
						}
					}
					s.A[Kp * LD + k] = v;
				}
			}
			__syncthreads();
			if (warp == 0) {
				bool const ok = WarpLdltSolve<NT8>(s.A, K, s.cvec, lane);
				if (lane == 0) {
					s.iscr[NW] = ok ? 1 : 0;
				}
			}
			__syncthreads();
			if (s.iscr[NW] == 0) {
				fallback = true;
				break;
			}
			for (int e = tid; e < npiece * 4; e += NTHR) {
				int const q = e >> 2;
				int const b = e & 3;
				double v = 0.0;
				for (int k = 0; k < K; ++k) {
					v = fma(s.cvec[k], s.M[static_cast<size_t>(q * Kp + k) * 4 + b], v);
				}
				s.loc[e] = v;
			}
			__syncthreads();
			// ---- model, residual, statistics: one warp per piece -------------------------------------
			int c = 0;
			double sum = 0.0, sq = 0.0;
			for (int q = warp; q < npiece; q += NW) {
				int const b0 = s.bound[q];
				int const b1 = s.bound[q + 1];
				double const l0 = s.loc[4 * q], l1 = s.loc[4 * q + 1], l2 = s.loc[4 * q + 2], l3 = s.loc[4 * q + 3];
				auto one = [&](int i, float yf, bool valid) {
					double const u = int_to_double(i - b0) * inv;
					double const m = fma(fma(fma(l3, u, l2), u, l1), u, l0);
					float const rr = __fsub_rn(yf, static_cast<float>(m));
					// flagged channels contribute +0.0 (a select: NaN/Inf under the mask must not leak)
					double const rd = static_cast<double>(__int_as_float(valid ? __float_as_int(rr) : 0));
					c += valid ? 1 : 0;
					sum += rd;
					sq = fma(rd, rd, sq);
					return rr;
				};
				ForPiece(b0, b1, lane, [&](int i) {
					float4 const y = *reinterpret_cast<float4 const *>(s.y + i);
					unsigned const m = *reinterpret_cast<unsigned const *>(s.wm + i);
					float4 rr;
					rr.x = one(i, y.x, (m & 0xffu) != 0u);
					rr.y = one(i + 1, y.y, (m & 0xff00u) != 0u);
					rr.z = one(i + 2, y.z, (m & 0xff0000u) != 0u);
					rr.w = one(i + 3, y.w, (m & 0xff000000u) != 0u);
					*reinterpret_cast<float4 *>(s.r + i) = rr;
				}, [&](int i) {
					s.r[i] = one(i, s.y[i], s.wm[i] != 0);
				});
			}
			block_sum3(c, sum, sq, s.part);
			ClipThresholds const th = make_thresholds(c, sum, sq, p.clip_threshold_sigma);
			rms_d = th.rms;
			if (iter == p.num_fitting_max) {
				break;
			}
			// ---- clip (not the last channel of a piece, baseline.cc:1081-1096) + moment downdate ----
			int nclip = 0;
			for (int q = warp; q < npiece; q += NW) {
				int const b0 = s.bound[q];
				int const last = s.bound[q + 1] - 1;
				double h[7] = { 0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0 };
				double t[4] = { 0.0, 0.0, 0.0, 0.0 };
				bool mine = false;
				auto hit = [&](int i) { // channel i is clipped: unflag it, remember its moments
					s.wm[i] = 0;
					++nclip;
					mine = true;
					AddMoments(h, t, int_to_double(i - b0) * inv, static_cast<double>(s.y[i]));
				};
				ForPiece(b0, last, lane, [&](int i) {
					float4 const r = *reinterpret_cast<float4 const *>(s.r + i);
					unsigned const m = *reinterpret_cast<unsigned const *>(s.wm + i);
					unsigned bits = 0u;
					bits |= ((m & 0xffu) != 0u && is_clipped(r.x, th.lower, th.upper)) ? 1u : 0u;
					bits |= ((m & 0xff00u) != 0u && is_clipped(r.y, th.lower, th.upper)) ? 2u : 0u;
					bits |= ((m & 0xff0000u) != 0u && is_clipped(r.z, th.lower, th.upper)) ? 4u : 0u;
					bits |= ((m & 0xff000000u) != 0u && is_clipped(r.w, th.lower, th.upper)) ? 8u : 0u;
					while (bits != 0u) { // rare
						hit(i + __ffs(bits) - 1);
						bits &= bits - 1u;
					}
				}, [&](int i) {
					if (s.wm[i] && is_clipped(s.r[i], th.lower, th.upper)) {
						hit(i);
					}
				});
				if (__any_sync(kFull, mine)) {
#pragma unroll
					for (int m = 0; m < 7; ++m) {
						h[m] = warp_sum(h[m]);
					}
#pragma unroll
					for (int a = 0; a < 4; ++a) {
						t[a] = warp_sum(t[a]);
					}
					if (lane == 0) {
#pragma unroll
						for (int m = 0; m < 7; ++m) {
							s.H[q * 8 + m] -= h[m];
						}
#pragma unroll
						for (int a = 0; a < 4; ++a) {
							s.T[q * 4 + a] -= t[a];
						}
					}
				}
			}
			int const num_clipped = block_sum_int(nclip, s.iscr);
			if (num_clipped == 0) { // baseline.cc:1211-1213
				break;
			}
			num_unmasked = c - num_clipped;
			__syncthreads();
		}
		if (fallback) {
			if (tid == 0) {
				int const slot = atomicAdd(p.fallback_count, 1);
				p.fallback_rows[slot] = static_cast<int32_t>(row);
			}
			continue;
		}
		__syncthreads();
		uint8_t *fmask = p.final_mask + row * p.mask_stride;
		for (int i = tid * 4; i < n; i += NTHR * 4) {
			*reinterpret_cast<unsigned *>(fmask + i) = *reinterpret_cast<unsigned const *>(s.wm + i);
		}
		if (failed) {
			if (tid == 0) {
				p.status[row] = kStatusNG;
				p.lsqfit_status[row] = kLsqNotEnoughData;
				if (p.flags != nullptr) {
					p.flags[row] = 4 | 2;
				}
			}
			continue;
		}
		// ---- outputs (baseline.cc:1389-1408) ---------------------------------------------------------
		if (p.residual != nullptr) {
			float *out = p.residual + row * p.data_stride;
			for (int i = tid * 4; i < n; i += NTHR * 4) {
				*reinterpret_cast<float4 *>(out + i) = *reinterpret_cast<float4 const *>(s.r + i);
			}
		}
		if (p.best_fit != nullptr) {
			float *out = p.best_fit + row * p.data_stride;
			for (int q = warp; q < npiece; q += NW) {
				int const b0 = s.bound[q];
				int const b1 = s.bound[q + 1];
				double const l0 = s.loc[4 * q], l1 = s.loc[4 * q + 1], l2 = s.loc[4 * q + 2], l3 = s.loc[4 * q + 3];
				for (int i = b0 + lane; i < b1; i += 32) {
					double const u = int_to_double(i - b0) * inv;
					out[i] = static_cast<float>(fma(fma(fma(l3, u, l2), u, l1), u, l0));
				}
			}
		}
		if (p.coeff != nullptr) {
			if (tid == 0) {
				FullCoefficients(n, npiece, s);
			}
			__syncthreads();
			double *co = p.coeff + row * p.coeff_stride;
			for (int e = tid; e < 4 * npiece; e += NTHR) {
				int const j = e & 3;
				double factor = 1.0;
				for (int t = 0; t < j; ++t) {
					factor *= max_x;
				}
				co[e] = s.cfull[e] / factor;
			}
		}
		if (tid == 0) {
			p.rms[row] = static_cast<float>(rms_d);
			p.status[row] = kStatusOK;
			p.lsqfit_status[row] = kLsqOK;
			if (p.flags != nullptr) {
				p.flags[row] = 4 | 3;
			}
		}
	}
}

template<int NT8>
cudaError_t LaunchNT(SplineFitParams const &p, int grid, cudaStream_t stream) {
	size_t const smem = SplineSharedBytes(p.n, p.npiece);
	cudaError_t err = cudaFuncSetAttribute(SplineFitKernel<NT8>,
			cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem));
	if (err != cudaSuccess) {
		return err;
	}
	SplineFitKernel<NT8><<<grid, NTHR, smem, stream>>>(p);
	return cudaGetLastError();
}

template<int NT8>
int ResidentNT(size_t smem) {
	int blocks = 0;
	if (cudaFuncSetAttribute(SplineFitKernel<NT8>, cudaFuncAttributeMaxDynamicSharedMemorySize,
			static_cast<int>(smem)) != cudaSuccess) {
		return 0;
	}
	if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&blocks, SplineFitKernel<NT8>, NTHR, smem)
			!= cudaSuccess) {
		return 0;
	}
	return blocks;
}

} // namespace

bool SplineFastSupported(int n, int npiece) {
	return npiece >= 1 && npiece <= kSplineMaxPieces && n % 4 == 0 && n >= 32
			&& n <= kSplineMaxChannels;
}

size_t SplineSharedBytes(int n, int npiece) {
	return SmemDoubles(npiece) * sizeof(double) + static_cast<size_t>(n) * 9
			+ (static_cast<size_t>(npiece + 1 + 3) / 4 * 4 + NW + 4) * sizeof(int) + 16;
}

int SplineResidentCtas(int n, int npiece, int num_sms) {
	size_t const smem = SplineSharedBytes(n, npiece);
	int per_sm = 0;
	switch (PaddedBases(npiece) / 8) {
	case 1:
		per_sm = ResidentNT<1>(smem);
		break;
	case 2:
		per_sm = ResidentNT<2>(smem);
		break;
	case 3:
		per_sm = ResidentNT<3>(smem);
		break;
	default:
		per_sm = ResidentNT<4>(smem);
		break;
	}
	return (per_sm < 1 ? 1 : per_sm) * num_sms;
}

cudaError_t LaunchSplineFit(SplineFitParams const &p, int grid, cudaStream_t stream) {
	if (!SplineFastSupported(p.n, p.npiece)) {
		return cudaErrorInvalidValue;
	}
	switch (PaddedBases(p.npiece) / 8) {
	case 1:
		return LaunchNT<1>(p, grid, stream);
	case 2:
		return LaunchNT<2>(p, grid, stream);
	case 3:
		return LaunchNT<3>(p, grid, stream);
	case 4:
		return LaunchNT<4>(p, grid, stream);
	default:
		return cudaErrorInvalidValue;
	}
}

} // namespace sakura_b200
