// Batched cubic-spline baseline fit (sakura_LSQFitCubicSplineFloat semantics, baseline.cc:1761-1833
// -> LSQFitCubicSpline :1342 -> DoLSQFit :1115), one 256-thread CTA per spectrum, three CTAs per SM,
// the spectrum, its residual and its working mask (one bit per channel) resident in shared memory for
// the whole fit.
//
// The reference builds a per-spectrum basis table [n][npiece+3] (baseline.cc:911-980: 1, x, x^2,
// x^3 - (x-b_1)^3_+, then (x-b_{j-1})^3_+ - (x-b_j)^3_+), accumulates all (npiece+3)^2 products per
// channel and solves the dense system (cond ~2e11 at 20 pieces) with a full-pivot LU. Those bases
// span the cubic splines with simple knots at the piece boundaries, and the least-squares fit is a
// property of that space, not of the basis. This kernel fits in the B-SPLINE basis of the same
// space (knots b_1..b_{npiece-1}, fourfold end knots at 0 and n): four bases are alive in a piece,
// each a cubic in the piece-local coordinate u = x - x(b_q), B_{q+r}(x) = sum_a M_q[r][a] u^a
// (Cox-de Boor recursion on the coefficients, one thread per piece and row), so with the local
// power sums and data moments of a piece
//   h_q[m] = sum_{valid i in piece q} u_i^m (m = 0..6),   t_q[a] = sum_valid y_i u_i^a (a = 0..3)
// the normal equations are G = sum_q M_q Hank(h_q) M_q^T (4 x 4 blocks on the diagonal band, half
// bandwidth 3) and b = sum_q M_q t_q: 11 moments per channel instead of 552 products, a banded
// LDL^T of ~25 dependent steps instead of a dense 23 x 23 one, and cond(G) ~1e2. A clip pass
// only subtracts the clipped channels from the moments of their piece (downdate, as
// numeric_operation.cc:283-314 does for the matrix). All expansions are about the piece origin
// with non-negative shifts. A row whose pivots are not safely positive is handed to the general
// kernel (full-pivot LU restating Eigen's on the reference's own bases). The model is evaluated in
// the piece-local form; the per-piece global cubic coefficients the API returns
// (baseline.cc:868-893) are the local cubics re-expanded about x = 0.
//
// Boundaries (baseline.cc:808-844), "the last channel of each piece is never examined"
// (:1081-1096), thresholds, early exit, not-enough-data and which outputs a failing row leaves
// untouched follow the general kernel (lsqfit_general.cu), which remains the restatement of the
// reference's arithmetic order.
#include "lsqfit_spline.cuh"
#include "device_common.cuh"

namespace sakura_b200 {

namespace {

constexpr int NTHR = kSplineThreads;
constexpr int NW = NTHR / 32;
constexpr unsigned kFull = 0xffffffffu;
constexpr double kPivotTol = 1e-9; // pivot / diagonal entry below which a row goes to the LU kernel

// exact for 0 <= v < 2^31, without the (slow) conversion pipe
__device__ __forceinline__ double int_to_double(int v) {
	return __hiloint2double(0x43300000, v) - 4503599627370496.0;
}

__device__ __forceinline__ double rcp_pos(double x) {
	double r;
	asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x));
	double e = fma(-x, r, 1.0);
	r = fma(r, e, r);
	e = fma(-x, r, 1.0);
	r = fma(r, e, r);
	return r;
}

struct Smem {
	double *M; // [npiece][4][4] local cubic of the four B-splines alive in every piece
	double *H; // [npiece][8] local power sums (7 used)
	double *T; // [npiece][4] local data moments
	double *E; // [npiece][16] M_q Hank(h_q) M_q^T (lower triangle used); dead after the assembly:
	double *loc; //   [npiece][4] local cubic of the fitted model (aliases E)
	double *F; // [npiece][4] M_q t_q; dead after the assembly:
	double *cvec; //  [32] solution (aliases F; npiece >= 8 or padded)
	double *band; // [32][5] band rows G[k][k-d], d = 0..3, right-hand side; then the factors;
	double *part; //  [3 * NW] block reduction scratch (aliases band)
	float *y; // [n]
	float *r; // [n]
	int *bound; // [npiece + 1]
	int *iscr; // [NW + 4]
	unsigned *bits; // [ceil(n / 32)] working mask, one bit per channel
};

__host__ __device__ inline size_t SmemDoubles(int npiece) {
	int const f = 4 * npiece > 32 ? 4 * npiece : 32;
	return static_cast<size_t>(npiece) * (16 + 8 + 4 + 16) + f + 32 * 5;
}

__host__ __device__ inline size_t SmemInts(int n, int npiece) {
	return static_cast<size_t>(npiece + 1 + 3) / 4 * 4 + NW + 4 + (n + 31) / 32 + 3;
}

__device__ __forceinline__ Smem Carve(unsigned char *base, int n, int npiece) {
	Smem s;
	double *d = reinterpret_cast<double *>(base);
	s.M = d;
	d += npiece * 16;
	s.H = d;
	d += npiece * 8;
	s.T = d;
	d += npiece * 4;
	s.E = d;
	s.loc = d;
	d += npiece * 16;
	s.F = d;
	s.cvec = d;
	d += 4 * npiece > 32 ? 4 * npiece : 32;
	s.band = d;
	s.part = d;
	d += 32 * 5;
	s.y = reinterpret_cast<float *>(d);
	s.r = s.y + n;
	s.bound = reinterpret_cast<int *>(s.r + n);
	s.iscr = s.bound + (npiece + 1 + 3) / 4 * 4;
	s.bits = reinterpret_cast<unsigned *>(s.iscr + NW + 4);
	return s;
}

__device__ __forceinline__ unsigned Nibble(unsigned const *bits, int i) { // i % 4 == 0
	return (bits[i >> 5] >> (i & 31)) & 15u;
}
__device__ __forceinline__ bool Bit(unsigned const *bits, int i) {
	return ((bits[i >> 5] >> (i & 31)) & 1u) != 0u;
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
// the unmasked channels equals ceil(N*idx/npiece). Every thread owns a run of mask words.
__device__ void Boundaries(int n, int npiece, int num_valid, Smem const &s, int tid) {
	int const lane = tid & 31;
	int const warp = tid >> 5;
	int const nwords = (n + 31) >> 5;
	int const chunk = (nwords + NTHR - 1) / NTHR;
	int const lo = min(nwords, tid * chunk);
	int const hi = min(nwords, lo + chunk);
	int cnt = 0;
	for (int w = lo; w < hi; ++w) {
		cnt += __popc(s.bits[w]);
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
	// rank targets once per row (64-bit divisions), then every thread only compares
	int *tgt = reinterpret_cast<int *>(s.E);
	if (tid >= 1 && tid < npiece) {
		long long const N = num_valid;
		tgt[tid] = static_cast<int>((N * tid + npiece - 1) / npiece);
	}
	if (tid == 0) {
		s.bound[0] = 0;
		s.bound[npiece] = n;
	}
	__syncthreads();
	int start = incl - cnt;
	for (int w = 0; w < warp; ++w) {
		start += s.iscr[w];
	}
	for (int idx = 1; idx < npiece; ++idx) {
		int const target = tgt[idx];
		if (target >= start && target < start + cnt) {
			int rem = target - start;
			for (int w = lo; w < hi; ++w) {
				unsigned const v = s.bits[w];
				int const c = __popc(v);
				if (rem < c) {
					// position of the set bit of rank rem
					s.bound[idx] = 32 * w + (__fns(v, 0, rem + 1));
					break;
				}
				rem -= c;
			}
		}
	}
	__syncthreads();
}

// Local cubics of the four B-splines alive in piece q (Cox-de Boor on polynomial coefficients in
// u = x - tau_mu, knots tau = boundary positions / max_x with fourfold end knots): M[r][a],
// basis index q + r.
__device__ void PieceBasis(int q, int npiece, Smem const &s, double inv) {
	auto tau = [&](int j) { // knot j of the extended knot vector, in channels
		return s.bound[min(max(j - 3, 0), npiece)];
	};
	int const mu = q + 3;
	int const t0 = tau(mu);
	double L[4], R[4];
#pragma unroll
	for (int j = 1; j <= 3; ++j) {
		L[j] = int_to_double(t0 - tau(mu + 1 - j)) * inv;
		R[j] = int_to_double(tau(mu + j) - t0) * inv;
	}
	double N[4][4];
#pragma unroll
	for (int r = 0; r < 4; ++r) {
#pragma unroll
		for (int a = 0; a < 4; ++a) {
			N[r][a] = 0.0;
		}
	}
	N[0][0] = 1.0;
#pragma unroll
	for (int j = 1; j <= 3; ++j) {
		double saved[4] = { 0.0, 0.0, 0.0, 0.0 };
#pragma unroll
		for (int r = 0; r < j; ++r) {
			double const rd = 1.0 / (R[r + 1] + L[j - r]);
			double temp[4];
#pragma unroll
			for (int a = 0; a < 4; ++a) {
				temp[a] = N[r][a] * rd;
			}
			// N[r] = saved + (R[r+1] - u) temp;  saved = (u + L[j-r]) temp
#pragma unroll
			for (int a = 3; a >= 0; --a) {
				double const up = (a > 0) ? temp[a - 1] : 0.0;
				N[r][a] = fma(R[r + 1], temp[a], saved[a]) - up;
				saved[a] = fma(L[j - r], temp[a], up);
			}
		}
#pragma unroll
		for (int a = 0; a < 4; ++a) {
			N[j][a] = saved[a];
		}
	}
	double *m = s.M + q * 16;
#pragma unroll
	for (int r = 0; r < 4; ++r) {
#pragma unroll
		for (int a = 0; a < 4; ++a) {
			m[4 * r + a] = N[r][a];
		}
	}
}

// Warp 0: band rows from the piece blocks, banded LDL^T bordered by the right-hand side, back
// substitution, local cubic of the model per piece. Returns false (warp-uniform) when a pivot is not
// safely positive. The serial loops run redundantly in all lanes on broadcast shared-memory reads;
// lane k keeps what belongs to basis k.
__device__ bool BandSolve(int npiece, int K, Smem const &s, int lane) {
	double *band = s.band;
	if (lane < K) {
		int const k = lane;
		int const q0 = max(0, k - 3);
#pragma unroll
		for (int d = 0; d < 4; ++d) {
			double v = 0.0;
			if (k - d >= 0) {
				int const q1 = min(npiece - 1, k - d);
				for (int q = q0; q <= q1; ++q) {
					v += s.E[q * 16 + 4 * (k - q) + (k - d - q)];
				}
			}
			band[5 * k + d] = v;
		}
		double v = 0.0;
		int const q1 = min(npiece - 1, k);
		for (int q = q0; q <= q1; ++q) {
			v += s.F[4 * q + (k - q)];
		}
		band[5 * k + 4] = v;
	}
	__syncwarp();
	// row i: w_d = G[i][i-d] - (earlier columns), l_d = w_d / D[i-d], D[i] = G[i][i] - sum w_d l_d,
	// z[i] = b[i] - sum l_d z[i-d]
	double p12 = 0.0, p13 = 0.0, p23 = 0.0; // L[i-1][i-2], L[i-1][i-3], L[i-2][i-3]
	double r1 = 0.0, r2 = 0.0, r3 = 0.0; // 1 / D[i-1], 1 / D[i-2], 1 / D[i-3]
	double z1 = 0.0, z2 = 0.0, z3 = 0.0;
	double ml1 = 0.0, ml2 = 0.0, ml3 = 0.0, mzr = 0.0; // lane i: L[i][i-1..i-3], z[i] / D[i]
	bool ok = true;
#pragma unroll 4
	for (int i = 0; i < K; ++i) {
		double const g0 = band[5 * i], g1 = band[5 * i + 1], g2 = band[5 * i + 2], g3 = band[5 * i + 3];
		double const bi = band[5 * i + 4];
		double const w3 = g3;
		double const w2 = fma(-w3, p23, g2);
		double const w1 = fma(-w2, p12, fma(-w3, p13, g1));
		double const l3 = w3 * r3;
		double const l2 = w2 * r2;
		double const l1 = w1 * r1;
		double const di = fma(-w1, l1, fma(-w2, l2, fma(-w3, l3, g0)));
		double const zi = fma(-l1, z1, fma(-l2, z2, fma(-l3, z3, bi)));
		ok = ok && (di > kPivotTol * g0);
		double const ri = rcp_pos(di);
		if (lane == i) {
			ml1 = l1;
			ml2 = l2;
			ml3 = l3;
			mzr = zi * ri;
		}
		p23 = p12;
		p13 = l2;
		p12 = l1;
		r3 = r2;
		r2 = r1;
		r1 = ri;
		z3 = z2;
		z2 = z1;
		z1 = zi;
	}
	if (!ok) {
		return false;
	}
	// L^T x = D^-1 z: x[i] = z[i]/D[i] - L[i+1][i] x[i+1] - L[i+2][i] x[i+2] - L[i+3][i] x[i+3]
	double const c1 = __shfl_down_sync(kFull, ml1, 1);
	double const c2 = __shfl_down_sync(kFull, ml2, 2);
	double const c3 = __shfl_down_sync(kFull, ml3, 3);
	__syncwarp();
	if (lane < K) {
		band[5 * lane] = mzr;
		band[5 * lane + 1] = (lane + 1 < K) ? c1 : 0.0;
		band[5 * lane + 2] = (lane + 2 < K) ? c2 : 0.0;
		band[5 * lane + 3] = (lane + 3 < K) ? c3 : 0.0;
	}
	__syncwarp();
	double x1 = 0.0, x2 = 0.0, x3 = 0.0, mine = 0.0;
#pragma unroll 4
	for (int i = K - 1; i >= 0; --i) {
		double const xi = fma(-band[5 * i + 1], x1, fma(-band[5 * i + 2], x2, fma(-band[5 * i + 3], x3, band[5 * i])));
		if (lane == i) {
			mine = xi;
		}
		x3 = x2;
		x2 = x1;
		x1 = xi;
	}
	s.cvec[lane] = mine;
	__syncwarp();
	for (int e = lane; e < 4 * npiece; e += 32) { // E is dead: loc may be written
		int const q = e >> 2;
		int const a = e & 3;
		double const *m = s.M + q * 16 + a;
		s.loc[e] = fma(s.cvec[q + 3], m[12], fma(s.cvec[q + 2], m[8], fma(s.cvec[q + 1], m[4], s.cvec[q] * m[0])));
	}
	return true;
}

__global__ void __launch_bounds__(NTHR, 3)
SplineFitKernel(SplineFitParams const p) {
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
		for (int i0 = warp * 128; i0 < n; i0 += NTHR * 4) { // warp-uniform bounds: shuffles inside
			int const i = i0 + 4 * lane;
			unsigned v = 0u;
			if (i < n) {
				v = *reinterpret_cast<unsigned const *>(mask + i);
				*reinterpret_cast<float4 *>(s.y + i) = *reinterpret_cast<float4 const *>(data + i);
			}
			v = __vcmpne4(v, 0u) & 0x01010101u;
			v = ((v * 0x10204080u) >> 28) << (4 * (lane & 7)); // the four flags as a nibble
			v |= __shfl_xor_sync(kFull, v, 1);
			v |= __shfl_xor_sync(kFull, v, 2);
			v |= __shfl_xor_sync(kFull, v, 4);
			if ((lane & 7) == 0 && i < n) {
				s.bits[i >> 5] = v;
				cnt += __popc(v);
			}
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
		// ---- B-spline pieces (the last warp has the fewest pieces to sum below) -------------------------
		if (warp == NW - 1 && lane < npiece) {
			PieceBasis(lane, npiece, s, inv);
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
				unsigned const m = Nibble(s.bits, i);
				one(i, y.x, (m & 1u) != 0u);
				one(i + 1, y.y, (m & 2u) != 0u);
				one(i + 2, y.z, (m & 4u) != 0u);
				one(i + 3, y.w, (m & 8u) != 0u);
			}, [&](int i) {
				one(i, s.y[i], Bit(s.bits, i));
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
			// ---- piece blocks of the normal equations from the moments ------------------------------
			if (tid < 4 * npiece) {
				int const q = tid >> 2;
				int const c = tid & 3;
				double const *m = s.M + q * 16;
				double const *h = s.H + q * 8;
				double const *t = s.T + q * 4;
				double const m0 = m[4 * c], m1 = m[4 * c + 1], m2 = m[4 * c + 2], m3 = m[4 * c + 3];
				double w[4];
#pragma unroll
				for (int a = 0; a < 4; ++a) {
					w[a] = fma(m3, h[a + 3], fma(m2, h[a + 2], fma(m1, h[a + 1], m0 * h[a])));
				}
#pragma unroll
				for (int r = 0; r < 4; ++r) {
					s.E[q * 16 + 4 * r + c] = fma(m[4 * r + 3], w[3], fma(m[4 * r + 2], w[2], fma(m[4 * r + 1], w[1], m[4 * r] * w[0])));
				}
				s.F[tid] = fma(m3, t[3], fma(m2, t[2], fma(m1, t[1], m0 * t[0])));
			}
			__syncthreads();
			if (warp == 0) {
				bool const ok = BandSolve(npiece, K, s, lane);
				if (lane == 0) {
					s.iscr[NW] = ok ? 1 : 0;
				}
			}
			__syncthreads();
			if (s.iscr[NW] == 0) {
				fallback = true;
				break;
			}
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
					unsigned const m = Nibble(s.bits, i);
					float4 rr;
					rr.x = one(i, y.x, (m & 1u) != 0u);
					rr.y = one(i + 1, y.y, (m & 2u) != 0u);
					rr.z = one(i + 2, y.z, (m & 4u) != 0u);
					rr.w = one(i + 3, y.w, (m & 8u) != 0u);
					*reinterpret_cast<float4 *>(s.r + i) = rr;
				}, [&](int i) {
					s.r[i] = one(i, s.y[i], Bit(s.bits, i));
				});
			}
			block_sum3(c, sum, sq, s.part); // band is dead: part may be written
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
					atomicAnd(s.bits + (i >> 5), ~(1u << (i & 31)));
					++nclip;
					mine = true;
					AddMoments(h, t, int_to_double(i - b0) * inv, static_cast<double>(s.y[i]));
				};
				ForPiece(b0, last, lane, [&](int i) {
					float4 const r = *reinterpret_cast<float4 const *>(s.r + i);
					unsigned const m = Nibble(s.bits, i);
					unsigned bits = 0u;
					bits |= ((m & 1u) != 0u && is_clipped(r.x, th.lower, th.upper)) ? 1u : 0u;
					bits |= ((m & 2u) != 0u && is_clipped(r.y, th.lower, th.upper)) ? 2u : 0u;
					bits |= ((m & 4u) != 0u && is_clipped(r.z, th.lower, th.upper)) ? 4u : 0u;
					bits |= ((m & 8u) != 0u && is_clipped(r.w, th.lower, th.upper)) ? 8u : 0u;
					while (bits != 0u) { // rare
						hit(i + __ffs(bits) - 1);
						bits &= bits - 1u;
					}
				}, [&](int i) {
					if (Bit(s.bits, i) && is_clipped(s.r[i], th.lower, th.upper)) {
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
			*reinterpret_cast<unsigned *>(fmask + i) = (Nibble(s.bits, i) * 0x00204081u) & 0x01010101u;
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
		if (p.coeff != nullptr && tid < npiece) {
			// the piece's cubic about x = 0 (what baseline.cc:868-893 derives from its own bases),
			// then per power of max_x as the reference's epilogue
			int const q = tid;
			double const xq = int_to_double(s.bound[q]) * inv;
			double const l0 = s.loc[4 * q], l1 = s.loc[4 * q + 1], l2 = s.loc[4 * q + 2], l3 = s.loc[4 * q + 3];
			double const c3 = l3;
			double const c2 = fma(-3.0 * l3, xq, l2);
			double const c1 = fma(fma(3.0 * l3, xq, -2.0 * l2), xq, l1);
			double const c0 = fma(fma(fma(-l3, xq, l2), xq, -l1), xq, l0);
			double *co = p.coeff + row * p.coeff_stride + 4 * q;
			co[0] = c0;
			co[1] = c1 / max_x;
			co[2] = c2 / (max_x * max_x);
			co[3] = c3 / (max_x * max_x * max_x);
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

cudaError_t LaunchImpl(SplineFitParams const &p, int grid, cudaStream_t stream) {
	size_t const smem = SplineSharedBytes(p.n, p.npiece);
	cudaError_t err = cudaFuncSetAttribute(SplineFitKernel,
			cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem));
	if (err != cudaSuccess) {
		return err;
	}
	SplineFitKernel<<<grid, NTHR, smem, stream>>>(p);
	return cudaGetLastError();
}

} // namespace

bool SplineFastSupported(int n, int npiece) {
	return npiece >= 1 && npiece <= kSplineMaxPieces && n % 4 == 0 && n >= 32
			&& n <= kSplineMaxChannels;
}

size_t SplineSharedBytes(int n, int npiece) {
	return SmemDoubles(npiece) * sizeof(double) + static_cast<size_t>(n) * 8
			+ SmemInts(n, npiece) * sizeof(int);
}

int SplineResidentCtas(int n, int npiece, int num_sms) {
	size_t const smem = SplineSharedBytes(n, npiece);
	int per_sm = 0;
	if (cudaFuncSetAttribute(SplineFitKernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
			static_cast<int>(smem)) != cudaSuccess
			|| cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, SplineFitKernel, NTHR, smem)
					!= cudaSuccess) {
		per_sm = 0;
	}
	return (per_sm < 1 ? 1 : per_sm) * num_sms;
}

cudaError_t LaunchSplineFit(SplineFitParams const &p, int grid, cudaStream_t stream) {
	if (!SplineFastSupported(p.n, p.npiece)) {
		return cudaErrorInvalidValue;
	}
	return LaunchImpl(p, grid, stream);
}

} // namespace sakura_b200
