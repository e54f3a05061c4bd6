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
#include <stdio.h>
#include "lsqfit_spline.cuh"
#include "device_common.cuh"

namespace sakura_b200 {

namespace {

constexpr int NTHR = kSplineThreads;
constexpr int NW = NTHR / 32;
constexpr unsigned kFull = 0xffffffffu;
constexpr double kPivotTol = 1e-9; // pivot / diagonal entry below which a row goes to the LU kernel

// -DSPLINE_CLOCKS: CTA 0 prints the cycles its warp 0 spent per phase (development only)
#ifdef SPLINE_CLOCKS
#define CLK(k) do { long long const t_ = clock64(); clk[k] += t_ - tprev; tprev = t_; } while (0)
#define CLK_ARGS , long long (&clk)[16], long long &tprev
#define CLK_PASS , clk, tprev
#else
#define CLK(k) do { } while (0)
#define CLK_ARGS
#define CLK_PASS
#endif

// exact for 0 <= v < 2^31, without the (slow) conversion pipe
__device__ __forceinline__ double int_to_double(int v) {
	return __hiloint2double(0x43300000, v) - 4503599627370496.0;
}

// reciprocal of a positive, normal double: hardware estimate (e ~ 2^-23) + one third-order step,
// r (1 + e + e^2) = (1 - e^3) / x: ~1 ulp with three dependent operations after the estimate
__device__ __forceinline__ double rcp_pos(double x) {
	double r;
	asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x));
	double const e = fma(-x, r, 1.0);
	return fma(r, fma(e, e, e), r);
}

constexpr int kGpStride = 11; // 10 band contributions per basis row (odd: conflict-free lane reads)
constexpr int kFpStride = 5;
constexpr int kBandStride = 6;

struct Smem {
	double *M; // [npiece][4][4] local cubic of the four B-splines alive in every piece
	double *H; // [npiece][8] local power sums (7 used)
	double *T; // [npiece][4] local data moments
	double *Gp; // [K][11] piece contributions to band row k: slot kGpBase[d] + c is E_q[c + d][c], q = k - d - c;
	double *loc; //   dead after the assembly: [npiece][4] local cubic of the fitted model (aliases Gp)
	double *Fp; // [K][5] piece contributions to the right-hand side: slot r is F_q[r], q = k - r;
	double *cvec; //  dead after the assembly: [32] solution (aliases Fp, K >= 7 or padded)
	double *band; // [K][6] band rows G[k][k-d], d = 0..3, right-hand side; then the factors;
	double *part; //  [3 * NW] block reduction scratch (aliases band, K >= 4)
	float *y; // [n]
	float *r; // [n]
	int *bound; // [npiece + 1]
	int *iscr; // [NW + 4]
	unsigned *bits; // [ceil(n / 32)] working mask, one bit per channel
};

// (all sizes even: the arrays behind them stay 16-byte aligned)
__host__ __device__ inline int GpDoubles(int npiece) {
	return ((npiece + 3) * kGpStride + 1) & ~1;
}
__host__ __device__ inline int FpDoubles(int npiece) {
	int const f = ((npiece + 3) * kFpStride + 1) & ~1;
	return f > 32 ? f : 32;
}
__host__ __device__ inline int BandDoubles(int npiece) {
	return (npiece + 3) * kBandStride > 3 * NW ? (npiece + 3) * kBandStride : 3 * NW;
}

__host__ __device__ inline size_t SmemDoubles(int npiece) {
	return static_cast<size_t>(npiece) * (16 + 8 + 4) + GpDoubles(npiece) + FpDoubles(npiece)
			+ BandDoubles(npiece);
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
	s.Gp = d;
	s.loc = d;
	d += GpDoubles(npiece);
	s.Fp = d;
	s.cvec = d;
	d += FpDoubles(npiece);
	s.band = d;
	s.part = d;
	d += BandDoubles(npiece);
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
template<int UNROLL = 1, class Vec, class One>
__device__ __forceinline__ void ForPiece(int b0, int b1, int lane, Vec vec, One one) {
	int const a0 = min((b0 + 3) & ~3, b1);
	int const a1 = max(b1 & ~3, a0);
	if (lane < a0 - b0) {
		one(b0 + lane);
	} else if (lane >= 16 && lane - 16 < b1 - a1) {
		one(a1 + lane - 16);
	}
#pragma unroll UNROLL
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

// The same sums for four channels at once, flagged channels silenced by u = 0, y = 0 (they then add
// nothing to any sum but h[0], which is counted from the mask bits instead): every accumulator takes
// ONE dependent addition per call, the four channels' terms are combined in a tree first.
__device__ __forceinline__ void AddMoments4(double (&h)[7], double (&t)[4], double d0, double inv,
		float4 yf, unsigned m) {
	double u[4], y[4], sq[4], cu[4];
	float const ys[4] = { yf.x, yf.y, yf.z, yf.w };
#pragma unroll
	for (int j = 0; j < 4; ++j) {
		bool const valid = (m >> j) & 1u;
		double const uj = (d0 + static_cast<double>(j)) * inv;
		u[j] = valid ? uj : 0.0;
		y[j] = static_cast<double>(valid ? ys[j] : 0.0f);
		sq[j] = u[j] * u[j];
		cu[j] = sq[j] * u[j];
	}
	h[1] += (u[0] + u[1]) + (u[2] + u[3]);
	h[2] += (sq[0] + sq[1]) + (sq[2] + sq[3]);
	h[3] += (cu[0] + cu[1]) + (cu[2] + cu[3]);
	h[4] += fma(sq[0], sq[0], sq[1] * sq[1]) + fma(sq[2], sq[2], sq[3] * sq[3]);
	h[5] += fma(sq[0], cu[0], sq[1] * cu[1]) + fma(sq[2], cu[2], sq[3] * cu[3]);
	h[6] += fma(cu[0], cu[0], cu[1] * cu[1]) + fma(cu[2], cu[2], cu[3] * cu[3]);
	t[0] += (y[0] + y[1]) + (y[2] + y[3]);
	t[1] += fma(y[0], u[0], y[1] * u[1]) + fma(y[2], u[2], y[3] * u[3]);
	t[2] += fma(y[0], sq[0], y[1] * sq[1]) + fma(y[2], sq[2], y[3] * sq[3]);
	t[3] += fma(y[0], cu[0], y[1] * cu[1]) + fma(y[2], cu[2], y[3] * cu[3]);
}

// Block-wide sums without a leading barrier (callers guarantee that the scratch is free): warp totals
// on the tensor cores, the eight warp results added by one more warp total in every warp.
__device__ __forceinline__ int BlockSumInt(int v, int *scratch, int lane, int warp) {
	v = __reduce_add_sync(kFull, v);
	if (lane == 0) {
		scratch[warp] = v;
	}
	__syncthreads();
	return __reduce_add_sync(kFull, lane < NW ? scratch[lane] : 0);
}

__device__ __forceinline__ void BlockSum3(int &cnt, double &s1, double &s2, double *scratch, int lane, int warp) {
	cnt = __reduce_add_sync(kFull, cnt);
	s1 = warp_total(s1);
	s2 = warp_total(s2);
	if (lane == 0) {
		scratch[warp] = s1;
		scratch[NW + warp] = s2;
		reinterpret_cast<int *>(scratch + 2 * NW)[warp] = cnt;
	}
	__syncthreads();
	cnt = __reduce_add_sync(kFull, lane < NW ? reinterpret_cast<int const *>(scratch + 2 * NW)[lane] : 0);
	s1 = warp_total(lane < NW ? scratch[lane] : 0.0);
	s2 = warp_total(lane < NW ? scratch[NW + lane] : 0.0);
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
	int *tgt = reinterpret_cast<int *>(s.Gp);
	unsigned const N = static_cast<unsigned>(num_valid); // N * npiece < 2^31: 32-bit divisions
	if (tid >= 1 && tid < npiece) {
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
	if (cnt > 0) {
		// the targets ceil(N idx / npiece) ascend: the first one >= start has idx > (start - 1) npiece / N
		int idx = (start == 0) ? 1 : static_cast<int>((static_cast<unsigned>(start - 1) * npiece) / N) + 1;
		for (; idx < npiece; ++idx) {
			int const target = tgt[idx];
			if (target >= start + cnt) {
				break;
			}
			if (target >= start) {
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

// Warp 0: band rows from the piece contributions, banded LDL^T bordered by the right-hand side, back
// substitution. Returns false (warp-uniform) when a pivot is not safely positive. The serial loops run
// redundantly in all lanes on broadcast shared-memory reads; lane k keeps what belongs to basis k.
__device__ bool BandSolve(int npiece, int K, Smem const &s, int lane CLK_ARGS) {
	double *band = s.band;
	if (lane < K) {
		int const k = lane;
		double const *gp = s.Gp + k * kGpStride;
		double const *fp = s.Fp + k * kFpStride;
		// contribution (d, c) comes from piece q = k - d - c, slot r of the right-hand side from piece k - r
		auto alive = [&](int r) {
			return r <= k && k - r < npiece;
		};
		double const g0 = ((alive(0) ? gp[0] : 0.0) + (alive(1) ? gp[1] : 0.0))
				+ ((alive(2) ? gp[2] : 0.0) + (alive(3) ? gp[3] : 0.0));
		double const g1 = ((alive(1) ? gp[4] : 0.0) + (alive(2) ? gp[5] : 0.0)) + (alive(3) ? gp[6] : 0.0);
		double const g2 = (alive(2) ? gp[7] : 0.0) + (alive(3) ? gp[8] : 0.0);
		double const g3 = alive(3) ? gp[9] : 0.0;
		double const bk = ((alive(0) ? fp[0] : 0.0) + (alive(1) ? fp[1] : 0.0))
				+ ((alive(2) ? fp[2] : 0.0) + (alive(3) ? fp[3] : 0.0));
		double *row = band + kBandStride * k;
		row[0] = g0;
		row[1] = g1;
		row[2] = g2;
		row[3] = g3;
		row[4] = bk;
	}
	__syncwarp();
	CLK(2);
	// row i: w_d = G[i][i-d] - (earlier columns), l_d = w_d / D[i-d], D[i] = G[i][i] - sum w_d l_d,
	// z[i] = b[i] - sum l_d z[i-d]
	double p12 = 0.0, p13 = 0.0, p23 = 0.0; // L[i-1][i-2], L[i-1][i-3], L[i-2][i-3]
	double r1 = 0.0, r2 = 0.0, r3 = 0.0; // 1 / D[i-1], 1 / D[i-2], 1 / D[i-3]
	double z1 = 0.0, z2 = 0.0, z3 = 0.0;
	double ml1 = 0.0, ml2 = 0.0, ml3 = 0.0, mzr = 0.0; // lane i: L[i][i-1..i-3], z[i] / D[i]
	bool ok = true;
#pragma unroll 4
	for (int i = 0; i < K; ++i) {
		double const *row = band + kBandStride * i;
		double const g0 = row[0], g1 = row[1], g2 = row[2], g3 = row[3];
		double const bi = row[4];
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
	CLK(3);
	if (!ok) {
		return false;
	}
	// L^T x = D^-1 z: x[i] = z[i]/D[i] - L[i+1][i] x[i+1] - L[i+2][i] x[i+2] - L[i+3][i] x[i+3]
	double const c1 = __shfl_down_sync(kFull, ml1, 1);
	double const c2 = __shfl_down_sync(kFull, ml2, 2);
	double const c3 = __shfl_down_sync(kFull, ml3, 3);
	__syncwarp();
	if (lane < K) {
		double *row = band + kBandStride * lane;
		row[0] = mzr;
		row[1] = (lane + 1 < K) ? c1 : 0.0;
		row[2] = (lane + 2 < K) ? c2 : 0.0;
		row[3] = (lane + 3 < K) ? c3 : 0.0;
	}
	__syncwarp();
	double x1 = 0.0, x2 = 0.0, x3 = 0.0, mine = 0.0;
#pragma unroll 4
	for (int i = K - 1; i >= 0; --i) {
		double const *row = band + kBandStride * i;
		double const xi = fma(-row[1], x1, fma(-row[2], x2, fma(-row[3], x3, row[0])));
		if (lane == i) {
			mine = xi;
		}
		x3 = x2;
		x2 = x1;
		x1 = xi;
	}
	s.cvec[lane] = mine;
	CLK(4);
	return true;
}

__device__ __forceinline__ void LoadRowAsync(float *sy, float const *data, int n, int tid) {
	for (int i = tid * 4; i < n; i += NTHR * 4) {
		asm volatile("cp.async.ca.shared.global [%0], [%1], 16;"
				:: "r"(static_cast<unsigned>(__cvta_generic_to_shared(sy + i))), "l"(data + i) : "memory");
	}
	asm volatile("cp.async.commit_group;" ::: "memory");
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
#ifdef SPLINE_CLOCKS
	long long clk[16] = { 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0 };
	long long tprev = clock64();
#endif

	bool preloaded = false;
	for (size_t row = blockIdx.x; row < p.num_spectra; row += gridDim.x) {
		float const *data = p.data + row * p.data_stride;
		uint8_t const *mask = p.mask + row * p.mask_stride;
		__syncthreads(); // previous row's shared state fully consumed
		CLK(11);
		// the spectrum: asynchronous 16-byte copies straight into shared memory, all in flight at once
		// (already under way when the previous row of this CTA started them behind its last pass)
		if (!preloaded) {
			LoadRowAsync(s.y, data, n, tid);
		}
		preloaded = false;
		// the mask: bytes -> one bit per channel, four 128-channel groups per warp and step in flight
		int cnt = 0;
		constexpr int kGroups = 4;
		for (int i0 = warp * 128; i0 < n; i0 += NTHR * 4 * kGroups) { // warp-uniform bounds: shuffles inside
			unsigned v[kGroups];
#pragma unroll
			for (int g = 0; g < kGroups; ++g) {
				int const i = i0 + g * NTHR * 4 + 4 * lane;
				v[g] = (i < n) ? *reinterpret_cast<unsigned const *>(mask + i) : 0u;
			}
#pragma unroll
			for (int g = 0; g < kGroups; ++g) {
				int const i = i0 + g * NTHR * 4 + 4 * lane;
				if (i0 + g * NTHR * 4 >= n) {
					break;
				}
				unsigned w = __vcmpne4(v[g], 0u) & 0x01010101u;
				w = ((w * 0x10204080u) >> 28) << (4 * (lane & 7)); // the four flags as a nibble
				w |= __shfl_xor_sync(kFull, w, 1);
				w |= __shfl_xor_sync(kFull, w, 2);
				w |= __shfl_xor_sync(kFull, w, 4);
				if ((lane & 7) == 0 && i < n) {
					s.bits[i >> 5] = w;
					cnt += __popc(w);
				}
			}
		}
		asm volatile("cp.async.wait_group 0;" ::: "memory");
		if (row + gridDim.x < p.num_spectra) { // the next spectrum of this CTA: into L2 while this one is fitted
			char const *nd = reinterpret_cast<char const *>(data + gridDim.x * p.data_stride);
			char const *nm = reinterpret_cast<char const *>(mask + gridDim.x * p.mask_stride);
			for (int b = tid * 128; b < n * 4; b += NTHR * 128) {
				asm volatile("prefetch.global.L2 [%0];" :: "l"(nd + b));
			}
			for (int b = tid * 128; b < n; b += NTHR * 128) {
				asm volatile("prefetch.global.L2 [%0];" :: "l"(nm + b));
			}
		}
		int const num_valid = BlockSumInt(cnt, s.iscr, lane, warp);
		CLK(12);
		if (num_valid < npiece) { // baseline.cc:822-824 -> Status_kNG, lsqfit_status stays kNG
			if (tid == 0) {
				p.status[row] = kStatusNG;
				p.lsqfit_status[row] = kLsqNG;
			}
			continue;
		}
		Boundaries(n, npiece, num_valid, s, tid);
		CLK(13);
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
		CLK(14);
		// ---- local moments, one warp per piece ---------------------------------------------------------
		for (int q = warp; q < npiece; q += NW) {
			int const b0 = s.bound[q];
			int const b1 = s.bound[q + 1];
			double h[7] = { 0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0 };
			double t[4] = { 0.0, 0.0, 0.0, 0.0 };
			int nv = 0;
			ForPiece(b0, b1, lane, [&](int i) {
				unsigned const m = Nibble(s.bits, i);
				nv += __popc(m);
				AddMoments4(h, t, int_to_double(i - b0), inv, *reinterpret_cast<float4 const *>(s.y + i), m);
			}, [&](int i) {
				if (Bit(s.bits, i)) {
					AddMoments(h, t, int_to_double(i - b0) * inv, static_cast<double>(s.y[i]));
				}
			});
			h[0] += static_cast<double>(nv);
			// warp totals: every lane gets them, lane m < 7 keeps h[m], lane 8 + a keeps t[a]
			double mine = 0.0;
#pragma unroll
			for (int m = 0; m < 7; ++m) {
				double const tot = warp_total(h[m]);
				mine = (lane == m) ? tot : mine;
			}
#pragma unroll
			for (int a = 0; a < 4; ++a) {
				double const tot = warp_total(t[a]);
				mine = (lane == 8 + a) ? tot : mine;
			}
			if (lane < 12) { // H[q][0..7] and T[q][0..3] are contiguous per kind
				(lane < 8 ? s.H + q * 8 + lane : s.T + q * 4 + (lane - 8))[0] = mine;
			}
		}
		__syncthreads();
		CLK(0); // load, boundaries, bases, moments

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
				// E_q[r][c] (r >= c) belongs to band row q + r, offset d = r - c: slot kGpBase[d] + c
#pragma unroll
				for (int d = 0; d < 4; ++d) {
					int const r = c + d;
					if (r < 4) {
						int const base = d == 0 ? 0 : (d == 1 ? 4 : (d == 2 ? 7 : 9));
						s.Gp[(q + r) * kGpStride + base + c] = fma(m[4 * r + 3], w[3],
								fma(m[4 * r + 2], w[2], fma(m[4 * r + 1], w[1], m[4 * r] * w[0])));
					}
				}
				s.Fp[(q + c) * kFpStride + c] = fma(m3, t[3], fma(m2, t[2], fma(m1, t[1], m0 * t[0])));
			}
			__syncthreads();
			CLK(1);
			if (warp == 0) {
				bool const ok = BandSolve(npiece, K, s, lane CLK_PASS);
				if (lane == 0) {
					s.iscr[NW] = ok ? 1 : 0;
				}
			}
			__syncthreads();
			CLK(6);
			if (s.iscr[NW] == 0) {
				fallback = true;
				break;
			}
			// ---- local cubic of the model in the pieces of this warp ------------------------------------
			for (int e = lane; e < 4 * ((npiece - warp + NW - 1) / NW); e += 32) { // Gp is dead: loc may be written
				int const q = warp + NW * (e >> 2);
				int const a = e & 3;
				double const *m = s.M + q * 16 + a;
				s.loc[4 * q + a] = fma(s.cvec[q + 3], m[12], fma(s.cvec[q + 2], m[8], fma(s.cvec[q + 1], m[4], s.cvec[q] * m[0])));
			}
			__syncwarp();
			// ---- model, residual, statistics: one warp per piece -------------------------------------
			int c = 0;
			double sum = 0.0, sq = 0.0;
			for (int q = warp; q < npiece; q += NW) {
				int const b0 = s.bound[q];
				int const b1 = s.bound[q + 1];
				double const l0 = s.loc[4 * q], l1 = s.loc[4 * q + 1], l2 = s.loc[4 * q + 2], l3 = s.loc[4 * q + 3];
				auto resid = [&](double d, float yf) {
					double const u = d * inv;
					return __fsub_rn(yf, static_cast<float>(fma(fma(fma(l3, u, l2), u, l1), u, l0)));
				};
				ForPiece(b0, b1, lane, [&](int i) {
					float4 const y = *reinterpret_cast<float4 const *>(s.y + i);
					unsigned const m = Nibble(s.bits, i);
					double const d0 = int_to_double(i - b0);
					float4 rr;
					rr.x = resid(d0, y.x);
					rr.y = resid(d0 + 1.0, y.y);
					rr.z = resid(d0 + 2.0, y.z);
					rr.w = resid(d0 + 3.0, y.w);
					*reinterpret_cast<float4 *>(s.r + i) = rr;
					// flagged channels contribute +0.0 (a select: NaN/Inf under the mask must not leak)
					double const r0 = static_cast<double>((m & 1u) ? rr.x : 0.0f);
					double const r1 = static_cast<double>((m & 2u) ? rr.y : 0.0f);
					double const r2 = static_cast<double>((m & 4u) ? rr.z : 0.0f);
					double const r3 = static_cast<double>((m & 8u) ? rr.w : 0.0f);
					c += __popc(m);
					sum += (r0 + r1) + (r2 + r3);
					sq += fma(r0, r0, r1 * r1) + fma(r2, r2, r3 * r3);
				}, [&](int i) {
					float const rr = resid(int_to_double(i - b0), s.y[i]);
					s.r[i] = rr;
					bool const valid = Bit(s.bits, i);
					double const rd = static_cast<double>(valid ? rr : 0.0f);
					c += valid ? 1 : 0;
					sum += rd;
					sq = fma(rd, rd, sq);
				});
			}
			CLK(7);
			BlockSum3(c, sum, sq, s.part, lane, warp); // band is dead: part may be written
			CLK(8);
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
				int pend = -1; // a clipped channel whose moments are not yet taken (most lanes meet at most one
				// per piece: all of them are then processed together after the scan, not one by one)
				auto hit = [&](int i) { // channel i is clipped: unflag it, remember its moments
					atomicAnd(s.bits + (i >> 5), ~(1u << (i & 31)));
					++nclip;
					mine = true;
					if (pend >= 0) {
						AddMoments(h, t, int_to_double(pend - b0) * inv, static_cast<double>(s.y[pend]));
					}
					pend = i;
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
				if (pend >= 0) {
					AddMoments(h, t, int_to_double(pend - b0) * inv, static_cast<double>(s.y[pend]));
				}
				if (__any_sync(kFull, mine)) {
					double keep = 0.0;
#pragma unroll
					for (int m = 0; m < 7; ++m) {
						double const tot = warp_total(h[m]);
						keep = (lane == m) ? tot : keep;
					}
#pragma unroll
					for (int a = 0; a < 4; ++a) {
						double const tot = warp_total(t[a]);
						keep = (lane == 8 + a) ? tot : keep;
					}
					if (lane < 12) {
						(lane < 8 ? s.H + q * 8 + lane : s.T + q * 4 + (lane - 8))[0] -= keep;
					}
				}
			}
			CLK(9);
			int const num_clipped = BlockSumInt(nclip, s.iscr, lane, warp);
			CLK(10);
			if (num_clipped == 0) { // baseline.cc:1211-1213
				break;
			}
			num_unmasked = c - num_clipped;
			__syncthreads();
		}
		// every warp is past its last pass over the spectrum (a block barrier lies behind each way out of the
		// loop): the next row's copy can run while this row's results are written
		if (row + gridDim.x < p.num_spectra) {
			LoadRowAsync(s.y, data + gridDim.x * p.data_stride, n, tid);
			preloaded = true;
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
#ifdef SPLINE_CLOCKS
	CLK(11);
	if (blockIdx.x == 0 && tid == 0) {
		printf("spline clocks (CTA 0, warp 0): load %lld bounds %lld basis %lld moments %lld | blocks %lld assemble %lld factor %lld backsub %lld - %lld "
				"(unused) barrier %lld | model %lld sum3 %lld | clip %lld sumint %lld | outputs+load %lld\n", clk[12], clk[13], clk[14], clk[0], clk[1], clk[2],
				clk[3], clk[4], clk[5], clk[6], clk[7], clk[8], clk[9], clk[10], clk[11]);
	}
#endif
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
