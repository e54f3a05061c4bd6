// Polynomial baseline fit with *screened* clip iterations (sakura_LSQFitPolynomialFloat semantics,
// libsakura/src/baseline.cc:1115-1224, 1268-1329): the second generation of the north-star kernel.
//
// PolyFitKernel (lsqfit_poly.cu) evaluates the FP64 model, the float residual and the FP64 residual
// statistics at every channel in every clip iteration: ~16 instructions per channel and iteration,
// 7 of them on the FP64 pipe. That work, not HBM, bounds it. This kernel produces the same clip
// decisions with a fraction of the work:
//
//   * statistics of an intermediate iteration come from the moments the fit already maintains
//     (sum r = T_0 - c.S, sum r^2 = Q - 2 c.T + c^T G c, Q = sum y^2) together with a rigorous
//     bound on what the reference's float residuals would add to them. That gives an INTERVAL for
//     each of the two float clip thresholds of baseline.cc:1200-1207 (usually 0-2 ulp wide);
//   * every channel keeps an approximate float residual r~ in shared memory with a known error
//     bound. A "refresh" pass recomputes it from the current model in FP32 (local quadratic
//     around the thread's first channel, the constant term split in two floats, the terms of
//     order >= 3 -- ~1e-5 of the model's variation over 32 channels -- in the band); a "stale" pass
//     reuses the stored value and widens the band by a bound on how far the model has moved since
//     (maximum over 32 Chebyshev nodes, Ehlich-Zeller inequality). A channel whose r~ lies inside
//     both thresholds by more than the band is provably not clipped and costs ~3-10 FP32
//     instructions; only the others (the clipped channels and a handful of near-threshold ones)
//     are evaluated exactly (FP64 Horner, float residual, baseline.cc:1088 predicate);
//   * an exactly evaluated residual that falls inside a threshold interval, a degenerate row, a
//     non-finite moment, or a normal-equation matrix that is not safely positive definite makes the
//     kernel give the ROW up: it is appended to a list and fitted by PolyFitKernel afterwards
//     (~1e-4..1e-3 of the rows). A row this kernel finishes has exactly the clip decisions the
//     per-channel evaluation would have taken.
//   * the last fit's residual, best-fit model and statistics (the reported rms) are evaluated
//     exactly, once, in FP64 as PolyFitKernel's pass A does.
//
// Layout: one CTA of NT threads per spectrum of n = 32 NT channels, persistent grid, 6 CTAs per SM.
// The row is loaded with coalesced 128-bit loads (thread t: float4 groups t, t+NT, ...; the power
// sums of the flagged channels are taken in that interleaved ownership, which spreads flagged runs
// evenly over the lanes), then every thread owns 32 CONSECUTIVE channels (XOR-swizzled rows in
// shared memory: conflict-free 128-bit accesses in both patterns) so that the local polynomial
// variable runs over 0..31 and FP32 accuracy suffices for the screening value. One warp per row
// (rotating) runs the serial part of an iteration -- downdate, solve, bounds, thresholds -- between
// two block barriers; everything per channel is a rolled loop (instruction-cache footprint).
#include "lsqfit_poly.cuh"
#include "device_common.cuh"
#include "poly_device.cuh"

#include <float.h>
#include <stdlib.h>

namespace sakura_b200 {

namespace {

constexpr int kListCap = 8; // clipped-channel entries per warp (the light path; >= kHeavyClipped)
constexpr double kU32 = 5.9604644775390625e-8; // 2^-24: unit roundoff of float
constexpr double kSlack = 1.0009765625; // 1 + 2^-10
constexpr double kTiny40 = 9.094947017729282e-13; // 2^-40
constexpr double kTiny44 = 5.684341886080802e-14; // 2^-44
constexpr double kTiny50 = 8.881784197001252e-16; // 2^-50
// a stale pass is used while the model has moved by less than this fraction of the clip threshold
constexpr double kStaleFraction = 0.15;
constexpr double kOutside = 1.00000095367431640625; // 1 + 2^-20 >= 1 + 4u: relative part of the r~ error
// a warp whose threads hold this many undecided channels / that clipped more than this many channels
// in one pass switches from per-thread bit loops to the vectorised / redistributed variants
constexpr int kHeavyCandidates = 8;
constexpr int kHeavyClipped = 8;

enum ScreenMode : int {
	kModeRefresh = 0, kModeStale = 1, kModeFinal = 2, kModeGiveUp = 3
};

template<int K, int NT>
struct ScreenShared {
	static constexpr int NW = NT / 32;
	static constexpr int NS = 2 * K - 1;
	static constexpr int NV = 3 * K; // S (2K-1 power sums), T (K data moments), Q (sum y^2)
	double wtot[NW][NV]; // per-warp moment totals of the channels flagged / clipped in this pass
	double st[NV]; // current S, T, Q over the valid channels (owned by the control warp)
	double stat[NW][2]; // final pass: sum, sum of squares per warp
	double2 list[NW][kListCap]; // (x, y) of the channels a warp clipped in this pass
	// ---- control block: written by the control warp, read by everybody after a barrier
	double cz[K]; // model coefficients in the channel index z
	double cx[K]; // ... in x = z/(n-1)
	double cxs[K]; // coefficients of the model the stored r~ refers to
	double kappa, ikappa; // rms of the previous fit and its reciprocal (scale of the next fit's bounds)
	double c1s; // sum |cxs_k|
	double es_block; // 2^-40 (A + C1) of the stored model
	double ebase; // thread-independent part of the screening band
	float lo_m, lo_p, hi_m, hi_p; // interval ends of the lower / upper clip threshold
	int mode;
	int icnt[NW];
	int wclip[NW];
	int giveup[2]; // [iteration parity] set by any thread that met an undecidable channel
};

template<int K, int NT>
__host__ __device__ constexpr size_t ScreenResidualBytes() {
	// r~ rows; doubles as the prologue's reduction scratch (NW x NS x 33 doubles)
	size_t const need = static_cast<size_t>(NT / 32) * (2 * K - 1) * 33 * sizeof(double);
	size_t const res = static_cast<size_t>(NT) * 32 * 4;
	return ((res > need ? res : need) + 15) / 16 * 16;
}

template<int K, int NT>
__host__ __device__ constexpr size_t ScreenSharedBytes() {
	return static_cast<size_t>(NT) * 32 * 4 + ScreenResidualBytes<K, NT>() + static_cast<size_t>(NT) * 8
			+ sizeof(ScreenShared<K, NT>) + 16;
}

__device__ __forceinline__ double warp_max(double v) {
#pragma unroll
	for (int o = 16; o > 0; o >>= 1) {
		v = fmax(v, __shfl_xor_sync(0xffffffffu, v, o));
	}
	return v;
}

/** Shared-memory position (in float4 units) of slot q (4 channels) of thread t's 32 consecutive
 *  channels: rows of 8 float4, the slot XOR-swizzled with the row number so that both the coalesced
 *  pattern of the load / store passes (consecutive groups) and the per-thread pattern (one row per
 *  lane) touch 8 different 16-byte columns per quarter warp: conflict-free without padding. */
__device__ __forceinline__ int slot_index(int t, int q) {
	return 8 * t + (q ^ (t & 7));
}

/** ... of a single channel (in floats) */
__device__ __forceinline__ int channel_index(int t, int u) {
	return 32 * t + 4 * ((u >> 2) ^ (t & 7)) + (u & 3);
}

/** value if bit `pos` of `word` is set, else +0.0f: an AND with the sign-extended bit (a NaN/Inf under
 *  the mask must not leak, so this cannot be a multiplication by 0/1). */
__device__ __forceinline__ float keep_if_bit(float value, uint32_t word, int pos) {
	int m;
	asm("bfe.s32 %0, %1, %2, 1;" : "=r"(m) : "r"(word), "r"(pos));
	return __int_as_float(__float_as_int(value) & m);
}

/** word |= bit if lo < r < hi (two compares and one predicated OR; false for NaN). */
__device__ __forceinline__ void or_if_inside(uint32_t &word, float r, float lo, float hi, uint32_t bit) {
	asm("{\n\t.reg .pred p;\n\tsetp.gt.f32 p, %1, %2;\n\tsetp.lt.and.f32 p, %1, %3, p;\n\t"
			"@p or.b32 %0, %0, %4;\n\t}" : "+r"(word) : "f"(r), "f"(lo), "f"(hi), "r"(bit));
}

/** word |= bit if r < lo or r > hi (false for NaN). */
__device__ __forceinline__ void or_if_outside(uint32_t &word, float r, float lo, float hi, uint32_t bit) {
	asm("{\n\t.reg .pred p;\n\tsetp.lt.f32 p, %1, %2;\n\tsetp.gt.or.f32 p, %1, %3, p;\n\t"
			"@p or.b32 %0, %0, %4;\n\t}" : "+r"(word) : "f"(r), "f"(lo), "f"(hi), "r"(bit));
}

/** pw[k] = x^k, k < N, as a product tree (x^k = x^(k/2) x^(k - k/2)): N - 2 multiplications like the
 *  running product, but a dependent chain of log2 N instead of N of them (FP64 latency is 8 cycles) */
template<int N>
__device__ __forceinline__ void PowerTree(double x, double (&pw)[N]) {
	pw[0] = 1.0;
	if (N > 1) {
		pw[1] = x;
	}
#pragma unroll
	for (int k = 2; k < N; ++k) {
		pw[k] = pw[k >> 1] * pw[k - (k >> 1)];
	}
}

// Power sums sum x^k (k < 2K-1) over the channels whose bit is set in `bits`, in the INTERLEAVED
// ownership of the load pass (bit 4j+e of thread t <-> channel 4(t + NT j) + e); per-lane partials
// are transposed through `scratch` (NS rows of 33 doubles per warp) and row k is added up by lane k
// in a fixed order: deterministic, no atomics.
template<int K, int NT>
__device__ __forceinline__ void WarpFlaggedPowerSums(uint32_t bits, double zbase, double h, int lane,
		double *scratch, double *wtot) {
	constexpr int NS = 2 * K - 1;
	uint32_t const active = __ballot_sync(0xffffffffu, bits != 0u);
	if (active == 0u) {
		if (lane < NS) {
			wtot[lane] = 0.0;
		}
		return;
	}
	{
		double ps[NS];
#pragma unroll
		for (int k = 0; k < NS; ++k) {
			ps[k] = 0.0;
		}
		while (bits != 0u) {
			int const s = __ffs(bits) - 1;
			bits &= bits - 1u;
			uint32_t const off = static_cast<uint32_t>((s >> 2) * (4 * NT) + (s & 3));
			double const x = (zbase + u32_to_double(off)) * h;
			double pw[NS];
			PowerTree<NS>(x, pw);
#pragma unroll
			for (int k = 0; k < NS; ++k) {
				ps[k] += pw[k];
			}
		}
#pragma unroll
		for (int k = 0; k < NS; ++k) {
			scratch[k * 33 + lane] = ps[k]; // every lane writes (zeros without bits)
		}
	}
	__syncwarp();
	if (lane < NS) {
		// fixed order, four independent partial sums (a single running sum is a chain of 32 FP64 adds)
		double const *rowp = scratch + lane * 33;
		double t0 = 0.0, t1 = 0.0, t2 = 0.0, t3 = 0.0;
#pragma unroll
		for (int l = 0; l < 32; l += 4) {
			t0 += rowp[l];
			t1 += rowp[l + 1];
			t2 += rowp[l + 2];
			t3 += rowp[l + 3];
		}
		double const tot = (t0 + t1) + (t2 + t3);
		wtot[lane] = tot;
	}
	__syncwarp();
}

/** >= the maximum over the warp of a non-negative value: rounded up to float, whose bit patterns
 *  order like the values (a NaN orders above +Inf and survives), one REDUX instruction. */
__device__ __forceinline__ double warp_max_upper(double a) {
	unsigned const bits = __float_as_uint(__double2float_ru(a));
	return static_cast<double>(__uint_as_float(__reduce_max_sync(0xffffffffu, bits)));
}

/** 1/sqrt(v) for v in [1e-30, 1e30]; sqrt(v) = v * rsqrt_fast(v). FP64 hardware estimate
 *  (rsqrt.approx.ftz.f64: only the upper 32 bits of v are used, relative error e0 < 2^-20) and two
 *  Newton steps, each mapping an error e to 1.5 e^2: < 2^-39 after the first, at the FP64 rounding level
 *  (~2^-52) after the second. The threshold intervals' safety factors (1 -/+ 3 * 2^-40) need the result
 *  to better than 2^-41: that holds for any estimate better than 2^-11, i.e. with a margin of 2^9 on the
 *  hardware's accuracy (tools/rcp_check.cu on B200, [1e-30, 1e30]: estimate 2^-20.1, result 2.2e-16). */
__device__ __forceinline__ double rsqrt_fast(double v) {
	// FP64 hardware estimate (relative error ~2^-20; no float round trip: a conversion costs as much
	// as the estimate itself)
	double r;
	asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(v));
	double const hv = 0.5 * v;
	r = fma(r, fma(-hv * r, r, 0.5), r);
	r = fma(r, fma(-hv * r, r, 0.5), r);
	return r;
}

// One warp: apply the per-warp moment totals, solve, derive the control block of the next pass.
//   first      : the totals are the prologue's (flagged-channel power sums, data moments in z)
//   count      : number of valid channels the system is built from
//   want_clip  : false for the last fit (no thresholds needed)
//   have_stored: r~ in shared memory refers to the model sh.cxs
template<int K, int NT>
__device__ __forceinline__ void Control(ScreenShared<K, NT> &sh, PolyConst const &pc, bool first,
		bool complement, int lane, int count, float clip, bool want_clip, bool have_stored, int parity) {
	constexpr int NW = NT / 32;
	constexpr int NS = 2 * K - 1;
	constexpr int NV = 3 * K;
	if (lane < NV) {
		double tot = 0.0;
#pragma unroll
		for (int w = 0; w < NW; ++w) {
			tot += sh.wtot[w][lane];
		}
		double v;
		if (first) {
			if (lane < NS) {
				v = complement ? (pc.s_all[lane] - tot) : tot;
			} else if (lane < NS + K) {
				v = tot * pc.hpow[lane - NS];
			} else {
				v = tot;
			}
		} else {
			v = sh.st[lane] - tot;
		}
		sh.st[lane] = v;
	}
	__syncwarp();
	double S[NS];
	double T[K];
#pragma unroll
	for (int k = 0; k < NS; ++k) {
		S[k] = sh.st[k];
	}
#pragma unroll
	for (int k = 0; k < K; ++k) {
		T[k] = sh.st[NS + k];
	}
	double const Q = sh.st[NV - 1];
	double cx[K];
	bool const solved = SolveLDLT<K>(S, T, cx);
	double c1 = 0.0;
#pragma unroll
	for (int k = 0; k < K; ++k) {
		c1 += fabs(cx[k]);
	}
	// A >= max |model| and D >= max |model - stored model| over the row: maxima over 32 Chebyshev
	// nodes (one per lane) times the Ehlich-Zeller factor, plus the evaluation's own rounding
	double const xq = pc.node[lane];
	double mv = cx[K - 1];
	double dv = have_stored ? cx[K - 1] - sh.cxs[K - 1] : 0.0;
#pragma unroll
	for (int k = K - 2; k >= 0; --k) {
		mv = fma(mv, xq, cx[k]);
		dv = fma(dv, xq, have_stored ? cx[k] - sh.cxs[k] : 0.0);
	}
	mv = warp_max_upper(fabs(mv));
	dv = warp_max_upper(fabs(dv));
	double const A = fma(mv, pc.ez[K - 1], kTiny44 * c1);
	double const D = fma(dv, pc.ez[K - 1], kTiny44 * (c1 + sh.c1s));
	int mode = kModeGiveUp;
	double ebase = 0.0;
	float lo_m = 0.0f, lo_p = 0.0f, hi_m = 0.0f, hi_p = 0.0f;
	// every comparison below is written so that a NaN anywhere leaves mode == kModeGiveUp
	if (solved && c1 < DBL_MAX && A < DBL_MAX) {
		if (!want_clip) {
			mode = kModeFinal;
		} else {
			double const N = static_cast<double>(count);
			double const inv_n = fast_rcp(N);
			// real-arithmetic residual sums from the moments, with the moments' own FP64 error
			double sum_r = T[0];
			double ss = Q;
#pragma unroll
			for (int k = 0; k < K; ++k) {
				sum_r = fma(-cx[k], S[k], sum_r);
				double g = 0.0; // (G c)_k
#pragma unroll
				for (int j = 0; j < K; ++j) {
					g = fma(cx[j], S[j + k], g);
				}
				ss = fma(cx[k], g - 2.0 * T[k], ss);
			}
			// The bounds are kept free of square roots and divisions on the dependent chain (this warp is
			// what the other three wait for): products of the form a sqrt(N ss) are bounded by
			// (a^2 N + ss) / 2, and sum |rho| / N <= sqrt(ss / N) reuses the rms computed anyway.
			double const e64_ss = kTiny40 * 2.0 * (Q + c1 * c1 * N); // >= 2^-40 (Q + 2 c1 sum|y| + c1^2 N)
			double const ss_up = ss + e64_ss;
			double const ssc = (ss_up > 0.0) ? ss_up : 0.0; // (NaN -> 0: caught by the tests on q_m below)
			// what rounding the model and the residual to float can change (|e_i| <= u (A + |rho_i|)):
			// 2 u (A sum|rho| + ss) + 2 u^2 (N A^2 + ss). sum|rho| <= sqrt(N ss) <= (N kappa + ss / kappa) / 2
			// for any kappa > 0: the previous fit's rms (tight within a few per cent), A for the first fit
			double const a2n = A * A * N;
			double const a_sum_rho = first ? 0.5 * (a2n + ssc) : 0.5 * A * fma(N, sh.kappa, ssc * sh.ikappa);
			double const d_ss = 2.0 * kU32 * kSlack * (a_sum_rho + ssc) + 2.0 * kU32 * kU32 * (a2n + ssc) + e64_ss
					+ kTiny44 * ssc;
			double const q_m = (ss - d_ss) * inv_n; // (1/N is good to an ulp or two: absorbed by the nudges)
			double const q_p = (ss + d_ss) * inv_n;
			if (q_m > 1e-30 && q_p < 1e30 && clip > 0.0f) {
				double const r_m = rsqrt_fast(q_m);
				double const r_p = rsqrt_fast(q_p);
				double const s_p = q_p * r_p * (1.0 + kTiny40); // >= sqrt(ss/N) >= sum|rho| / N
				if (lane == 0) {
					sh.kappa = s_p;
					sh.ikappa = r_p;
				}
				// |mean| <= |sum r| / N + u (A + sum|rho| / N) + 2^-40 (sum|y| / N + c1), sum|y| <= N A + sum|rho|
				double const d_mean = kU32 * kSlack * (A + s_p) + kTiny40 * (A + s_p + c1);
				double const mean_c = sum_r * inv_n;
				double const mean_m = mean_c - d_mean;
				double const mean_p = mean_c + d_mean;
				double const mabs = fabs(mean_c) + d_mean;
				double const m2max = mabs * mabs;
				// var = q - mean^2 in [q_m - m2max, q_p]; sqrt(q_m - m2) >= sqrt(q_m) - m2 / sqrt(q_m)
				double const sigma = static_cast<double>(clip);
				// (the safety factors of rms_m and of the product with sigma are folded into sigma_m / sigma_p)
				double const rms_m = fma(-m2max * (1.0 + kTiny40), r_m, q_m * r_m);
				double const rms_p = s_p;
				// thresholds as make_thresholds (device_common.cuh) forms them; conversions to float
				// round to nearest and are monotone, the FP64 roundings are covered by the nudges
				float const thr_m = static_cast<float>(rms_m * (sigma * (1.0 - 3.0 * kTiny40)));
				float const thr_p = static_cast<float>(rms_p * (sigma * (1.0 + 3.0 * kTiny40)));
				double const nudge = kTiny40 * (mabs + static_cast<double>(thr_p));
				lo_m = static_cast<float>(mean_m - static_cast<double>(thr_p) - nudge);
				lo_p = static_cast<float>(mean_p - static_cast<double>(thr_m) + nudge);
				hi_m = static_cast<float>(mean_m + static_cast<double>(thr_m) - nudge);
				hi_p = static_cast<float>(mean_p + static_cast<double>(thr_p) + nudge);
				// (r-lo)*(hi-r) < 0 (baseline.cc:1088) is (r < lo) || (r > hi) when the product cannot
				// underflow: thresholds not tiny, window not degenerate (see PolyFitKernel pass B)
				// (the whole interval of the lower threshold negative, of the upper one positive)
				bool const by_compare = lo_p <= -0x1p-80f && hi_m >= 0x1p-80f
						&& fabsf(lo_m) <= FLT_MAX && fabsf(hi_p) <= FLT_MAX
						&& __fsub_rn(hi_m, lo_p) >= 0x1p-40f;
				if (by_compare) {
					// lo_m <= lo_p < 0 < hi_m <= hi_p here
					double const theta = static_cast<double>(fmaxf(-lo_m, hi_p));
					if (have_stored && D <= kStaleFraction * static_cast<double>(thr_m)) {
						mode = kModeStale;
						ebase = kU32 * kSlack * (A + 6.0 * (theta + D)) + kTiny40 * (A + c1) + sh.es_block + D;
					} else {
						mode = kModeRefresh;
						ebase = kU32 * kSlack * (A + 6.0 * theta) + kTiny40 * (A + c1);
						if (lane < K) {
							sh.cxs[lane] = cx[lane];
						}
						if (lane == 0) {
							sh.c1s = c1;
							sh.es_block = kTiny40 * (A + c1);
						}
					}
				}
			}
		}
	}
	if (lane == 0) {
#pragma unroll
		for (int k = 0; k < K; ++k) {
			sh.cx[k] = cx[k];
			sh.cz[k] = cx[k] * pc.hpow[k];
		}
		sh.mode = mode;
		sh.ebase = ebase;
		sh.lo_m = lo_m;
		sh.lo_p = lo_p;
		sh.hi_m = hi_m;
		sh.hi_p = hi_p;
		sh.giveup[parity] = 0; // last read two iterations (two barriers) ago
	}
}

template<int K, int NT, bool CAL>
__global__ void __launch_bounds__(NT, (NT == 128 ? 6 : 3))
PolyScreenFitKernel(PolyFitParams const p, PolyConst const pc, int const num_sms) {
	constexpr int NW = NT / 32;
	constexpr int NS = 2 * K - 1;
	constexpr int NV = 3 * K;
	constexpr int n = 32 * NT;
	extern __shared__ __align__(16) unsigned char smem_raw[];
	// layout: data rows [NT][32] | r~ rows [NT][32] (also prologue scratch) | mask nibbles | control
	float *sdata = reinterpret_cast<float *>(smem_raw);
	float *sres = reinterpret_cast<float *>(smem_raw + static_cast<size_t>(NT) * 32 * 4);
	uint8_t *snib = smem_raw + static_cast<size_t>(NT) * 32 * 4 + ScreenResidualBytes<K, NT>();
	ScreenShared<K, NT> &sh = *reinterpret_cast<ScreenShared<K, NT> *>(snib + static_cast<size_t>(NT) * 8);

	int const tid = threadIdx.x;
	int const lane = tid & 31;
	int const warp = tid >> 5;
	double *scratch = reinterpret_cast<double *>(sres) + warp * (NS * 33);
	double const zbase_il = static_cast<double>(4 * tid); // interleaved ownership (load pass)
	float const clip = p.clip_threshold_sigma;
	// consecutive ownership: channels 32 tid + u, u < 32; slot q (channels 4q..4q+3) at own_*[q ^ t7]
	float4 *own_data = reinterpret_cast<float4 *>(sdata) + 8 * tid;
	float4 *own_res = reinterpret_cast<float4 *>(sres) + 8 * tid;
	int const t7 = tid & 7;
	uint32_t const examine = (tid == NT - 1) ? 0x7fffffffu : 0xffffffffu; // channel n-1 is never clipped

	// The per-channel loops below are deliberately NOT unrolled over the eight float4 groups of a
	// thread and use no per-channel immediates: the unrolled form of this kernel (70 KB of SASS)
	// ran at a 69 % instruction-cache hit rate with the GPC instruction cache at 93 % of its
	// request peak (profiles/r01/screen_v2_ncu_summary.txt); the loop bodies fit the caches.
	// The warp that runs Control (a long dependent FP64 chain) rotates from row to row and differs
	// between the CTAs resident on one SM: warp w always issues on scheduler w % 4, and with a fixed
	// control warp every CTA of the SM would queue its chain on the same scheduler.
	int cwarp = (blockIdx.x / num_sms) % NW;
	for (size_t row = blockIdx.x; row < p.num_spectra; row += gridDim.x, cwarp = (cwarp + 1) % NW) {
		float4 const *gdata = reinterpret_cast<float4 const *>(p.data + row * p.data_stride);
		uint32_t const *gmask = reinterpret_cast<uint32_t const *>(p.mask + row * p.mask_stride);
		if (tid == 0 && row + gridDim.x < p.num_spectra) {
			size_t const nrow = row + gridDim.x;
			asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;"
					:: "l"(p.data + nrow * p.data_stride), "r"(static_cast<unsigned>(n) * 4u) : "memory");
			asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;"
					:: "l"(p.mask + nrow * p.mask_stride), "r"(static_cast<unsigned>(n)) : "memory");
		}
		__syncthreads(); // the previous row's shared state is dead

		// ---- load the row: coalesced, interleaved ownership (thread t: groups t, t + NT, ...) --------
		uint32_t vbits_il = 0u;
#pragma unroll
		for (int j = 0; j < 8; ++j) {
			int const g = tid + NT * j;
			float4 y = __ldg(gdata + g);
			uint32_t m = __ldg(gmask + g);
			m = __vcmpne4(m, 0u) & 0x01010101u; // any non-zero byte is "true"
			if (CAL) { // calibration fused into the load
				float4 const r = __ldg(reinterpret_cast<float4 const *>(p.cal.reference
						+ row * p.cal.reference_stride) + g);
				float4 s = make_float4(p.cal.const_scaling, p.cal.const_scaling,
						p.cal.const_scaling, p.cal.const_scaling);
				if (p.cal.scaling != nullptr) {
					s = __ldg(reinterpret_cast<float4 const *>(p.cal.scaling
							+ row * p.cal.scaling_stride) + g);
				}
				y.x = CalibrateOne(s.x, y.x, r.x);
				y.y = CalibrateOne(s.y, y.y, r.y);
				y.z = CalibrateOne(s.z, y.z, r.z);
				y.w = CalibrateOne(s.w, y.w, r.w);
				if (p.cal.flag_nonfinite) { // sakura_SetFalseIfNanOrInfFloat merged into the mask
					auto finite = [](float v) {
						return (__float_as_uint(v) & 0x7f800000u) != 0x7f800000u;
					};
					m &= (finite(y.x) ? 0x1u : 0u) | (finite(y.y) ? 0x100u : 0u)
							| (finite(y.z) ? 0x10000u : 0u) | (finite(y.w) ? 0x1000000u : 0u);
				}
			}
			// group g belongs to thread g/8 of the consecutive ownership, slot g%8
			reinterpret_cast<float4 *>(sdata)[slot_index(g >> 3, g & 7)] = y;
			uint32_t const nib = (m | (m >> 7) | (m >> 14) | (m >> 21)) & 0xfu;
			snib[g] = static_cast<uint8_t>(nib);
			vbits_il |= nib << (4 * j);
		}
		int const cnt = warp_sum(__popc(vbits_il));
		if (lane == 0) {
			sh.icnt[warp] = cnt;
		}
		__syncthreads();
		int count = 0;
#pragma unroll
		for (int w = 0; w < NW; ++w) {
			count += sh.icnt[w];
		}
		if (p.num_fitting_max == 0) { // baseline.cc:1295-1296: kOK, nothing written
			if (tid == 0) {
				p.status[row] = kStatusOK;
				p.lsqfit_status[row] = kLsqOK;
			}
			continue;
		}
		if (count < K) { // baseline.cc:1147-1150
			if (tid == 0) {
				p.status[row] = kStatusNG;
				p.lsqfit_status[row] = kLsqNotEnoughData;
			}
			continue;
		}
		// valid bits of this thread's 32 consecutive channels: 8 nibbles, one per byte
		uint32_t vbits;
		{
			uint2 const w = reinterpret_cast<uint2 const *>(snib)[tid];
			uint32_t a = w.x | (w.x >> 4); // bytes 0 and 2 hold nibble pairs (0,1) and (2,3)
			a = (a & 0xffu) | ((a >> 8) & 0xff00u);
			uint32_t b = w.y | (w.y >> 4);
			b = (b & 0xffu) | ((b >> 8) & 0xff00u);
			vbits = a | (b << 16);
		}
		// ---- data moments over the thread's 32 consecutive channels -------------------------------
		// Running sums of running sums, channels in descending order: s_k = sum_u y_u C(u + k, k)
		// (additions only, no per-channel constants); converted to sum y_u u^k with the integer matrix
		// pc.binom (u^j = sum_k binom[j][k] C(u + k, k)) and shifted to the global channel index.
		double t[K];
		double qsum = 0.0;
		{
			double s[K];
#pragma unroll
			for (int k = 0; k < K; ++k) {
				s[k] = 0.0;
			}
#pragma unroll 1
			for (int q = 7; q >= 0; --q) {
				float4 const y = own_data[q ^ t7];
				float const ye[4] = { y.x, y.y, y.z, y.w };
				uint32_t const nib = vbits >> (4 * q);
#pragma unroll
				for (int e = 3; e >= 0; --e) {
					// select (on the bits), not multiply: a NaN/Inf under the mask must not leak
					double const yd = static_cast<double>(keep_if_bit(ye[e], nib, e));
					s[0] += yd;
#pragma unroll
					for (int k = 1; k < K; ++k) {
						s[k] += s[k - 1];
					}
					qsum = fma(yd, yd, qsum);
				}
			}
#pragma unroll
			for (int j = 0; j < K; ++j) {
				double v = 0.0;
#pragma unroll
				for (int k = 0; k <= j; ++k) {
					v = fma(pc.binom[j][k], s[k], v);
				}
				t[j] = v;
			}
			double const zb = static_cast<double>(32 * tid);
#pragma unroll
			for (int i = 0; i < K - 1; ++i) {
#pragma unroll
				for (int j = K - 1; j > i; --j) {
					t[j] = fma(zb, t[j - 1], t[j]);
				}
			}
		}
		// S over the valid channels = constant over all channels - sum over the flagged ones
		// (or the direct sum when more than half of the row is flagged); interleaved ownership,
		// which spreads flagged runs over the lanes
		bool const complement = (2 * count >= n);
		WarpFlaggedPowerSums<K, NT>(complement ? ~vbits_il : vbits_il, zbase_il, pc.h, lane, scratch,
				sh.wtot[warp]);
#pragma unroll
		for (int k = 0; k < K; ++k) {
			t[k] = warp_sum(t[k]);
		}
		qsum = warp_sum(qsum);
		if (lane == 0) {
#pragma unroll
			for (int k = 0; k < K; ++k) {
				sh.wtot[warp][NS + k] = t[k];
			}
			sh.wtot[warp][NV - 1] = qsum;
		}
		__syncthreads();

		// ---- clip iterations --------------------------------------------------------------------
		int num_unmasked = n; // baseline.cc:1152 (sic)
		bool failed = false;
		bool giveup = false;
		bool have_stored = false;
		float vs_t = 0.0f; // 16 u V_t of the stored r~ (this thread's share of its error bound)
		for (int iter = 1; iter <= p.num_fitting_max; ++iter) {
			if (num_unmasked < K) { // baseline.cc:1162-1165
				failed = true;
				break;
			}
			if (warp == cwarp) { // moments of all warps are visible (barrier above / at the end of the pass)
				Control<K, NT>(sh, pc, iter == 1, complement, lane, count, clip, iter < p.num_fitting_max,
						have_stored, iter & 1);
			}
			__syncthreads(); // control block of this iteration is visible
			int const mode = sh.mode;
			if (mode == kModeGiveUp) {
				giveup = true;
				break;
			}
			if (mode == kModeFinal) {
				break;
			}
			float const lo_m = sh.lo_m, lo_p = sh.lo_p, hi_m = sh.hi_m, hi_p = sh.hi_p;
			double const ebase = sh.ebase;
			uint32_t inside = 0u; // channels proven to lie strictly inside both thresholds
			float c_hi, c_lo; // r~ strictly between these: inside whatever the exact thresholds are
			float o_hi, o_lo; // r~ beyond these: outside whatever the exact thresholds are
			if (mode == kModeRefresh) {
				double d[K];
				{
					double cz[K];
#pragma unroll
					for (int k = 0; k < K; ++k) {
						cz[k] = sh.cz[k];
					}
					TaylorShift<K>(cz, static_cast<double>(32 * tid), d);
				}
				float const dhi = static_cast<float>(d[0]);
				float const dlo = static_cast<float>(d[0] - static_cast<double>(dhi));
				// Over 32 channels the terms of third and higher order of the local expansion are of the
				// order of 1e-5 of the model's variation: they are not evaluated per channel but added to
				// this thread's band (V3), which costs a fraction of a candidate per row and pass.
				constexpr int KS = (K < 3) ? K : 3; // local terms evaluated per channel: 1, u, u^2
				float df[KS];
				double V = 0.0, V3 = 0.0;
				{
					double upow = 1.0;
#pragma unroll
					for (int k = 1; k < K; ++k) {
						upow *= 31.0;
						if (k < KS) {
							df[k] = static_cast<float>(d[k]);
							V = fma(fabs(d[k]), upow, V);
						} else {
							V3 = fma(fabs(d[k]), upow, V3);
						}
					}
				}
				vs_t = __double2float_ru(V * (16.0 * kU32 * kSlack) + V3 * kSlack);
				double const et = ebase + static_cast<double>(vs_t);
				c_hi = __double2float_rd(static_cast<double>(hi_m) - et);
				c_lo = __double2float_ru(static_cast<double>(lo_p) + et);
				o_hi = __double2float_ru((static_cast<double>(hi_p) + et) * kOutside);
				o_lo = __double2float_rd((static_cast<double>(lo_m) - et) * kOutside);
				float ub = 0.0f; // local channel index of the group's first channel (exact in float)
#pragma unroll 2
				for (int q = 0; q < 8; ++q) {
					float4 const y = own_data[q ^ t7];
					float const ye[4] = { y.x, y.y, y.z, y.w };
					float re[4];
					uint32_t nib = 0u;
#pragma unroll
					for (int e = 0; e < 4; ++e) {
						float delta = dlo;
						if (KS > 1) {
							float const uf = __fadd_rn(ub, static_cast<float>(e));
							float acc = df[KS - 1];
#pragma unroll
							for (int k = KS - 2; k >= 1; --k) {
								acc = __fmaf_rn(acc, uf, df[k]);
							}
							delta = __fmaf_rn(acc, uf, dlo);
						}
						float const r = __fsub_rn(__fsub_rn(ye[e], dhi), delta);
						re[e] = r;
						or_if_inside(nib, r, c_lo, c_hi, 1u << e);
					}
					own_res[q ^ t7] = make_float4(re[0], re[1], re[2], re[3]);
					inside |= nib << (4 * q);
					ub = __fadd_rn(ub, 4.0f);
				}
				have_stored = true;
			} else {
				double const et = ebase + static_cast<double>(vs_t);
				c_hi = __double2float_rd(static_cast<double>(hi_m) - et);
				c_lo = __double2float_ru(static_cast<double>(lo_p) + et);
				o_hi = __double2float_ru((static_cast<double>(hi_p) + et) * kOutside);
				o_lo = __double2float_rd((static_cast<double>(lo_m) - et) * kOutside);
#pragma unroll 2
				for (int q = 0; q < 8; ++q) {
					float4 const r4 = own_res[q ^ t7];
					uint32_t nib = 0u;
					or_if_inside(nib, r4.x, c_lo, c_hi, 1u);
					or_if_inside(nib, r4.y, c_lo, c_hi, 2u);
					or_if_inside(nib, r4.z, c_lo, c_hi, 4u);
					or_if_inside(nib, r4.w, c_lo, c_hi, 8u);
					inside |= nib << (4 * q);
				}
			}
			// ---- the channels the screen could not clear ----------------------------------------
			// r~ beyond the outer band: clipped for certain. Otherwise evaluate exactly (FP64 Horner,
			// float residual): outside both threshold intervals -> clipped, inside both -> kept,
			// in between -> the row cannot be decided here.
			uint32_t cand = ~inside & vbits & examine;
			uint32_t clipped = 0u;
			if (__any_sync(0xffffffffu, __popc(cand) >= kHeavyCandidates)) {
				// a line covers whole threads: one vectorised pass instead of long per-thread bit loops
				uint32_t outside = 0u;
#pragma unroll 2
				for (int q = 0; q < 8; ++q) {
					float4 const r4 = own_res[q ^ t7];
					uint32_t nib = 0u;
					or_if_outside(nib, r4.x, o_lo, o_hi, 1u);
					or_if_outside(nib, r4.y, o_lo, o_hi, 2u);
					or_if_outside(nib, r4.z, o_lo, o_hi, 4u);
					or_if_outside(nib, r4.w, o_lo, o_hi, 8u);
					outside |= nib << (4 * q);
				}
				clipped = cand & outside;
				cand &= ~outside;
			}
			while (cand != 0u) {
				int const s = __ffs(cand) - 1;
				cand &= cand - 1u;
				float const rt = sres[channel_index(tid, s)];
				if (rt > o_hi || rt < o_lo) {
					clipped |= 1u << s;
					continue;
				}
				float const y = sdata[channel_index(tid, s)];
				double const z = static_cast<double>(32 * tid + s);
				double acc = sh.cz[K - 1];
#pragma unroll
				for (int k = K - 2; k >= 0; --k) {
					acc = fma(acc, z, sh.cz[k]);
				}
				float const r = __fsub_rn(y, static_cast<float>(acc));
				if (r > hi_p || r < lo_m) {
					clipped |= 1u << s; // outside whatever the exact thresholds are
				} else if (!(r <= hi_m && r >= lo_p)) {
					sh.giveup[iter & 1] = 1; // inside a threshold interval (or NaN): undecidable here
				}
			}
			vbits &= ~clipped;
			// ---- moments of the clipped channels (downdate, numeric_operation.cc:283-314, 448-479) ----
			int const mine = __popc(clipped);
			int const total = __reduce_add_sync(0xffffffffu, mine);
			if (total > kHeavyClipped) {
				// Many channels, typically runs that fill a few threads' words: the warp's 1024 clipped
				// bits are redistributed by nibbles (lane l takes nibbles l, l+32, ..., l+224, so a run of
				// 100 channels spreads over 25 lanes, four channels each), every lane accumulates private
				// partial moments, and the partials are added up with a butterfly of shuffles.
				uint32_t *cw = reinterpret_cast<uint32_t *>(sh.list[warp]);
				cw[lane] = clipped;
				__syncwarp();
				uint8_t const *cb = reinterpret_cast<uint8_t const *>(cw);
				uint32_t w = 0u; // bit 4 m + e <-> channel 4 (lane + 32 m) + e of this warp
#pragma unroll
				for (int m = 0; m < 8; ++m) {
					int const nibble = lane + 32 * m;
					w |= ((static_cast<uint32_t>(cb[nibble >> 1]) >> (4 * (nibble & 1))) & 0xfu) << (4 * m);
				}
				__syncwarp();
				double ps[NS];
				double pt[K];
				double qq = 0.0;
#pragma unroll
				for (int k = 0; k < NS; ++k) {
					ps[k] = 0.0;
				}
#pragma unroll
				for (int k = 0; k < K; ++k) {
					pt[k] = 0.0;
				}
				while (w != 0u) {
					int const s = __ffs(w) - 1;
					w &= w - 1u;
					int const ch = 1024 * warp + 4 * (lane + 32 * (s >> 2)) + (s & 3);
					double const x = static_cast<double>(ch) * pc.h;
					double const yd = static_cast<double>(sdata[channel_index(ch >> 5, ch & 31)]);
					double pw[NS];
					PowerTree<NS>(x, pw);
#pragma unroll
					for (int k = 0; k < NS; ++k) {
						ps[k] += pw[k];
						if (k < K) {
							pt[k] = fma(yd, pw[k], pt[k]);
						}
					}
					qq = fma(yd, yd, qq);
				}
#pragma unroll
				for (int k = 0; k < NS; ++k) {
					ps[k] = warp_sum(ps[k]);
				}
#pragma unroll
				for (int k = 0; k < K; ++k) {
					pt[k] = warp_sum(pt[k]);
				}
				qq = warp_sum(qq);
				if (lane == 0) {
#pragma unroll
					for (int k = 0; k < NS; ++k) {
						sh.wtot[warp][k] = ps[k];
					}
#pragma unroll
					for (int k = 0; k < K; ++k) {
						sh.wtot[warp][NS + k] = pt[k];
					}
					sh.wtot[warp][NV - 1] = qq;
				}
			} else {
				// Few channels: (x, y) of the clipped channels are listed per warp in channel order;
				// lane k then accumulates sum x^k and sum y x^k over the list (lane 0 also sum y^2):
				// a fixed order, no atomics, no reduction scratch.
				double s_acc = 0.0, t_acc = 0.0, q_acc = 0.0;
				if (total > 0) {
					// position of this lane's first entry: one ballot when no lane holds more than one
					// channel (the usual case late in a fit), else a prefix sum over the lanes
					int first_entry;
					if (__all_sync(0xffffffffu, mine <= 1)) {
						first_entry = __popc(__ballot_sync(0xffffffffu, mine != 0) & ((1u << lane) - 1u));
					} else {
						int incl = mine;
#pragma unroll
						for (int o = 1; o < 32; o <<= 1) {
							int const v = __shfl_up_sync(0xffffffffu, incl, o);
							incl += (lane >= o) ? v : 0;
						}
						first_entry = incl - mine;
					}
					{
						uint32_t b = clipped;
						int r = first_entry;
						while (b != 0u) {
							int const s = __ffs(b) - 1;
							b &= b - 1u;
							double const x = static_cast<double>(32 * tid + s) * pc.h;
							sh.list[warp][r] = make_double2(x, static_cast<double>(sdata[channel_index(tid, s)]));
							++r;
						}
					}
					__syncwarp();
					for (int i = 0; i < total; ++i) {
						double2 const xy = sh.list[warp][i];
						double const x2 = xy.x * xy.x;
						double const x4 = x2 * x2;
						double const x8 = x4 * x4;
						// x^lane from the binary digits of the lane number, as a two-level product
						double const f01 = ((lane & 1) ? xy.x : 1.0) * ((lane & 2) ? x2 : 1.0);
						double const f23 = ((lane & 4) ? x4 : 1.0) * ((lane & 8) ? x8 : 1.0);
						double const pw = f01 * f23;
						s_acc += pw;
						t_acc = fma(xy.y, pw, t_acc);
						q_acc = fma(xy.y, xy.y, q_acc);
					}
					__syncwarp();
				}
				if (lane < NS) {
					sh.wtot[warp][lane] = s_acc;
				}
				if (lane < K) {
					sh.wtot[warp][NS + lane] = t_acc;
				}
				if (lane == 0) {
					sh.wtot[warp][NV - 1] = q_acc;
				}
			}
			if (lane == 0) {
				sh.wclip[warp] = total;
			}
			__syncthreads();
			if (sh.giveup[iter & 1] != 0) {
				giveup = true;
				break;
			}
			int num_clipped = 0;
#pragma unroll
			for (int w = 0; w < NW; ++w) {
				num_clipped += sh.wclip[w];
			}
			if (num_clipped == 0) {
				break; // converged (baseline.cc:1211-1213); sh.cz is the model of the last fit
			}
			num_unmasked = count - num_clipped;
			count = num_unmasked;
		}
		if (giveup) {
			// PolyFitKernel fits this row afterwards; nothing of it has been written
			if (tid == 0) {
				int const pos = atomicAdd(p.fallback_count, 1);
				p.fallback_rows[pos] = static_cast<int32_t>(row);
			}
			continue;
		}

		// ---- final mask (bits -> bytes), 32 consecutive bytes per thread ------------------------------
		{
			uint32_t w[8];
#pragma unroll
			for (int q = 0; q < 8; ++q) {
				w[q] = (((vbits >> (4 * q)) & 0xfu) * 0x00204081u) & 0x01010101u;
			}
			uint4 *gfm = reinterpret_cast<uint4 *>(p.final_mask + row * p.mask_stride) + 2 * tid;
			gfm[0] = make_uint4(w[0], w[1], w[2], w[3]);
			gfm[1] = make_uint4(w[4], w[5], w[6], w[7]);
		}
		if (failed) {
			// the reference has already clipped final_mask in place (SURVEY A-13 ii)
			if (tid == 0) {
				p.status[row] = kStatusNG;
				p.lsqfit_status[row] = kLsqNotEnoughData;
				if (p.flags != nullptr) {
					p.flags[row] = 2;
				}
			}
			continue;
		}
		// ---- the last fit, exactly: FP64 model -> float, float residual, FP64 statistics over the
		// valid channels (PolyFitKernel pass A in the consecutive ownership) -------------------------
		double sum = 0.0, sq = 0.0;
		{
			double d[K];
			{
				double cz[K];
#pragma unroll
				for (int k = 0; k < K; ++k) {
					cz[k] = sh.cz[k];
				}
				TaylorShift<K>(cz, static_cast<double>(32 * tid), d);
			}
			bool const want_best_fit = p.best_fit != nullptr;
			double ub = 0.0;
#pragma unroll 2
			for (int q = 0; q < 8; ++q) {
				float4 const y = own_data[q ^ t7];
				float const ye[4] = { y.x, y.y, y.z, y.w };
				float re[4], me[4];
				uint32_t const nib = vbits >> (4 * q);
#pragma unroll
				for (int e = 0; e < 4; ++e) {
					double const u = ub + static_cast<double>(e);
					double acc = d[K - 1];
#pragma unroll
					for (int k = K - 2; k >= 0; --k) {
						acc = fma(acc, u, d[k]);
					}
					me[e] = static_cast<float>(acc);
					float const r = __fsub_rn(ye[e], me[e]);
					re[e] = r;
					double const rd = static_cast<double>(keep_if_bit(r, nib, e));
					sum += rd;
					sq = fma(rd, rd, sq);
				}
				own_res[q ^ t7] = make_float4(re[0], re[1], re[2], re[3]);
				if (want_best_fit) {
					own_data[q ^ t7] = make_float4(me[0], me[1], me[2], me[3]);
				}
				ub += 4.0;
			}
		}
		sum = warp_sum(sum);
		sq = warp_sum(sq);
		if (lane == 0) {
			sh.stat[warp][0] = sum;
			sh.stat[warp][1] = sq;
		}
		__syncthreads();
		if (p.residual != nullptr) {
			float4 *outr = reinterpret_cast<float4 *>(p.residual + row * p.data_stride);
#pragma unroll
			for (int j = 0; j < 8; ++j) {
				int const g = tid + NT * j;
				outr[g] = reinterpret_cast<float4 const *>(sres)[slot_index(g >> 3, g & 7)];
			}
		}
		if (p.best_fit != nullptr) {
			float4 *outb = reinterpret_cast<float4 *>(p.best_fit + row * p.data_stride);
#pragma unroll
			for (int j = 0; j < 8; ++j) {
				int const g = tid + NT * j;
				outb[g] = reinterpret_cast<float4 const *>(sdata)[slot_index(g >> 3, g & 7)];
			}
		}
		if (p.coeff != nullptr && tid < K) {
			if (p.chebyshev) {
				double a = 0.0;
				for (int k = K - 1; k >= tid; --k) {
					a = fma(pc.cheb[tid][k], sh.cx[k], a);
				}
				p.coeff[row * K + tid] = a;
			} else {
				p.coeff[row * K + tid] = sh.cx[tid] / pc.factor[tid];
			}
		}
		if (tid == 0) {
			sum = 0.0;
			sq = 0.0;
#pragma unroll
			for (int w = 0; w < NW; ++w) {
				sum += sh.stat[w][0];
				sq += sh.stat[w][1];
			}
			ClipThresholds const th = make_thresholds(count, sum, sq, clip);
			p.rms[row] = static_cast<float>(th.rms);
			p.status[row] = kStatusOK;
			p.lsqfit_status[row] = kLsqOK;
			if (p.flags != nullptr) {
				p.flags[row] = 3;
			}
		}
	}
}

template<int K, int NT, bool CAL>
cudaError_t LaunchScreenKNC(PolyFitParams const &p, int num_sms, cudaStream_t stream) {
	constexpr size_t smem = ScreenSharedBytes<K, NT>();
	static int per_sm = 0;
	// (per device: a process may drive several GPUs)
	cudaError_t err = cudaFuncSetAttribute(PolyScreenFitKernel<K, NT, CAL>,
			cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem));
	if (err != cudaSuccess) {
		return err;
	}
	if (per_sm == 0) {
		int v = 0;
		err = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&v, PolyScreenFitKernel<K, NT, CAL>, NT, smem);
		if (err != cudaSuccess) {
			return err;
		}
		if (v < 1) {
			return cudaErrorInvalidConfiguration;
		}
		per_sm = v;
	}
	size_t grid = static_cast<size_t>(num_sms) * per_sm;
	{
		// tuning knob: resident CTAs per SM (persistent grid = #SM x this), at most the occupancy limit
		char const *e = getenv("SAKURA_B200_SCREEN_CTAS");
		int const v = (e != nullptr) ? atoi(e) : 0;
		if (v >= 1 && v < per_sm) {
			grid = static_cast<size_t>(num_sms) * v;
		}
	}
	if (grid > p.num_spectra) {
		grid = p.num_spectra;
	}
	PolyScreenFitKernel<K, NT, CAL><<<static_cast<unsigned>(grid), NT, smem, stream>>>(p,
			GetPolyConst(p.n), num_sms);
	return cudaGetLastError();
}

template<int K>
cudaError_t LaunchScreenK(PolyFitParams const &p, int num_sms, cudaStream_t stream) {
	bool const cal = p.cal.reference != nullptr;
	if (p.n == 32 * 128) {
		return cal ? LaunchScreenKNC<K, 128, true>(p, num_sms, stream) :
				LaunchScreenKNC<K, 128, false>(p, num_sms, stream);
	}
	return cal ? LaunchScreenKNC<K, 256, true>(p, num_sms, stream) :
			LaunchScreenKNC<K, 256, false>(p, num_sms, stream);
}

} // namespace

bool PolyScreenSupported(PolyFitParams const &p) {
	if (p.n != 32 * 128 && p.n != 32 * 256) {
		return false;
	}
	if (p.K < 1 || p.K > kMaxPolyK || p.num_spectra > static_cast<size_t>(INT32_MAX)) {
		return false;
	}
	// 128-bit stores of the final mask, bulk prefetch of the mask rows
	return p.mask_stride % 16 == 0 && (reinterpret_cast<uintptr_t>(p.final_mask) % 16) == 0
			&& (reinterpret_cast<uintptr_t>(p.mask) % 16) == 0;
}

cudaError_t LaunchPolyScreenFit(PolyFitParams const &p, int num_sms, cudaStream_t stream) {
	switch (p.K) {
	case 1:
		return LaunchScreenK<1>(p, num_sms, stream);
	case 2:
		return LaunchScreenK<2>(p, num_sms, stream);
	case 3:
		return LaunchScreenK<3>(p, num_sms, stream);
	case 4:
		return LaunchScreenK<4>(p, num_sms, stream);
	case 5:
		return LaunchScreenK<5>(p, num_sms, stream);
	case 6:
		return LaunchScreenK<6>(p, num_sms, stream);
	case 7:
		return LaunchScreenK<7>(p, num_sms, stream);
	case 8:
		return LaunchScreenK<8>(p, num_sms, stream);
	default:
		return cudaErrorInvalidValue;
	}
}

} // namespace sakura_b200
