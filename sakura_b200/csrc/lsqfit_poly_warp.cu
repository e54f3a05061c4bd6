// Polynomial baseline fit, third generation: ONE WARP PER SPECTRUM, no block barrier anywhere
// (sakura_LSQFitPolynomialFloat semantics, libsakura/src/baseline.cc:1115-1224, 1268-1329).
//
// PolyScreenFitKernel (lsqfit_poly_screen.cu) spends 42 % of its warp time at the two block barriers
// around a serial control section and issues 19.5 k warp instructions per 4096-channel spectrum. This
// kernel keeps its arithmetic contract -- clip decisions identical to the per-channel evaluation,
// threshold INTERVALS from the moments, rows it cannot decide handed to PolyFitKernel -- and changes
// the structure:
//
//   * a warp owns a spectrum from load to store; twelve warps (= twelve spectra) are resident per SM
//     and never synchronise with each other, so the serial section of one spectrum (solve, bounds,
//     thresholds) overlaps the channel passes of the eleven others;
//   * the row arrives by ONE bulk asynchronous copy (cp.async.bulk, mbarrier completion) and the
//     residual leaves by one bulk store: no per-element load/store instructions;
//   * flagged and clipped channels are overwritten by NaN in shared memory: the screening pass then
//     needs no mask at all, because max(m, |NaN|) = m. Per four channels it issues one 128-bit load,
//     four FFMA2 (local quadratic of the model, two channels per instruction), two FADD2, two
//     FMNMX3 and a compare: 2.75 instructions per channel instead of 8;
//   * the few groups the screen cannot clear are listed and resolved COOPERATIVELY, eight groups
//     (32 channels, one per lane) per step, with the exact evaluation of the per-channel kernel
//     (FP64 Horner, float residual, the predicate of baseline.cc:1088 against both ends of each
//     threshold interval); clipped channels are downdated from the moments in the same step;
//   * data moments sum y u^k come from the FP64 tensor cores (DMMA m8n8k4: A = 8 blocks x 4
//     channels of data, B = integer powers of the block-local channel index, exact), Taylor-shifted
//     to the row origin per 32-channel block; warp totals of FP64 values are two DMMAs with a matrix
//     of ones instead of five shuffle rounds;
//   * power sums of the flagged channels: whole flagged 16-channel segments come from a table,
//     partly flagged ones from the smaller of {flagged, valid} channels, spread evenly over the lanes
//     through a list.
//
// Shared memory per warp: the row (4 n bytes), a 2 KiB scratch (lists / moment staging), the valid
// bits (n / 8 bytes) and a mbarrier: 19.2 KB at n = 4096.
#include "lsqfit_poly.cuh"
#include "device_common.cuh"
#include "poly_device.cuh"
#include "async_copy.cuh"

#include <float.h>
#include <stdlib.h>
#include <map>
#include <mutex>
#include <vector>

namespace sakura_b200 {

namespace {

constexpr double kU32 = 5.9604644775390625e-8; // 2^-24: unit roundoff of float
constexpr double kSlack = 1.0009765625; // 1 + 2^-10
constexpr double kTiny40 = 9.094947017729282e-13; // 2^-40
constexpr double kTiny44 = 5.684341886080802e-14; // 2^-44
constexpr int kScratchBytes = 2048;
constexpr int kListCap = 960; // u16 entries of the work lists; the last 128 scratch bytes hold the dirty bits
constexpr int kSegStride = 16; // doubles per row of the segment table (>= kMaxMoments)

enum WarpMode : int {
	kWarpClip = 0, kWarpFinal = 1, kWarpGiveUp = 2
};

__host__ __device__ constexpr int WarpsPerCta(int NB) {
	return NB >= 4 ? 12 : 16;
}

__host__ __device__ constexpr size_t SliceBytes(int NB) {
	// row | scratch | valid bits | mbarrier (+ pad to 16)
	return static_cast<size_t>(NB) * 4096 + kScratchBytes + static_cast<size_t>(NB) * 128 + 16;
}

__host__ __device__ constexpr size_t TableBytes() {
	return 32 * 8 * sizeof(double); // DMMA B fragments: 8 doubles per lane
}

__device__ __forceinline__ double warp_max_upper(double a) {
	unsigned const bits = __float_as_uint(__double2float_ru(a));
	return static_cast<double>(__uint_as_float(__reduce_max_sync(0xffffffffu, bits)));
}

__device__ __forceinline__ double rsqrt_newton(double v) {
	// FP64 hardware estimate (relative error ~2^-20) + two Newton steps: ~1e-13 relative
	double r;
	asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(v));
	double const hv = 0.5 * v;
	r = fma(r, fma(-hv * r, r, 0.5), r);
	r = fma(r, fma(-hv * r, r, 0.5), r);
	return r;
}

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

__host__ __device__ constexpr double Binomial(int n, int k) {
	double r = 1.0;
	for (int i = 1; i <= k; ++i) {
		r = r * static_cast<double>(n - k + i) / static_cast<double>(i);
	}
	return r;
}

__device__ __forceinline__ float keep_if_bit(float value, uint32_t word, int pos) {
	int m;
	asm("bfe.s32 %0, %1, %2, 1;" : "=r"(m) : "r"(word), "r"(pos));
	return __int_as_float(__float_as_int(value) & m);
}

__device__ __forceinline__ void bulk_store_s2g(void *gdst, void const *ssrc, unsigned bytes) {
	asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;"
			:: "l"(gdst), "r"(smem_u32(ssrc)), "r"(bytes) : "memory");
	asm volatile("cp.async.bulk.commit_group;" ::: "memory");
}

__device__ __forceinline__ void bulk_wait_read_all() {
	asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
}

/** exclusive prefix sum over the lanes */
__device__ __forceinline__ int warp_exclusive_scan(int v, int lane) {
	int incl = v;
#pragma unroll
	for (int o = 1; o < 32; o <<= 1) {
		int const t = __shfl_up_sync(0xffffffffu, incl, o);
		incl += (lane >= o) ? t : 0;
	}
	return incl - v;
}

template<int K>
struct WarpFit {
	double cz[K]; // model coefficients in the channel index z
	double cx[K]; // ... in x = z / (n - 1)
	float lo_m, lo_p, hi_m, hi_p; // interval ends of the lower / upper clip threshold
	float c_in; // |screening residual| below this: provably inside both thresholds
	int mode;
};

// Solve for the current moments and derive what the next pass needs. Every lane does the same work
// (the warp owns the row; there is nobody to wait for it), except the maximum of |model| over 32
// Chebyshev nodes, one per lane. Follows Control<> of lsqfit_poly_screen.cu; the screening band is
// the one of this kernel's single-float local quadratic (derivation in DESIGN.md 3.1c).
template<int K>
__device__ __forceinline__ void ControlWarp(double const (&S)[2 * K - 1], double const (&T)[K], double Q,
		PolyConst const &pc, double xq, bool first, int count, float clip, bool want_clip, double &kappa,
		double &ikappa, WarpFit<K> &f) {
	double cx[K];
	bool const solved = SolveLDLT<K>(S, T, cx);
	double c1 = 0.0;
#pragma unroll
	for (int k = 0; k < K; ++k) {
		c1 += fabs(cx[k]);
	}
	double mv = cx[K - 1];
#pragma unroll
	for (int k = K - 2; k >= 0; --k) {
		mv = fma(mv, xq, cx[k]);
	}
	mv = warp_max_upper(fabs(mv));
	double const A = fma(mv, pc.ez[K - 1], kTiny44 * c1);
	int mode = kWarpGiveUp;
	float lo_m = 0.0f, lo_p = 0.0f, hi_m = 0.0f, hi_p = 0.0f, c_in = 0.0f;
	// every comparison below is written so that a NaN anywhere leaves mode == kWarpGiveUp
	if (solved && c1 < DBL_MAX && A < DBL_MAX) {
		if (!want_clip) {
			mode = kWarpFinal;
		} else {
			double const N = static_cast<double>(count);
			double const inv_n = fast_rcp(N);
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
			double const e64_ss = kTiny40 * 2.0 * (Q + c1 * c1 * N);
			double const ss_up = ss + e64_ss;
			double const ssc = (ss_up > 0.0) ? ss_up : 0.0;
			double const a2n = A * A * N;
			double const a_sum_rho = first ? 0.5 * (a2n + ssc) : 0.5 * A * fma(N, kappa, ssc * ikappa);
			double const d_ss = 2.0 * kU32 * kSlack * (a_sum_rho + ssc) + 2.0 * kU32 * kU32 * (a2n + ssc) + e64_ss
					+ kTiny44 * ssc;
			double const q_m = (ss - d_ss) * inv_n;
			double const q_p = (ss + d_ss) * inv_n;
			if (q_m > 1e-30 && q_p < 1e30 && clip > 0.0f) {
				double const r_m = rsqrt_newton(q_m);
				double const r_p = rsqrt_newton(q_p);
				double const s_p = q_p * r_p * (1.0 + kTiny40);
				kappa = s_p;
				ikappa = r_p;
				double const d_mean = kU32 * kSlack * (A + s_p) + kTiny40 * (A + s_p + c1);
				double const mean_c = sum_r * inv_n;
				double const mean_m = mean_c - d_mean;
				double const mean_p = mean_c + d_mean;
				double const mabs = fabs(mean_c) + d_mean;
				double const m2max = mabs * mabs;
				double const sigma = static_cast<double>(clip);
				double const rms_m = fma(-m2max * (1.0 + kTiny40), r_m, q_m * r_m);
				double const rms_p = s_p;
				float const thr_m = static_cast<float>(rms_m * (sigma * (1.0 - 3.0 * kTiny40)));
				float const thr_p = static_cast<float>(rms_p * (sigma * (1.0 + 3.0 * kTiny40)));
				double const nudge = kTiny40 * (mabs + static_cast<double>(thr_p));
				lo_m = static_cast<float>(mean_m - static_cast<double>(thr_p) - nudge);
				lo_p = static_cast<float>(mean_p - static_cast<double>(thr_m) + nudge);
				hi_m = static_cast<float>(mean_m + static_cast<double>(thr_m) - nudge);
				hi_p = static_cast<float>(mean_p + static_cast<double>(thr_p) + nudge);
				bool const by_compare = lo_p <= -0x1p-80f && hi_m >= 0x1p-80f
						&& fabsf(lo_m) <= FLT_MAX && fabsf(hi_p) <= FLT_MAX
						&& __fsub_rn(hi_m, lo_p) >= 0x1p-40f;
				if (by_compare) {
					// lo_m <= lo_p < 0 < hi_m <= hi_p here
					double const theta = static_cast<double>(fmaxf(-lo_m, hi_p));
					// bounds over the row on the terms of the local expansion of the model about a block
					// origin x0: |d_j| <= h^j sum_{k >= j} C(k, j) |c_k| (x0 in [0, 1]), local index <= 31
					double const e1 = 31.0 * pc.h;
					double V = 0.0, V3 = 0.0;
					double ep = 1.0;
#pragma unroll
					for (int j = 1; j < K; ++j) {
						ep *= e1;
						double b = 0.0;
#pragma unroll
						for (int k = j; k < K; ++k) {
							b = fma(Binomial(k, j), fabs(cx[k]), b);
						}
						if (j < 3) {
							V = fma(ep, b, V);
						} else {
							V3 = fma(ep, b, V3);
						}
					}
					double const band = kSlack * (kU32 * (3.0 * A + 3.0 * V + 6.0 * theta) + V3)
							+ kTiny40 * (A + c1);
					float const c0 = fminf(hi_m, -lo_p);
					c_in = __double2float_rd(static_cast<double>(c0) - band);
					if (c_in > 0.0f) {
						mode = kWarpClip;
					}
				}
			}
		}
	}
#pragma unroll
	for (int k = 0; k < K; ++k) {
		f.cx[k] = cx[k];
		f.cz[k] = cx[k] * pc.hpow[k];
	}
	f.mode = mode;
	f.lo_m = lo_m;
	f.lo_p = lo_p;
	f.hi_m = hi_m;
	f.hi_p = hi_p;
	f.c_in = c_in;
}

// 16 mask bytes -> 16 bits (any non-zero byte is "true"), bit b <-> byte b
__device__ __forceinline__ uint32_t MaskBits16(uint4 const m) {
	auto nib = [](uint32_t w) {
		// bit 7 of every non-zero byte, then the four flags gathered into the top nibble by a product
		uint32_t const t = (((w & 0x7f7f7f7fu) + 0x7f7f7f7fu) | w) & 0x80808080u;
		return (t * 0x00204081u) >> 28;
	};
	return nib(m.x) | (nib(m.y) << 4) | (nib(m.z) << 8) | (nib(m.w) << 12);
}

// CAL: position-switch calibration (and NaN/Inf flagging) applied in shared memory as soon
// as the row has arrived: the fit then sees scaling * (data - reference) / reference, the calibrated
// spectrum is never written to global memory.
template<int K, int NB, bool CAL>
__global__ void __launch_bounds__(32 * WarpsPerCta(NB), 1)
PolyWarpFitKernel(PolyFitParams const p, PolyConst const pc, double const *__restrict__ seg_sums) {
	constexpr int NS = 2 * K - 1;
	constexpr int NV = 3 * K;
	constexpr int n = 1024 * NB;
	constexpr int W = WarpsPerCta(NB);
	constexpr int NSEG = 2 * NB; // 16-channel segments per lane (interleaved ownership of the mask)
	constexpr int NPW = NSEG / 2; // ... packed two per 32-bit word
	extern __shared__ __align__(128) unsigned char smem_raw[];
	int const lane = threadIdx.x & 31;
	int const warp = threadIdx.x >> 5;
	double *bpow_tab = reinterpret_cast<double *>(smem_raw); // [8][32]
	unsigned char *slice = smem_raw + TableBytes() + static_cast<size_t>(warp) * SliceBytes(NB);
	float *sdata = reinterpret_cast<float *>(slice);
	unsigned char *scratch = slice + static_cast<size_t>(n) * 4;
	uint16_t *slist = reinterpret_cast<uint16_t *>(scratch);
	double *sstage = reinterpret_cast<double *>(scratch);
	uint32_t *sdirty = reinterpret_cast<uint32_t *>(scratch + 2 * kListCap); // bit g: group g holds a NaN
	uint32_t *svalid = reinterpret_cast<uint32_t *>(scratch + kScratchBytes);
	uint64_t *mbar = reinterpret_cast<uint64_t *>(scratch + kScratchBytes + static_cast<size_t>(NB) * 128);

	// DMMA B fragments of the data-moment pass: B_j[k][m] = (8 k + j)^m, k = lane % 4, m = lane / 4
	if (warp == 0) {
#pragma unroll 1
		for (int j = 0; j < 8; ++j) {
			double const u = static_cast<double>(8 * (lane & 3) + j);
			double v = 1.0;
#pragma unroll 1
			for (int m = 0; m < (lane >> 2); ++m) {
				v *= u;
			}
			bpow_tab[j * 32 + lane] = v;
		}
	}
	if (lane == 0) {
		mbar_init(mbar, 1);
		mbar_fence_init();
	}
	__syncthreads(); // the only block barrier: constant table and barriers initialised
	unsigned phase = 0u;

	int const l7 = lane & 7;
	float const clip = p.clip_threshold_sigma;
	double const xq = pc.node[lane];
	// local channel index pairs of the lane's groups in the order it visits them: at step q the lane
	// reads group q ^ l7 of a block (conflict-free 128-bit shared-memory accesses at a lane stride of
	// 128 bytes)
	float2 u01[8], u23[8];
#pragma unroll
	for (int q = 0; q < 8; ++q) {
		float const ua = static_cast<float>(4 * (q ^ l7));
		u01[q] = make_float2(ua, ua + 1.0f);
		u23[q] = make_float2(ua + 2.0f, ua + 3.0f);
	}
	float4 *sd4 = reinterpret_cast<float4 *>(sdata);
	int const own4 = lane * (8 * NB); // first float4 group of the lane's consecutive channels
	// shared-memory addresses of the eight groups of the lane's first block in visiting order
	uint32_t qaddr[8];
#pragma unroll
	for (int q = 0; q < 8; ++q) {
		qaddr[q] = smem_u32(sd4 + own4 + (q ^ l7));
	}

	for (size_t row = static_cast<size_t>(blockIdx.x) * W + warp; row < p.num_spectra;
			row += static_cast<size_t>(gridDim.x) * W) {
		float const *grow = p.data + row * p.data_stride;
		// ---- start the row's copy; the previous row's bulk store must have read the buffer ----
		if (lane == 0) {
			bulk_wait_read_all();
			fence_proxy_async();
			mbar_expect_tx(mbar, static_cast<unsigned>(n) * 4u);
			bulk_copy_g2s(sdata, grow, static_cast<unsigned>(n) * 4u, mbar);
		}
		// ---- mask bytes -> bits, interleaved ownership: segment lane + 32 j = channels 16 (lane + 32 j).. ----
		uint32_t vseg[NPW]; // valid bits of segments (2 w, 2 w + 1) in the low / high half
		int count;
		{
			uint4 const *gmask = reinterpret_cast<uint4 const *>(p.mask + row * p.mask_stride);
			uint4 mb[NSEG];
#pragma unroll
			for (int j = 0; j < NSEG; ++j) {
				mb[j] = __ldg(gmask + lane + 32 * j);
			}
			uint32_t fin[NSEG]; // CAL: finite flags of the calibrated values of segment j
			if (CAL) {
				// the row has arrived: calibrate the lane's mask segments in place (16 channels each)
				mbar_wait(mbar, phase);
				phase ^= 1u;
				float4 const *ref4 = reinterpret_cast<float4 const *>(p.cal.reference + row * p.cal.reference_stride);
				float4 const *scl4 = (p.cal.scaling != nullptr) ?
						reinterpret_cast<float4 const *>(p.cal.scaling + row * p.cal.scaling_stride) : nullptr;
#pragma unroll
				for (int j = 0; j < NSEG; ++j) {
					int const g0 = 4 * (lane + 32 * j);
					float4 r4[4], s4[4];
#pragma unroll
					for (int e = 0; e < 4; ++e) {
						r4[e] = __ldg(ref4 + g0 + e);
						s4[e] = (scl4 != nullptr) ? __ldg(scl4 + g0 + e) : make_float4(p.cal.const_scaling,
								p.cal.const_scaling, p.cal.const_scaling, p.cal.const_scaling);
					}
					uint32_t f = 0u;
#pragma unroll
					for (int e = 0; e < 4; ++e) {
						float4 y = sd4[g0 + e];
						y.x = CalibrateOne(s4[e].x, y.x, r4[e].x);
						y.y = CalibrateOne(s4[e].y, y.y, r4[e].y);
						y.z = CalibrateOne(s4[e].z, y.z, r4[e].z);
						y.w = CalibrateOne(s4[e].w, y.w, r4[e].w);
						sd4[g0 + e] = y;
						auto finite = [](float v) {
							return (__float_as_uint(v) & 0x7f800000u) != 0x7f800000u;
						};
						f |= ((finite(y.x) ? 1u : 0u) | (finite(y.y) ? 2u : 0u) | (finite(y.z) ? 4u : 0u)
								| (finite(y.w) ? 8u : 0u)) << (4 * e);
					}
					fin[j] = p.cal.flag_nonfinite ? f : 0xffffu; // sakura_SetFalseIfNanOrInfFloat merged into the mask
				}
				__syncwarp();
			}
			int c = 0;
#pragma unroll
			for (int w = 0; w < NPW; ++w) {
				uint32_t a = MaskBits16(mb[2 * w]);
				uint32_t b = MaskBits16(mb[2 * w + 1]);
				if (CAL) {
					a &= fin[2 * w];
					b &= fin[2 * w + 1];
				}
				reinterpret_cast<uint16_t *>(svalid)[lane + 32 * (2 * w)] = static_cast<uint16_t>(a);
				reinterpret_cast<uint16_t *>(svalid)[lane + 32 * (2 * w + 1)] = static_cast<uint16_t>(b);
				vseg[w] = a | (b << 16);
				c += __popc(vseg[w]);
			}
			count = __reduce_add_sync(0xffffffffu, c);
		}
		bool skip = false; // the row ends here (its copy is still awaited below)
		bool giveup = false;
		if (p.num_fitting_max == 0) { // baseline.cc:1295-1296: kOK, nothing written
			if (lane == 0) {
				p.status[row] = kStatusOK;
				p.lsqfit_status[row] = kLsqOK;
			}
			skip = true;
		} else if (count < K) { // baseline.cc:1147-1150
			if (lane == 0) {
				p.status[row] = kStatusNG;
				p.lsqfit_status[row] = kLsqNotEnoughData;
			}
			skip = true;
		}
		// ---- power sums over the valid channels (needs no data: overlaps the copy) ----------------
		// target set X = the flagged channels (S = s_all - R) or, when more than half of the row is
		// flagged, the valid ones (S = R); per 16-channel segment the smaller of X and its complement
		// is visited (the other side comes from the table of whole-segment sums)
		double ps[NS];
#pragma unroll
		for (int k = 0; k < NS; ++k) {
			ps[k] = 0.0;
		}
		bool const direct = (2 * count < n);
		if (!skip) {
			uint16_t const *sv16 = reinterpret_cast<uint16_t const *>(svalid);
			auto visited = [&](int s, bool &comp) {
				uint32_t const vb = sv16[lane + 32 * s]; // (this lane's own store above)
				uint32_t const x = direct ? vb : (~vb & 0xffffu);
				comp = __popc(x) > 8;
				return comp ? (~x & 0xffffu) : x;
			};
			uint32_t neg = 0u; // bit s: segment s of this lane is visited through its complement
			int mine = 0;
#pragma unroll 1
			for (int s = 0; s < NSEG; ++s) {
				bool comp;
				mine += __popc(visited(s, comp));
				if (__any_sync(0xffffffffu, comp)) {
					if (comp) {
						double const *rowp = seg_sums + static_cast<size_t>(lane + 32 * s) * kSegStride;
#pragma unroll
						for (int k = 0; k < NS; k += 2) {
							double2 const v = __ldg(reinterpret_cast<double2 const *>(rowp + k));
							ps[k] += v.x;
							if (k + 1 < NS) {
								ps[k + 1] += v.y;
							}
						}
						neg |= 1u << s;
					}
				}
			}
			int const total = __reduce_add_sync(0xffffffffu, mine);
			if (total > kListCap) {
				giveup = true; // (cannot happen below 8 visited channels per segment on average)
			} else if (total > 0) {
				int pos = warp_exclusive_scan(mine, lane);
#pragma unroll 1
				for (int w = 0; w < NPW; ++w) {
					bool c0, c1;
					uint32_t b = visited(2 * w, c0);
					b |= visited(2 * w + 1, c1) << 16;
					while (b != 0u) {
						int const s = __ffs(b) - 1;
						b &= b - 1u;
						int const seg = 2 * w + (s >> 4);
						int const ch = 16 * (lane + 32 * seg) + (s & 15);
						slist[pos] = static_cast<uint16_t>(ch | (((neg >> seg) & 1u) << 15));
						++pos;
					}
				}
				__syncwarp();
				for (int it = lane; it < total; it += 32) {
					uint32_t const e = slist[it];
					double const x = u32_to_double(e & 0x7fffu) * pc.h;
					double const sg = (e & 0x8000u) ? -1.0 : 1.0;
					double pw[NS];
					PowerTree<NS>(x, pw);
#pragma unroll
					for (int k = 0; k < NS; ++k) {
						ps[k] = fma(sg, pw[k], ps[k]);
					}
				}
				__syncwarp();
			}
		}
		// ---- the row has arrived ------------------------------------------------------------------
		if (!CAL) {
			mbar_wait(mbar, phase);
			phase ^= 1u;
		}
		if (giveup) {
			if (lane == 0) {
				int const pos = atomicAdd(p.fallback_count, 1);
				p.fallback_rows[pos] = static_cast<int32_t>(row);
			}
			skip = true;
		}
		if (skip) {
			continue;
		}
		// ---- data moments: per 32-channel block sum y u^m (u = 0..31, m < 8) on the FP64 tensor cores,
		// shifted to the row origin by the lane that takes the block ----------------------------------
		double tm[K];
		double qsum = 0.0;
#pragma unroll
		for (int k = 0; k < K; ++k) {
			tm[k] = 0.0;
		}
		{
			double bp[8];
#pragma unroll
			for (int j = 0; j < 8; ++j) {
				bp[j] = bpow_tab[j * 32 + lane];
			}
			int const mrow = lane >> 2;
			int const kcol = lane & 3;
#pragma unroll 1
			for (int r = 0; r < NB; ++r) {
				// four independent sequences of eight dependent products each, interleaved
				float ye[4][8];
				uint32_t vb[4];
#pragma unroll
				for (int q = 0; q < 4; ++q) {
					int const blk = 32 * r + 8 * q + mrow;
					float4 const ya = sd4[8 * blk + 2 * kcol];
					float4 const yb = sd4[8 * blk + 2 * kcol + 1];
					ye[q][0] = ya.x, ye[q][1] = ya.y, ye[q][2] = ya.z, ye[q][3] = ya.w;
					ye[q][4] = yb.x, ye[q][5] = yb.y, ye[q][6] = yb.z, ye[q][7] = yb.w;
					vb[q] = reinterpret_cast<uint8_t const *>(svalid)[4 * blk + kcol];
				}
				double c0[4] = { 0.0, 0.0, 0.0, 0.0 }, c1[4] = { 0.0, 0.0, 0.0, 0.0 };
#pragma unroll
				for (int j = 0; j < 8; ++j) {
#pragma unroll
					for (int q = 0; q < 4; ++q) {
						// (a select on the bits: a NaN / Inf under the mask must not leak)
						double const yd = static_cast<double>(keep_if_bit(ye[q][j], vb[q], j));
						qsum = fma(yd, yd, qsum);
						dmma8x8x4(c0[q], c1[q], yd, bp[j]);
					}
				}
#pragma unroll
				for (int q = 0; q < 4; ++q) {
					reinterpret_cast<double2 *>(sstage)[((8 * q + mrow) * 8 + 2 * kcol) >> 1] = make_double2(c0[q], c1[q]);
				}
				__syncwarp();
				{
					double t[8];
					double2 const *src = reinterpret_cast<double2 const *>(sstage + lane * 8);
#pragma unroll
					for (int k = 0; k < (K + 1) / 2; ++k) {
						double2 const v = src[k];
						t[2 * k] = v.x;
						t[2 * k + 1] = v.y;
					}
					double const zb = static_cast<double>(32 * (32 * r + lane));
#pragma unroll
					for (int i = 0; i < K - 1; ++i) {
#pragma unroll
						for (int j = K - 1; j > i; --j) {
							t[j] = fma(zb, t[j - 1], t[j]);
						}
					}
#pragma unroll
					for (int k = 0; k < K; ++k) {
						tm[k] += t[k];
					}
				}
				__syncwarp();
			}
		}
		sdirty[lane] = 0u; // (the staging area of the moments is dead: barrier at the end of the last round)
		// ---- totals: every lane gets S (2K-1), T (K), Q ---------------------------------------------
		double S[NS], T[K], Q;
#pragma unroll
		for (int k = 0; k < NS; ++k) {
			double const tot = warp_total(ps[k]);
			S[k] = direct ? tot : (pc.s_all[k] - tot);
		}
#pragma unroll
		for (int k = 0; k < K; ++k) {
			T[k] = warp_total(tm[k]) * pc.hpow[k];
		}
		Q = warp_total(qsum);

		// ---- clip iterations --------------------------------------------------------------------
		int num_unmasked = n; // baseline.cc:1152 (sic)
		bool failed = false;
		double kappa = 0.0, ikappa = 0.0;
		WarpFit<K> fit;
		for (int iter = 1; iter <= p.num_fitting_max; ++iter) {
			if (num_unmasked < K) { // baseline.cc:1162-1165
				failed = true;
				break;
			}
			ControlWarp<K>(S, T, Q, pc, xq, iter == 1, count, clip, iter < p.num_fitting_max, kappa, ikappa, fit);
			if (fit.mode == kWarpGiveUp) {
				giveup = true;
				break;
			}
			if (fit.mode == kWarpFinal) {
				break;
			}
			// ---- screening pass: |y - local quadratic| >= c_in somewhere in a group of four -> suspect ----
			uint32_t susp = 0u;
			float const c_in = fit.c_in;
#pragma unroll 1
			for (int j = 0; j < NB; ++j) {
				double d[K];
				TaylorShift<K>(fit.cz, static_cast<double>(32 * (NB * lane + j)), d);
				// negated model: r = y + (-model)
				float const n0 = -static_cast<float>(d[0]);
				float const n1 = (K > 1) ? -static_cast<float>(d[K > 1 ? 1 : 0]) : 0.0f;
				float const n2 = (K > 2) ? -static_cast<float>(d[K > 2 ? 2 : 0]) : 0.0f;
				float2 const d0 = make_float2(n0, n0);
				float2 const d1 = make_float2(n1, n1);
				float2 const d2 = make_float2(n2, n2);
				uint32_t sj = 0u; // bit q: the group visited at step q
#pragma unroll
				for (int q = 0; q < 8; ++q) {
					float4 y;
					// (volatile: the row changes between passes -- clipped channels become NaN)
					asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(y.x), "=f"(y.y), "=f"(y.z), "=f"(y.w)
							: "r"(qaddr[q] + 128u * static_cast<unsigned>(j)));
					float2 const t01 = __ffma2_rn(d2, u01[q], d1);
					float2 const t23 = __ffma2_rn(d2, u23[q], d1);
					float2 const p01 = __ffma2_rn(t01, u01[q], d0);
					float2 const p23 = __ffma2_rn(t23, u23[q], d0);
					float2 const r01 = __fadd2_rn(make_float2(y.x, y.y), p01);
					float2 const r23 = __fadd2_rn(make_float2(y.z, y.w), p23);
					// a NaN (flagged or clipped channel) is ignored by the maximum
					float const m = fmaxf(fmaxf(fmaxf(fabsf(r01.x), fabsf(r01.y)), fabsf(r23.x)), fabsf(r23.y));
					if (m >= c_in) {
						sj |= 1u << q;
					}
				}
				susp |= sj << (8 * j);
			}
			// ---- the groups the screen could not clear: listed, then resolved eight per step, one
			// channel per lane, exactly as the per-channel kernel evaluates it ----------------------
			int const mine = __popc(susp);
			int const total = __reduce_add_sync(0xffffffffu, mine);
			int num_clipped = 0;
			if (total > kListCap) {
				giveup = true; // (more than 94 % of the groups suspect: the per-channel kernel takes the row)
				break;
			}
			if (total > 0) {
				{
					int pos = warp_exclusive_scan(mine, lane);
					uint32_t b = susp;
					while (b != 0u) {
						int const s = __ffs(b) - 1;
						b &= b - 1u;
						slist[pos] = static_cast<uint16_t>(8 * NB * lane + (s ^ l7)); // (step -> group of the block)
						++pos;
					}
				}
				__syncwarp();
				double cs[NS], ct[K], cq = 0.0;
#pragma unroll
				for (int k = 0; k < NS; ++k) {
					cs[k] = 0.0;
				}
#pragma unroll
				for (int k = 0; k < K; ++k) {
					ct[k] = 0.0;
				}
				bool undecided = false;
				for (int base = 0; base < total; base += 8) {
					int const idx = base + (lane >> 2);
					bool clipped = false;
					float y = 0.0f;
					int ch = 0;
					bool mute = false; // a flagged channel far from the model: silenced for the later passes
					if (idx < total) {
						ch = 4 * static_cast<int>(slist[idx]) + (lane & 3);
						y = sdata[ch];
						bool const valid = ((svalid[ch >> 5] >> (ch & 31)) & 1u) != 0u;
						double const z = static_cast<double>(ch);
						double acc = fit.cz[K - 1];
#pragma unroll
						for (int k = K - 2; k >= 0; --k) {
							acc = fma(acc, z, fit.cz[k]);
						}
						float const r = __fsub_rn(y, static_cast<float>(acc));
						if (valid && ch != n - 1) { // (channel n-1 is never examined: baseline.cc:1078)
							if (r > fit.hi_p || r < fit.lo_m) {
								clipped = true; // outside whatever the exact thresholds are
							} else if (!(r <= fit.hi_m && r >= fit.lo_p)) {
								undecided = true; // inside a threshold interval
							}
						} else if (!(fabsf(r) < c_in) && y == y) {
							mute = true;
						}
					}
					uint32_t const cb = __ballot_sync(0xffffffffu, clipped);
					if (clipped || mute) {
						sdata[ch] = __int_as_float(0x7fc00000);
						atomicOr(sdirty + (ch >> 7), 1u << ((ch >> 2) & 31));
					}
					if (cb != 0u) {
						num_clipped += __popc(cb);
						if (clipped) {
							double const x = static_cast<double>(ch) * pc.h;
							double const yd = static_cast<double>(y);
							double pw[NS];
							PowerTree<NS>(x, pw);
#pragma unroll
							for (int k = 0; k < NS; ++k) {
								cs[k] += pw[k];
								if (k < K) {
									ct[k] = fma(yd, pw[k], ct[k]);
								}
							}
							cq = fma(yd, yd, cq);
							atomicAnd(svalid + (ch >> 5), ~(1u << (ch & 31)));
						}
					}
				}
				__syncwarp();
				if (__any_sync(0xffffffffu, undecided)) {
					giveup = true;
					break;
				}
				if (num_clipped > 0) {
					// downdate (numeric_operation.cc:283-314, 448-479)
#pragma unroll
					for (int k = 0; k < NS; ++k) {
						S[k] -= warp_total(cs[k]);
					}
#pragma unroll
					for (int k = 0; k < K; ++k) {
						T[k] -= warp_total(ct[k]);
					}
					Q -= warp_total(cq);
				}
			}
			if (num_clipped == 0) {
				break; // converged (baseline.cc:1211-1213); fit is the last fit
			}
			num_unmasked = count - num_clipped;
			count = num_unmasked;
		}
		if (giveup) {
			// PolyFitKernel fits this row afterwards; nothing of it has been written
			if (lane == 0) {
				int const pos = atomicAdd(p.fallback_count, 1);
				p.fallback_rows[pos] = static_cast<int32_t>(row);
			}
			continue;
		}
		// ---- groups that hold a NaN (clipped or silenced channels) get their data from global memory again:
		// asynchronous 16-byte copies, in flight while the final mask is expanded and written; the next
		// row's lines are requested into L2 now (not earlier: they would be evicted again before use) ----
		__syncwarp();
		if (p.residual != nullptr && !failed) {
			uint32_t dirty = (NB == 4) ? sdirty[lane] : ((sdirty[lane >> 1] >> (16 * (lane & 1))) & 0xffffu);
			float4 const *grow4 = reinterpret_cast<float4 const *>(grow);
			if (CAL) {
				// calibrated again from the three input rows (the same operations: the same bits), two groups
				// per step so that their six loads are in flight together
				float4 const *ref4 = reinterpret_cast<float4 const *>(p.cal.reference + row * p.cal.reference_stride);
				float4 const *scl4 = (p.cal.scaling != nullptr) ?
						reinterpret_cast<float4 const *>(p.cal.scaling + row * p.cal.scaling_stride) : nullptr;
				float4 const cs4 = make_float4(p.cal.const_scaling, p.cal.const_scaling, p.cal.const_scaling,
						p.cal.const_scaling);
				while (dirty != 0u) {
					int const ga = __ffs(dirty) - 1;
					dirty &= dirty - 1u;
					int const gb = (dirty != 0u) ? __ffs(dirty) - 1 : ga;
					dirty &= dirty - 1u; // (0 & anything = 0)
					float4 ya = __ldg(grow4 + own4 + ga), yb = __ldg(grow4 + own4 + gb);
					float4 const ra = __ldg(ref4 + own4 + ga), rb = __ldg(ref4 + own4 + gb);
					float4 const sa = (scl4 != nullptr) ? __ldg(scl4 + own4 + ga) : cs4;
					float4 const sb = (scl4 != nullptr) ? __ldg(scl4 + own4 + gb) : cs4;
					ya.x = CalibrateOne(sa.x, ya.x, ra.x);
					ya.y = CalibrateOne(sa.y, ya.y, ra.y);
					ya.z = CalibrateOne(sa.z, ya.z, ra.z);
					ya.w = CalibrateOne(sa.w, ya.w, ra.w);
					yb.x = CalibrateOne(sb.x, yb.x, rb.x);
					yb.y = CalibrateOne(sb.y, yb.y, rb.y);
					yb.z = CalibrateOne(sb.z, yb.z, rb.z);
					yb.w = CalibrateOne(sb.w, yb.w, rb.w);
					sd4[own4 + ga] = ya;
					sd4[own4 + gb] = yb;
				}
			} else {
				while (dirty != 0u) {
					int const g = __ffs(dirty) - 1;
					dirty &= dirty - 1u;
					asm volatile("cp.async.ca.shared.global [%0], [%1], 16;"
							:: "r"(smem_u32(sd4 + own4 + g)), "l"(grow4 + own4 + g) : "memory");
				}
			}
			asm volatile("cp.async.commit_group;" ::: "memory");
		}
		if (lane == 0) {
			size_t const nrow = row + static_cast<size_t>(gridDim.x) * W;
			if (nrow < p.num_spectra) {
				asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;"
						:: "l"(p.data + nrow * p.data_stride), "r"(static_cast<unsigned>(n) * 4u) : "memory");
				asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;"
						:: "l"(p.mask + nrow * p.mask_stride), "r"(static_cast<unsigned>(n)) : "memory");
				if (CAL) {
					if (p.cal.reference_stride != 0) {
						asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;"
								:: "l"(p.cal.reference + nrow * p.cal.reference_stride),
								"r"(static_cast<unsigned>(n) * 4u) : "memory");
					}
					if (p.cal.scaling != nullptr && p.cal.scaling_stride != 0) {
						asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;"
								:: "l"(p.cal.scaling + nrow * p.cal.scaling_stride),
								"r"(static_cast<unsigned>(n) * 4u) : "memory");
					}
				}
			}
		}
		// ---- final mask: 32 NB consecutive bytes per lane ---------------------------------------------
		uint32_t vw[NB];
		{
			__syncwarp();
#pragma unroll
			for (int j = 0; j < NB; ++j) {
				vw[j] = svalid[NB * lane + j];
			}
			uint4 *gfm = reinterpret_cast<uint4 *>(p.final_mask + row * p.mask_stride) + 2 * NB * lane;
#pragma unroll
			for (int j = 0; j < NB; ++j) {
				uint32_t w[8];
#pragma unroll
				for (int q = 0; q < 8; ++q) {
					w[q] = (((vw[j] >> (4 * q)) & 0xfu) * 0x00204081u) & 0x01010101u;
				}
				gfm[2 * j] = make_uint4(w[0], w[1], w[2], w[3]);
				gfm[2 * j + 1] = make_uint4(w[4], w[5], w[6], w[7]);
			}
		}
		if (failed) {
			// the reference has already clipped final_mask in place (SURVEY A-13 ii)
			if (lane == 0) {
				p.status[row] = kStatusNG;
				p.lsqfit_status[row] = kLsqNotEnoughData;
				if (p.flags != nullptr) {
					p.flags[row] = 2;
				}
			}
			continue;
		}
		// ---- the last fit, exactly: FP64 model -> float, float residual, FP64 statistics over the valid
		// channels. Flagged / clipped channels (NaN in shared memory) take their data from global memory
		// again (their lines are still in L2) ----------------------------------------------------------
		double sum = 0.0, sq = 0.0;
		{
			// four independent accumulator pairs (one per channel of a group): the additions of a single
			// pair would be one dependent chain of 8-cycle operations through the whole row
			double s4[4] = { 0.0, 0.0, 0.0, 0.0 }, q4[4] = { 0.0, 0.0, 0.0, 0.0 };
			float4 *gbest4 = (p.best_fit != nullptr) ?
					reinterpret_cast<float4 *>(p.best_fit + row * p.data_stride) : nullptr;
			asm volatile("cp.async.wait_all;" ::: "memory"); // (own copies into own groups: no warp barrier needed)
#pragma unroll 1
			for (int j = 0; j < NB; ++j) {
				double d[K];
				TaylorShift<K>(fit.cz, static_cast<double>(32 * (NB * lane + j)), d);
#pragma unroll 2
				for (int q = 0; q < 8; ++q) {
					int const g = q ^ l7;
					int const idx = own4 + 8 * j + g;
					uint32_t const nib = (vw[j] >> (4 * g)) & 0xfu;
					float4 const y = sd4[idx];
					float const ye[4] = { y.x, y.y, y.z, y.w };
					float re[4], me[4];
					double const ub = static_cast<double>(4 * g);
#pragma unroll
					for (int e = 0; e < 4; ++e) {
						double const u = (e == 0) ? ub : ub + static_cast<double>(e);
						double acc = d[K - 1];
#pragma unroll
						for (int k = K - 2; k >= 0; --k) {
							acc = fma(acc, u, d[k]);
						}
						me[e] = static_cast<float>(acc);
						float const r = __fsub_rn(ye[e], me[e]);
						re[e] = r;
						// statistics over the valid channels: a select on the float (whatever sits under the
						// mask -- NaN, Inf -- cannot leak), then unconditional FP64 accumulation
						double const rd = static_cast<double>(((nib >> e) & 1u) != 0u ? r : 0.0f);
						s4[e] += rd;
						q4[e] = fma(rd, rd, q4[e]);
					}
					sd4[idx] = make_float4(re[0], re[1], re[2], re[3]);
					if (gbest4 != nullptr) {
						gbest4[idx] = make_float4(me[0], me[1], me[2], me[3]);
					}
				}
			}
			sum = (s4[0] + s4[1]) + (s4[2] + s4[3]);
			sq = (q4[0] + q4[1]) + (q4[2] + q4[3]);
		}
		if (p.residual != nullptr) {
			fence_proxy_async();
			__syncwarp();
			if (lane == 0) {
				bulk_store_s2g(p.residual + row * p.data_stride, sdata, static_cast<unsigned>(n) * 4u);
			}
		}
		sum = warp_total(sum);
		sq = warp_total(sq);
		if (p.coeff != nullptr && lane < K) {
			// (static register indexing: select fit.cx[lane] with a chain of predicated moves)
			double cxl = 0.0;
#pragma unroll
			for (int k = 0; k < K; ++k) {
				cxl = (lane == k) ? fit.cx[k] : cxl;
			}
			if (p.chebyshev) {
				double a = 0.0;
#pragma unroll
				for (int k = K - 1; k >= 0; --k) {
					if (k >= lane) {
						a = fma(pc.cheb[lane][k], fit.cx[k], a);
					}
				}
				p.coeff[row * K + lane] = a;
			} else {
				p.coeff[row * K + lane] = cxl / pc.factor[lane];
			}
		}
		if (lane == 0) {
			ClipThresholds const th = make_thresholds(count, sum, sq, clip);
			p.rms[row] = static_cast<float>(th.rms);
			p.status[row] = kStatusOK;
			p.lsqfit_status[row] = kLsqOK;
			if (p.flags != nullptr) {
				p.flags[row] = 3;
			}
		}
		__syncwarp();
	}
	if (lane == 0) {
		bulk_wait_read_all(); // the last bulk store must have read shared memory before the CTA exits
		asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
	}
}

// ---- whole-segment power sums: seg[s][k] = sum_{i = 16 s}^{16 s + 15} x_i^k, x_i = fl(i h), per device ----
struct SegTable {
	double *dev = nullptr;
};

cudaError_t GetSegmentTable(int n, double const **out) {
	static std::mutex mu;
	static std::map<std::pair<int, int>, SegTable> cache; // (device, n)
	int device = 0;
	cudaError_t err = cudaGetDevice(&device);
	if (err != cudaSuccess) {
		return err;
	}
	std::lock_guard<std::mutex> lock(mu);
	auto key = std::make_pair(device, n);
	auto it = cache.find(key);
	if (it != cache.end()) {
		*out = it->second.dev;
		return cudaSuccess;
	}
	int const nseg = n / 16;
	std::vector<double> host(static_cast<size_t>(nseg) * kSegStride, 0.0);
	double const h = 1.0 / static_cast<double>(n - 1);
	for (int s = 0; s < nseg; ++s) {
		long double acc[kMaxMoments];
		for (int k = 0; k < kMaxMoments; ++k) {
			acc[k] = 0.0L;
		}
		for (int i = 16 * s; i < 16 * s + 16; ++i) {
			double const x = static_cast<double>(i) * h;
			long double pw = 1.0L;
			for (int k = 0; k < kMaxMoments; ++k) {
				acc[k] += pw;
				pw *= static_cast<long double>(x);
			}
		}
		for (int k = 0; k < kMaxMoments; ++k) {
			host[static_cast<size_t>(s) * kSegStride + k] = static_cast<double>(acc[k]);
		}
	}
	SegTable t;
	err = cudaMalloc(&t.dev, host.size() * sizeof(double));
	if (err != cudaSuccess) {
		return err;
	}
	// (synchronous, pageable source: done once per device and channel count)
	err = cudaMemcpy(t.dev, host.data(), host.size() * sizeof(double), cudaMemcpyHostToDevice);
	if (err != cudaSuccess) {
		cudaFree(t.dev);
		return err;
	}
	cache.emplace(key, t);
	*out = t.dev;
	return cudaSuccess;
}

template<int K, int NB, bool CAL>
cudaError_t LaunchWarpKN(PolyFitParams const &p, int num_sms, cudaStream_t stream) {
	constexpr int W = WarpsPerCta(NB);
	constexpr size_t smem = TableBytes() + W * SliceBytes(NB);
	double const *seg = nullptr;
	cudaError_t err = GetSegmentTable(p.n, &seg);
	if (err != cudaSuccess) {
		return err;
	}
	err = cudaFuncSetAttribute(PolyWarpFitKernel<K, NB, CAL>, cudaFuncAttributeMaxDynamicSharedMemorySize,
			static_cast<int>(smem));
	if (err != cudaSuccess) {
		return err;
	}
	size_t grid = static_cast<size_t>(num_sms);
	size_t const need = (p.num_spectra + W - 1) / W;
	if (grid > need) {
		grid = need;
	}
	PolyWarpFitKernel<K, NB, CAL><<<static_cast<unsigned>(grid), 32 * W, smem, stream>>>(p, GetPolyConst(p.n), seg);
	return cudaGetLastError();
}

template<int K>
cudaError_t LaunchWarpK(PolyFitParams const &p, int num_sms, cudaStream_t stream) {
	if (p.cal.reference != nullptr) {
		return p.n == 4096 ? LaunchWarpKN<K, 4, true>(p, num_sms, stream) : LaunchWarpKN<K, 2, true>(p, num_sms, stream);
	}
	if (p.n == 4096) {
		return LaunchWarpKN<K, 4, false>(p, num_sms, stream);
	}
	return LaunchWarpKN<K, 2, false>(p, num_sms, stream);
}

} // namespace

bool PolyWarpSupported(PolyFitParams const &p) {
	if (p.n != 4096 && p.n != 2048) {
		return false;
	}
	if (p.K < 1 || p.K > kMaxPolyK || p.num_spectra > static_cast<size_t>(INT32_MAX)) {
		return false;
	}
	if (p.cal.reference != nullptr) { // calibrated in shared memory with 128-bit loads of the two other rows
		auto a16c = [](void const *q) {return (reinterpret_cast<uintptr_t>(q) % 16) == 0;};
		if (!a16c(p.cal.reference) || p.cal.reference_stride % 4 != 0 || !a16c(p.cal.scaling)
				|| p.cal.scaling_stride % 4 != 0) {
			return false;
		}
	}
	// bulk copies of the data rows, 128-bit accesses to the mask rows
	auto a16 = [](void const *q) {return (reinterpret_cast<uintptr_t>(q) % 16) == 0;};
	return p.mask_stride % 16 == 0 && p.data_stride % 4 == 0 && a16(p.final_mask) && a16(p.mask)
			&& a16(p.data) && a16(p.residual) && a16(p.best_fit);
}

cudaError_t LaunchPolyWarpFit(PolyFitParams const &p, int num_sms, cudaStream_t stream) {
	switch (p.K) {
	case 1:
		return LaunchWarpK<1>(p, num_sms, stream);
	case 2:
		return LaunchWarpK<2>(p, num_sms, stream);
	case 3:
		return LaunchWarpK<3>(p, num_sms, stream);
	case 4:
		return LaunchWarpK<4>(p, num_sms, stream);
	case 5:
		return LaunchWarpK<5>(p, num_sms, stream);
	case 6:
		return LaunchWarpK<6>(p, num_sms, stream);
	case 7:
		return LaunchWarpK<7>(p, num_sms, stream);
	case 8:
		return LaunchWarpK<8>(p, num_sms, stream);
	default:
		return cudaErrorInvalidValue;
	}
}

} // namespace sakura_b200
