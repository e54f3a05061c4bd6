// Batched sinusoid baseline fit (sakura_LSQFitSinusoidFloat semantics, baseline.cc:1835-1893 ->
// LSQFit :1268 -> DoLSQFit :1115) for the high-basis-count configs (BASELINE C4: 8192 channels,
// wave numbers 0..20 = 41 bases).
//
// One 256-thread CTA fits a tile of 32 spectra together, so that the FP64 tables are read once
// per tile instead of once per spectrum, and every dense contraction runs on the FP64 tensor
// cores (mma.sync m8n8k4 = DMMA):
//   trig sums      t[r][e]   = sum_i  mask[r][i]     * ext[i][e]        (rows x channels x 2*mmax+1)
//   data moments   b[r][k]   = sum_i  (mask*y)[r][i] * table[i][k]      (rows x channels x bases)
//   model          m[r][i]   = sum_k  coeff[r][k]    * table[i][k]      (rows x bases x channels)
//   LDL^T trailing updates of the per-row K x K solve (8 x 8 blocks)
// The Gram matrix needs no per-channel outer products at all: with phi = {1, sin(w t), cos(w t)}
//   sin a sin b = (cos(a-b) - cos(a+b))/2,  cos a cos b = (cos(a-b) + cos(a+b))/2,
//   sin a cos b = (sin(a+b) + sin(a-b))/2
// every entry is half the sum of two signed masked sums C_m = sum_valid cos(m t_i), S_m = sum_valid
// sin(m t_i), m <= 2*max wave (the trig sums above; which two, and their signs, is a per-call
// table built once per CTA). A clip pass only subtracts the clipped channels from the sums
// (downdate, like the reference's numeric_operation.cc:283-314). The solve is a blocked LDL^T of
// the matrix bordered by the right-hand side, one warp per row; a row whose pivots are not safely
// positive is handed to the general kernel (full-pivot rank-revealing LU).
//
// Thresholds, the float clip predicate, "last channel never clipped", early exit, not-enough-data
// and which outputs a failing row leaves untouched follow baseline.cc:1142-1224 exactly as in the
// other kernels; results are staged per tile and published only for rows that succeed.
#include "lsqfit_sinusoid.cuh"
#include "device_common.cuh"
#include "ldlt_device.cuh"
#include "async_copy.cuh"

#include <float.h>
#include <stdio.h>
#include <stdlib.h>

namespace sakura_b200 {

namespace {

constexpr int RB = kSinRowsPerTile;
constexpr int NWARP = kSinThreads / 32;
constexpr int NGROUP = 4; // warp groups (pairs) that split the channels
constexpr int MT = 4 * NGROUP / NWARP; // 8-row groups of the tile handled by one warp
static_assert(MT == 1 || MT == 2 || MT == 4, "warps of a group split the four 8-row groups");
constexpr unsigned kFull = 0xffffffffu;

enum RowState : int {
	kActive = 0, kConverged = 1, kFailedAfterClip = 2, kFailedBefore = 3, kFallback = 4, kNoRow = 5
};

// bits of a Gram table entry: which two trig sums, their signs, padding, and the (i, j) it fills
constexpr unsigned kGramNeg1 = 1u << 16;
constexpr unsigned kGramNeg2 = 1u << 17;
constexpr unsigned kGramPadDiag = 1u << 18;
constexpr unsigned kGramPad = 1u << 19;

__device__ __forceinline__ double select_f64(bool keep, double v) {
	// v or +0.0 without a branch and without touching v's NaN payloads when dropped
	return __hiloint2double(keep ? __double2hiint(v) : 0, keep ? __double2loint(v) : 0);
}

// dst[0..count) = src[0..count) by one warp, eight loads in flight per lane
template<typename V>
__device__ __forceinline__ void WarpCopy(V *__restrict__ dst, V const *__restrict__ src, int count, int lane) {
	int i = lane;
	for (; i + 7 * 32 < count; i += 8 * 32) {
		V v[8];
#pragma unroll
		for (int u = 0; u < 8; ++u) {
			v[u] = src[i + 32 * u];
		}
#pragma unroll
		for (int u = 0; u < 8; ++u) {
			dst[i + 32 * u] = v[u];
		}
	}
	for (; i < count; i += 32) {
		dst[i] = src[i];
	}
}

__host__ __device__ constexpr int CoefLd(int Kp) {
	return Kp + 4; // 2*ld mod 32 == 8 or 24: DMMA A-fragment loads without bank conflicts
}
__host__ __device__ constexpr int LdA(int Kp) {
	return (Kp + 9) / 16 * 16 + 6; // smallest ld >= Kp with ld % 16 == 6, see the model pass
}
__host__ __device__ constexpr int TileDoubles(int Kp) {
	// largest staged tile: 16 channels of table_a, table_p or one extended-table chunk
	return 16 * (LdA(Kp) > kSinExtChunkLd ? LdA(Kp) : kSinExtChunkLd);
}
__host__ __device__ inline int GwStride(int Kp) {
	// bordered matrices of the solves (packed lower triangles) share the groups' buffers: room for six
	// solving warps (eight would push the CTA over half an SM's shared memory at 48 bases)
	int const a = (6 * LdltPackedDoubles(Kp) + NGROUP - 1) / NGROUP;
	int const b = RB * (Kp > 32 ? Kp : 32); // per-group partial sums of the contractions
	int const c = 2 * TileDoubles(Kp); // two stages of table tiles per warp pair
	int const ab = a > b ? a : b;
	return ab > c ? ab : c;
}
__host__ __device__ inline int GramEntries(int Kp) {
	return Kp * (Kp + 1) / 2;
}

struct Smem {
	double *coef; // [RB][CoefLd]
	double *bvec; // [RB][Kp]
	double *trig; // [RB][ext_ld]
	double *gw; // [NGROUP][gw_stride]
	double *stat; // [NGROUP][RB][3]
	double *rmsd; // [RB]
	float *lo; // [RB]
	float *hi; // [RB]
	int *state; // [RB]
	int *cnt; // [RB] valid channels (statistics count of the latest fit)
	uint64_t *bars; // [NGROUP][4] mbarriers of the table-tile rings
	int gw_stride;
};

__device__ __forceinline__ Smem Carve(unsigned char *base, int Kp, int ext_ld) {
	Smem s;
	s.gw_stride = GwStride(Kp);
	double *d = reinterpret_cast<double *>(base);
	s.coef = d;
	d += RB * CoefLd(Kp);
	s.bvec = d;
	d += RB * Kp;
	s.trig = d;
	d += RB * ext_ld;
	s.gw = d;
	d += NGROUP * s.gw_stride;
	s.stat = d;
	d += NGROUP * RB * 3;
	s.rmsd = d;
	d += RB;
	s.bars = reinterpret_cast<uint64_t *>(d);
	d += NGROUP * 4;
	s.lo = reinterpret_cast<float *>(d);
	s.hi = s.lo + RB;
	s.state = reinterpret_cast<int *>(s.hi + RB);
	s.cnt = s.state + RB;
	return s;
}

// Gram table entry for bases i >= j: G = (+-trig[i1] +- trig[i2]) / 2; `zero` is a column of the
// trig sums that is identically zero (S_0)
__device__ unsigned GramPack(SinusoidFitParams const &p, int i, int j, int zero) {
	unsigned const pos = (static_cast<unsigned>(i) << 26) | (static_cast<unsigned>(j) << 20);
	if (p.kind[i] < 0 || p.kind[j] < 0) { // padding slot: identity row / column
		return pos | kGramPad | (i == j ? kGramPadDiag : 0u);
	}
	int const ki = p.kind[i], wi = p.wave[i], kj = p.kind[j], wj = p.wave[j];
	bool const cos_only = p.cos_only != 0;
	auto C = [cos_only](int m) {return static_cast<unsigned>(cos_only ? m : (m == 0 ? 0 : 2 * m));};
	auto S = [zero](int m) {return static_cast<unsigned>(m == 0 ? zero : 2 * (m < 0 ? -m : m) - 1);};
	unsigned i1, i2, neg = 0u;
	if (ki == 0 && kj == 0) {
		i1 = i2 = C(0);
	} else if (ki == 0 || kj == 0) {
		int const k = (ki == 0) ? kj : ki;
		int const w = (ki == 0) ? wj : wi;
		i1 = i2 = (k == 1) ? S(w) : C(w);
	} else {
		int const d = wi > wj ? wi - wj : wj - wi;
		if (ki == 1 && kj == 1) {
			i1 = C(d);
			i2 = C(wi + wj);
			neg = kGramNeg2;
		} else if (ki == 2 && kj == 2) {
			i1 = C(d);
			i2 = C(wi + wj);
		} else { // one sine (wave a), one cosine (wave b): (S_{a+b} + S_{a-b}) / 2
			int const a = (ki == 1) ? wi : wj;
			int const b = (ki == 1) ? wj : wi;
			i1 = S(a + b);
			i2 = S(a - b);
			neg = (a - b < 0) ? kGramNeg2 : 0u;
		}
	}
	return pos | neg | i1 | (i2 << 8);
}

// Data moments b[r][k] = sum_i (mask*y)[r][i] * table[i][k] over the whole row from HALF of the table.
// Every slot k is even or odd about the middle of the row, table[n-1-i][k] = +-table[i][k] (constant and
// cosines / sines of integer wave numbers on x_i = i/(n-1); Chebyshev T_m on x_i = 2i/(n-1) - 1), so with the
// folded data
//     e[r][i] = (m y)[r][i] + (m y)[r][n-1-i],   o[r][i] = (m y)[r][i] - (m y)[r][n-1-i],   i < n/2
// (exact in FP64: sums of two floats) the even slots contract e and the odd slots o: half the tensor-core
// work of the plain contraction and half the table traffic. (The table's own two halves agree to a few
// ulps, not bit for bit: an O(eps) difference like that of any other summation order.)
// acc[m][nt] += sum over the warp's channel groups, 16 channels per step; lane (r8, c4) feeds channels
// c0 + 4 c4 + s and their mirror images. A warp pair shares the channel groups and splits the four 8-row
// groups of the tile (mt0 = 0 or 2). The B tiles (16 channels x ldb doubles, contiguous in global memory)
// arrive through the pair's two-stage bulk-copy ring; the A operands (mask words and float4s of y of both
// sides per row group) are loaded one step ahead into registers.
template<int NTE, int NTO>
__device__ __forceinline__ void FoldedMoments(double (&acc)[MT][NTE + NTO][2], SinusoidFitParams const &p,
		size_t row0, uint8_t const *wmask, double const *B, int ldb, PairPipe &pp, int warp, int lane) {
	constexpr int NTC = NTE + NTO;
	int const r8 = lane >> 2;
	int const c4 = lane & 3;
	int const n = p.n;
	int const half = n / 2;
	int const mt0 = MT * (warp % (4 / MT));
	unsigned const tile_bytes = 16u * static_cast<unsigned>(ldb) * sizeof(double);
	float const *yrow[MT];
	uint8_t const *mrow[MT];
#pragma unroll
	for (int mt = 0; mt < MT; ++mt) {
		size_t row = row0 + 8 * (mt0 + mt) + r8;
		row = row < p.num_spectra ? row : p.num_spectra - 1; // (its working mask is all false)
		yrow[mt] = p.data + row * p.data_stride;
		mrow[mt] = wmask + static_cast<size_t>(8 * (mt0 + mt) + r8) * n;
	}
	unsigned use = pp.it;
	int const first = (warp / (4 / MT)) * 16;
	unsigned mw_next[MT], mr_next[MT];
	float4 yv_next[MT], yr_next[MT];
	auto load = [&](int c0) {
		int const ch = c0 + 4 * c4;
#pragma unroll
		for (int mt = 0; mt < MT; ++mt) {
			mw_next[mt] = *reinterpret_cast<unsigned const *>(mrow[mt] + ch);
			yv_next[mt] = *reinterpret_cast<float4 const *>(yrow[mt] + ch);
			mr_next[mt] = *reinterpret_cast<unsigned const *>(mrow[mt] + n - 4 - ch); // channels n-1-ch-3 .. n-1-ch
			yr_next[mt] = *reinterpret_cast<float4 const *>(yrow[mt] + n - 4 - ch);
		}
	};
	if (first < half) {
		PipeIssue(pp, use, B + static_cast<size_t>(first) * ldb, tile_bytes, lane);
		load(first);
	}
	for (int c0 = first; c0 < half; c0 += 16 * NGROUP, ++use) {
		unsigned mw[MT], mr[MT];
		float4 yv[MT], yr[MT];
#pragma unroll
		for (int mt = 0; mt < MT; ++mt) {
			mw[mt] = mw_next[mt];
			yv[mt] = yv_next[mt];
			mr[mt] = mr_next[mt];
			yr[mt] = yr_next[mt];
		}
		int const cn = c0 + 16 * NGROUP;
		if (cn < half) {
			PipeIssue(pp, use + 1u, B + static_cast<size_t>(cn) * ldb, tile_bytes, lane);
			load(cn);
		}
		double const *brow = PipeWait(pp, use) + 4 * c4 * ldb + r8;
#pragma unroll
		for (int s = 0; s < 4; ++s) {
			double b[NTC];
#pragma unroll
			for (int nt = 0; nt < NTC; ++nt) {
				b[nt] = brow[s * ldb + 8 * nt];
			}
			double ae[MT], ao[MT];
#pragma unroll
			for (int mt = 0; mt < MT; ++mt) {
				// channel ch + s and its mirror image n-1-ch-s = element 3 - s of the mirror side's float4
				bool const mf = ((mw[mt] >> (8 * s)) & 0xffu) != 0u;
				bool const mm = ((mr[mt] >> (8 * (3 - s))) & 0xffu) != 0u;
				float const yf = s == 0 ? yv[mt].x : (s == 1 ? yv[mt].y : (s == 2 ? yv[mt].z : yv[mt].w));
				float const ym = s == 0 ? yr[mt].w : (s == 1 ? yr[mt].z : (s == 2 ? yr[mt].y : yr[mt].x));
				// (selects on the bits: a NaN / Inf under the mask must not leak)
				double const df = static_cast<double>(__int_as_float(mf ? __float_as_int(yf) : 0));
				double const dm = static_cast<double>(__int_as_float(mm ? __float_as_int(ym) : 0));
				ae[mt] = df + dm;
				ao[mt] = df - dm;
			}
#pragma unroll
			for (int nt = 0; nt < NTC; ++nt) {
#pragma unroll
				for (int mt = 0; mt < MT; ++mt) {
					dmma(acc[mt][nt][0], acc[mt][nt][1], nt < NTE ? ae[mt] : ao[mt], b[nt]);
				}
			}
		}
		PipeRelease(pp, use, lane);
	}
	pp.it = use;
}

// per-warp partial sums -> shared memory -> dst[r * ld_dst + col0 + c], summed over the warps
template<int NTC>
__device__ __forceinline__ void ReducePartials(double const (&acc)[MT][NTC][2], Smem const &s,
		double *dst, int ld_dst, int col0, int warp, int lane, int tid) {
	constexpr int W = 8 * NTC;
	double *mine = s.gw + (warp / (4 / MT)) * s.gw_stride; // one [RB][W] buffer per channel group
	int const mt0 = MT * (warp % (4 / MT));
	__syncthreads(); // the buffers alias the table-tile stages other warps may still be reading
#pragma unroll
	for (int mt = 0; mt < MT; ++mt) {
#pragma unroll
		for (int nt = 0; nt < NTC; ++nt) {
			int const r = 8 * (mt0 + mt) + (lane >> 2);
			int const k = 8 * nt + 2 * (lane & 3);
			*reinterpret_cast<double2 *>(mine + r * W + k) = make_double2(acc[mt][nt][0], acc[mt][nt][1]);
		}
	}
	__syncthreads();
	for (int e = tid; e < RB * W; e += kSinThreads) {
		double tot = 0.0;
#pragma unroll
		for (int w = 0; w < NGROUP; ++w) {
			tot += s.gw[w * s.gw_stride + e];
		}
		int const r = e / W;
		dst[r * ld_dst + col0 + (e - r * W)] = tot;
	}
	__syncthreads();
}

// 16 mask bytes -> 16 bits (any non-zero byte is "true"), bit b <-> byte b
__device__ __forceinline__ unsigned MaskBits16(uint4 const m) {
	auto nib = [](unsigned w) {
		return ((__vcmpne4(w, 0u) & 0x80808080u) * 0x00204081u) >> 28;
	};
	return nib(m.x) | (nib(m.y) << 4) | (nib(m.z) << 8) | (nib(m.w) << 12);
}

// Working mask and masked column sums of the extended table of ONE spectrum, by one warp:
//   trig[e] = sum over valid channels of ext[i][e]
// The sums over all channels (and over every block of kSinSumStep channels) are constants of the
// context, so only the FLAGGED channels of a block are visited -- a few per cent of the row -- and
// subtracted; a block with more flagged than valid channels is summed directly instead. This replaces
// the contraction mask x ext on the tensor cores (a third of the kernel's FP64 work) by ~3 loads and
// 3 additions per flagged channel. Column 0 of the table is 1, so trig[0] is the valid-channel count.
// Lane l accumulates columns l, l + 32, l + 64; fixed order: deterministic.
__device__ __forceinline__ void ComplementTrigSums(SinusoidFitParams const &p, size_t row, bool live,
		uint8_t *wm, int *list, double *trig, int lane) {
	int const n = p.n;
	int const ext_ld = p.ext_ld;
	size_t const ext_chunk = static_cast<size_t>(n) * kSinExtChunkLd;
	int const steps = (n + kSinSumStep - 1) / kSinSumStep;
	bool const c1 = 32 < ext_ld, c2 = 64 < ext_ld;
	double acc[3];
	{
		double const *tot = p.ext_sums + static_cast<size_t>(steps) * ext_ld + lane;
		acc[0] = live ? tot[0] : 0.0;
		acc[1] = (live && c1) ? tot[32] : 0.0;
		acc[2] = (live && c2) ? tot[64] : 0.0;
	}
	// the channels to visit are listed across the row (tagged: subtract a flagged one, add a valid one of a
	// dense block) and their table rows gathered eight at a time per lane
	int nlist = 0;
	double const *ext_lane = p.ext_p + lane;
	auto gather = [&](int total) {
		int t = 0;
		for (; t + 8 <= total; t += 8) {
			double e0[8], e1[8], e2[8], sg[8];
#pragma unroll
			for (int q = 0; q < 8; ++q) {
				unsigned const e = static_cast<unsigned>(list[t + q]);
				double const *er = ext_lane + (e & 0x7fffffffu) * static_cast<unsigned>(kSinExtChunkLd);
				sg[q] = (e & 0x80000000u) ? 1.0 : -1.0;
				e0[q] = __ldg(er);
				e1[q] = c1 ? __ldg(er + ext_chunk) : 0.0;
				e2[q] = c2 ? __ldg(er + 2 * ext_chunk) : 0.0;
			}
#pragma unroll
			for (int q = 0; q < 8; ++q) {
				acc[0] = fma(sg[q], e0[q], acc[0]);
				acc[1] = fma(sg[q], e1[q], acc[1]);
				acc[2] = fma(sg[q], e2[q], acc[2]);
			}
		}
		for (; t < total; ++t) {
			unsigned const e = static_cast<unsigned>(list[t]);
			double const *er = ext_lane + (e & 0x7fffffffu) * static_cast<unsigned>(kSinExtChunkLd);
			double const sg = (e & 0x80000000u) ? 1.0 : -1.0;
			acc[0] = fma(sg, __ldg(er), acc[0]);
			acc[1] = fma(sg, c1 ? __ldg(er + ext_chunk) : 0.0, acc[1]);
			acc[2] = fma(sg, c2 ? __ldg(er + 2 * ext_chunk) : 0.0, acc[2]);
		}
	};
	uint8_t const *gm = p.mask + (live ? row : 0) * p.mask_stride;
	uint4 m_next = make_uint4(0u, 0u, 0u, 0u);
	if (live && 16 * lane < n) {
		m_next = __ldg(reinterpret_cast<uint4 const *>(gm + 16 * lane));
	}
	for (int st = 0; st < steps; ++st) {
		int const base = st * kSinSumStep + 16 * lane;
		uint4 const m = m_next;
		if (live && base + kSinSumStep < n) {
			m_next = __ldg(reinterpret_cast<uint4 const *>(gm + base + kSinSumStep));
		}
		bool const in = base < n;
		unsigned const valid16 = (live && in) ? MaskBits16(m) : 0u;
		if (in) {
			uint4 w;
			w.x = live ? (__vcmpne4(m.x, 0u) & 0x01010101u) : 0u;
			w.y = live ? (__vcmpne4(m.y, 0u) & 0x01010101u) : 0u;
			w.z = live ? (__vcmpne4(m.z, 0u) & 0x01010101u) : 0u;
			w.w = live ? (__vcmpne4(m.w, 0u) & 0x01010101u) : 0u;
			*reinterpret_cast<uint4 *>(wm + base) = w;
		}
		if (!live) {
			continue;
		}
		unsigned const flag16 = in ? (~valid16 & 0xffffu) : 0u;
		int const nflag = __reduce_add_sync(kFull, __popc(flag16));
		if (nflag == 0) {
			continue;
		}
		int const nin = (n - st * kSinSumStep < kSinSumStep) ? n - st * kSinSumStep : kSinSumStep;
		bool const direct = 2 * nflag > nin; // more flagged than valid channels here: add the valid ones
		if (direct) {
			double const *blk = p.ext_sums + static_cast<size_t>(st) * ext_ld + lane;
			acc[0] -= blk[0];
			acc[1] -= c1 ? blk[32] : 0.0;
			acc[2] -= c2 ? blk[64] : 0.0;
		}
		unsigned bits = direct ? valid16 : flag16;
		unsigned const tag = direct ? 0x80000000u : 0u; // the channel is ADDED (valid one of a dense block)
		int const mine = __popc(bits);
		int incl = mine;
#pragma unroll
		for (int o = 1; o < 32; o <<= 1) {
			int const v = __shfl_up_sync(kFull, incl, o);
			incl += (lane >= o) ? v : 0;
		}
		int const total = __shfl_sync(kFull, incl, 31);
		int at = nlist + incl - mine;
		while (bits != 0u) {
			list[at++] = static_cast<int>(static_cast<unsigned>(base + __ffs(bits) - 1) | tag);
			bits &= bits - 1u;
		}
		nlist += total;
		if (nlist > kSinSumStep / 2) { // the next block adds at most half a block
			__syncwarp();
			gather(nlist);
			nlist = 0;
			__syncwarp();
		}
	}
	__syncwarp();
	gather(nlist);
	__syncwarp();
	trig[lane] = acc[0];
	if (c1) {
		trig[lane + 32] = acc[1];
	}
	if (c2) {
		trig[lane + 64] = acc[2];
	}
}

// Solve of one row by this warp: fills the bordered matrix [[G, .], [b^T, .]] from the trig sums,
// blocked LDL^T (8 x 8 blocks, trailing updates as DMMA), whose border row ends up as D^-1 L^-1 b,
// then L^T x = z. Returns false (warp-uniform) when a pivot is not safely positive.
template<int NT8>
__device__ bool WarpSolveRow(unsigned const *gtab, double const *trig, double const *bvec, double *A,
		double *xout, int K, int lane) {
	constexpr int Kp = 8 * NT8;
	auto Row = [](int i) {
		return (i * (i + 1)) / 2; // packed lower triangle, see WarpLdltSolve<NT8, true>
	};
	for (int e = lane; e < Kp * (Kp + 1) / 2; e += 32) {
		unsigned const g = __ldg(gtab + e);
		double a1 = trig[g & 0xffu];
		double a2 = trig[(g >> 8) & 0xffu];
		a1 = (g & kGramNeg1) ? -a1 : a1;
		a2 = (g & kGramNeg2) ? -a2 : a2;
		double v = 0.5 * (a1 + a2);
		if (g & kGramPad) {
			v = (g & kGramPadDiag) ? 1.0 : 0.0;
		}
		A[Row(static_cast<int>(g >> 26)) + ((g >> 20) & 63u)] = v;
	}
	for (int j = lane; j < Kp; j += 32) {
		A[Row(Kp) + j] = bvec[j];
	}
	__syncwarp();
	return WarpLdltSolve<NT8, true>(A, K, xout, lane);
}

// NTE / NTO: 8-slot groups of the bases that are even / odd about the middle of the row (p.kc_p = 8 NTE)
template<int NTE, int NTO>
__global__ void __launch_bounds__(kSinThreads, kSinCtasPerSm)
SinusoidFitKernel(SinusoidFitParams const p) {
	constexpr int NT8 = NTE + NTO;
	constexpr int Kp = 8 * NT8;
	constexpr int LDC = CoefLd(Kp);
	extern __shared__ __align__(16) unsigned char smem_raw[];
	Smem const s = Carve(smem_raw, Kp, p.ext_ld);
	int const tid = threadIdx.x;
	int const lane = tid & 31;
	int const warp = tid >> 5;
	int const n = p.n;
	int const K = p.K;
	int const ext_ld = p.ext_ld;
	size_t const num_tiles = (p.num_spectra + RB - 1) / RB;
	float *wres = p.ws_residual + static_cast<size_t>(blockIdx.x) * RB * n;
	float *wbest = p.ws_best_fit ? p.ws_best_fit + static_cast<size_t>(blockIdx.x) * RB * n : nullptr;
	uint8_t *wmask = p.ws_mask + static_cast<size_t>(blockIdx.x) * RB * n;
	float const clip = p.clip_threshold_sigma;

	if (p.num_fitting_max == 0) { // baseline.cc:1295-1296: kOK, nothing written
		for (size_t row = static_cast<size_t>(blockIdx.x) * kSinThreads + tid; row < p.num_spectra;
				row += static_cast<size_t>(gridDim.x) * kSinThreads) {
			p.status[row] = kStatusOK;
			p.lsqfit_status[row] = kLsqOK;
		}
		return;
	}
	// table-tile ring of this warp's pair (group); its stages live in the group's workspace
	int const group = warp / (4 / MT);
	if (tid < NGROUP) {
		PipeInit(s.bars + 4 * tid, 4 / MT);
	}
	mbar_fence_init();
	PairPipe pp;
	pp.full = s.bars + 4 * group;
	pp.empty = pp.full + 2;
	pp.stage = s.gw + group * s.gw_stride;
	pp.stage_doubles = TileDoubles(Kp);
	pp.it = 0u;
	pp.producer = (warp % (4 / MT)) == 0;
	size_t const ext_chunk = static_cast<size_t>(n) * kSinExtChunkLd;
	int const nsolve = (NGROUP * s.gw_stride / LdltPackedDoubles(Kp) < NWARP) ?
			NGROUP * s.gw_stride / LdltPackedDoubles(Kp) : NWARP;

	for (size_t tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
		size_t const row0 = tile * RB;
		__syncthreads();
		// ---- P0: working mask + masked column sums of the extended table (column 0 = valid count):
		// "all channels minus the flagged ones", one warp per spectrum ----------------------------------
		for (int r = warp; r < RB; r += NWARP) {
			ComplementTrigSums(p, row0 + r, row0 + r < p.num_spectra, wmask + static_cast<size_t>(r) * n,
					reinterpret_cast<int *>(s.gw) + warp * kSinSumStep, s.trig + r * ext_ld, lane);
		}
		__syncthreads(); // wmask (global) and the sums are visible to the whole CTA
		if (tid < RB) {
			int const cnt = static_cast<int>(s.trig[tid * ext_ld]);
			int st = kActive;
			if (row0 + tid >= p.num_spectra) {
				st = kNoRow;
			} else if (cnt < K) {
				st = kFailedBefore; // baseline.cc:1147-1150
			}
			s.state[tid] = st;
			s.cnt[tid] = cnt;
		}
		// ---- P1: data moments b = (mask*y) x table ---------------------------------------------------
		{
			double acc[MT][NT8][2];
#pragma unroll
			for (int mt = 0; mt < MT; ++mt) {
#pragma unroll
				for (int nt = 0; nt < NT8; ++nt) {
					acc[mt][nt][0] = 0.0;
					acc[mt][nt][1] = 0.0;
				}
			}
			FoldedMoments<NTE, NTO>(acc, p, row0, wmask, p.table_p, Kp + 1, pp, warp, lane);
			ReducePartials<NT8>(acc, s, s.bvec, Kp, 0, warp, lane, tid);
		}

		// ---- clip iterations -------------------------------------------------------------------------
		for (int iter = 1; iter <= p.num_fitting_max; ++iter) {
			// (a) solve, one warp per row, as many warps as packed workspaces fit into the groups' buffers
			bool any_active = false;
			for (int r = warp; r < RB && warp < nsolve; r += nsolve) {
				if (s.state[r] != kActive) {
					continue;
				}
				any_active = true;
				bool const ok = WarpSolveRow<NT8>(p.gram_table, s.trig + r * ext_ld, s.bvec + r * Kp,
						s.gw + warp * LdltPackedDoubles(Kp), s.coef + r * LDC, K, lane);
				if (!ok) {
					if (lane == 0) {
						s.state[r] = kFallback;
						int const slot = atomicAdd(p.fallback_count, 1);
						p.fallback_rows[slot] = static_cast<int32_t>(row0 + r);
					}
				}
			}
			if (__syncthreads_or(any_active ? 1 : 0) == 0) {
				break;
			}
			// (b) model + residual + statistics: coeff x table, 16 channels per warp step; lane
			// (r8, c4) ends up with channels c0 + 4 c4 .. + 3 of rows 8 mt + r8
			{
				int cnt[MT];
				double sum[MT], sq[MT];
				int const r8 = lane >> 2;
				int const c4 = lane & 3;
				int const mt0 = MT * (warp % (4 / MT));
				bool act[MT];
				float const *yrow[MT];
				// Where the residual / model of this pass go: the tile's staging rows, except in the last
				// iteration -- no clip follows it, a row that is active now WILL be published -- where they
				// are written straight to the caller's arrays (no copy afterwards, and nothing at all when
				// the caller does not want them).
				bool const last_iter = (iter == p.num_fitting_max);
				float *res_row[MT];
				float *best_row[MT];
#pragma unroll
				for (int mt = 0; mt < MT; ++mt) {
					cnt[mt] = 0;
					sum[mt] = 0.0;
					sq[mt] = 0.0;
					int const r = 8 * (mt0 + mt) + r8;
					act[mt] = (s.state[r] == kActive);
					size_t const row = act[mt] ? row0 + r : row0;
					yrow[mt] = p.data + row * p.data_stride;
					if (last_iter) {
						res_row[mt] = p.residual ? p.residual + row * p.data_stride : nullptr;
						best_row[mt] = p.best_fit ? p.best_fit + row * p.data_stride : nullptr;
					} else {
						res_row[mt] = wres + static_cast<size_t>(r) * n;
						best_row[mt] = wbest ? wbest + static_cast<size_t>(r) * n : nullptr;
					}
				}
				// B fragment: lane (n = r8, k = c4) of the table tile staged in shared memory (16 channels
				// x ld_a). Column n of half h is channel 4 (n / 2) + 2 (n & 1) + h, so that this lane's C
				// fragments {c[0][0], c[1][0], c[0][1], c[1][1]} are the 4 consecutive channels 4 c4 .. + 3,
				// and with ld_a % 16 == 6 the rows {0, 2, 4, 6} + h of a half warp fall into disjoint banks.
				constexpr int LDA = LdA(Kp);
				int const brow = 4 * (r8 >> 1) + 2 * (r8 & 1);
				unsigned const tile_bytes = 16u * LDA * sizeof(double);
				int const first = group * 16;
				unsigned use = pp.it;
				float4 yv_next[MT];
				unsigned mw_next[MT];
				// Half of the row: the even slots give E[i], the odd slots O[i]; model[i] = E + O,
				// model[n-1-i] = E - O (see FoldedMoments). The lane's four channels ch .. ch + 3 have their
				// mirror images in the float4 at n - 4 - ch, in reverse order.
				int const half = n / 2;
				float4 yr_next[MT];
				unsigned mr_next[MT];
				auto load = [&](int c0) {
#pragma unroll
					for (int mt = 0; mt < MT; ++mt) {
						uint8_t const *wm = wmask + static_cast<size_t>(8 * (mt0 + mt) + r8) * n;
						yv_next[mt] = *reinterpret_cast<float4 const *>(yrow[mt] + c0 + 4 * c4);
						mw_next[mt] = *reinterpret_cast<unsigned const *>(wm + c0 + 4 * c4);
						yr_next[mt] = *reinterpret_cast<float4 const *>(yrow[mt] + n - 4 - (c0 + 4 * c4));
						mr_next[mt] = *reinterpret_cast<unsigned const *>(wm + n - 4 - (c0 + 4 * c4));
					}
				};
				if (first < half) {
					PipeIssue(pp, use, p.table_a + static_cast<size_t>(first) * LDA, tile_bytes, lane);
					load(first);
				}
				for (int c0 = first; c0 < half; c0 += 16 * NGROUP, ++use) {
					int const ch = c0 + 4 * c4;
					int const chm = n - 4 - ch;
					float4 yv[MT], yr[MT];
					unsigned mw[MT], mr[MT];
#pragma unroll
					for (int mt = 0; mt < MT; ++mt) {
						yv[mt] = yv_next[mt];
						mw[mt] = mw_next[mt];
						yr[mt] = yr_next[mt];
						mr[mt] = mr_next[mt];
					}
					int const cn = c0 + 16 * NGROUP;
					if (cn < half) {
						PipeIssue(pp, use + 1u, p.table_a + static_cast<size_t>(cn) * LDA, tile_bytes, lane);
						load(cn);
					}
					double ce[MT][2][2], co[MT][2][2];
#pragma unroll
					for (int mt = 0; mt < MT; ++mt) {
						ce[mt][0][0] = ce[mt][0][1] = ce[mt][1][0] = ce[mt][1][1] = 0.0;
						co[mt][0][0] = co[mt][0][1] = co[mt][1][0] = co[mt][1][1] = 0.0;
					}
					double const *tb = PipeWait(pp, use) + brow * LDA + c4;
#pragma unroll
					for (int ks = 0; ks < Kp; ks += 4) {
						double const b0 = tb[ks];
						double const b1 = tb[LDA + ks];
#pragma unroll
						for (int mt = 0; mt < MT; ++mt) {
							double const a = s.coef[(8 * (mt0 + mt) + r8) * LDC + ks + c4];
							if (ks < 8 * NTE) {
								dmma(ce[mt][0][0], ce[mt][0][1], a, b0);
								dmma(ce[mt][1][0], ce[mt][1][1], a, b1);
							} else {
								dmma(co[mt][0][0], co[mt][0][1], a, b0);
								dmma(co[mt][1][0], co[mt][1][1], a, b1);
							}
						}
					}
					PipeRelease(pp, use, lane);
#pragma unroll
					for (int mt = 0; mt < MT; ++mt) {
						if (act[mt]) {
							// forward side: channels ch .. ch + 3; mirror side: chm .. chm + 3 = images of ch + 3 .. ch
							double const e4[4] = { ce[mt][0][0], ce[mt][1][0], ce[mt][0][1], ce[mt][1][1] };
							double const o4[4] = { co[mt][0][0], co[mt][1][0], co[mt][0][1], co[mt][1][1] };
							float4 md, rd, mdm, rdm;
							md.x = static_cast<float>(e4[0] + o4[0]);
							md.y = static_cast<float>(e4[1] + o4[1]);
							md.z = static_cast<float>(e4[2] + o4[2]);
							md.w = static_cast<float>(e4[3] + o4[3]);
							mdm.x = static_cast<float>(e4[3] - o4[3]);
							mdm.y = static_cast<float>(e4[2] - o4[2]);
							mdm.z = static_cast<float>(e4[1] - o4[1]);
							mdm.w = static_cast<float>(e4[0] - o4[0]);
							rd.x = __fsub_rn(yv[mt].x, md.x);
							rd.y = __fsub_rn(yv[mt].y, md.y);
							rd.z = __fsub_rn(yv[mt].z, md.z);
							rd.w = __fsub_rn(yv[mt].w, md.w);
							rdm.x = __fsub_rn(yr[mt].x, mdm.x);
							rdm.y = __fsub_rn(yr[mt].y, mdm.y);
							rdm.z = __fsub_rn(yr[mt].z, mdm.z);
							rdm.w = __fsub_rn(yr[mt].w, mdm.w);
							if (res_row[mt] != nullptr) {
								*reinterpret_cast<float4 *>(res_row[mt] + ch) = rd;
								*reinterpret_cast<float4 *>(res_row[mt] + chm) = rdm;
							}
							if (best_row[mt] != nullptr) {
								*reinterpret_cast<float4 *>(best_row[mt] + ch) = md;
								*reinterpret_cast<float4 *>(best_row[mt] + chm) = mdm;
							}
							float const rv[8] = { rd.x, rd.y, rd.z, rd.w, rdm.x, rdm.y, rdm.z, rdm.w };
#pragma unroll
							for (int q = 0; q < 8; ++q) {
								unsigned const word = q < 4 ? mw[mt] : mr[mt];
								bool const m = ((word >> (8 * (q & 3))) & 0xffu) != 0u;
								double const v = static_cast<double>(__int_as_float(m ? __float_as_int(rv[q]) : 0));
								cnt[mt] += m ? 1 : 0;
								sum[mt] += v;
								sq[mt] = fma(v, v, sq[mt]);
							}
						}
					}
				}
				pp.it = use;
#pragma unroll
				for (int mt = 0; mt < MT; ++mt) {
#pragma unroll
					for (int o = 1; o <= 2; o <<= 1) {
						cnt[mt] += __shfl_xor_sync(kFull, cnt[mt], o);
						sum[mt] += __shfl_xor_sync(kFull, sum[mt], o);
						sq[mt] += __shfl_xor_sync(kFull, sq[mt], o);
					}
					if (c4 == 0) {
						double *st = s.stat + ((warp / (4 / MT)) * RB + 8 * (mt0 + mt) + r8) * 3;
						st[0] = static_cast<double>(cnt[mt]);
						st[1] = sum[mt];
						st[2] = sq[mt];
					}
				}
			}
			__syncthreads();
			if (tid < RB && s.state[tid] == kActive) {
				double cd = 0.0, sum = 0.0, sq = 0.0;
				for (int w = 0; w < NGROUP; ++w) {
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
			// (c) clip + downdate, one warp per row: 256 channels per step, clipped channels listed in
			// shared memory, then subtracted four at a time from lane-private partial sums
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
				// clipped channels are listed across the whole row and subtracted from the sums afterwards,
				// four table rows in flight per lane (listing per 256-channel step left most groups of four
				// nearly empty and paid an L2 round trip per step)
				constexpr int kClipListCap = 512;
				int *list = reinterpret_cast<int *>(s.gw) + warp * kClipListCap;
				int nlist = 0;
				double dt[3] = { 0.0, 0.0, 0.0 };
				double db[2] = { 0.0, 0.0 };
				int num_clipped = 0;
				auto downdate = [&](int total) {
					for (int t = 0; t < total; t += 4) {
						int ch[4];
						bool v[4];
						float y[4];
						double e0[4], e1[4], e2[4], t0[4], t1[4];
#pragma unroll
						for (int q = 0; q < 4; ++q) {
							v[q] = t + q < total;
							ch[q] = list[v[q] ? t + q : t];
						}
#pragma unroll
						for (int q = 0; q < 4; ++q) {
							double const *er = p.ext_p + static_cast<size_t>(ch[q]) * kSinExtChunkLd + lane;
							double const *tr = p.table_p + static_cast<size_t>(ch[q]) * (Kp + 1);
							y[q] = yy[ch[q]];
							e0[q] = er[0];
							e1[q] = (32 < ext_ld) ? er[ext_chunk] : 0.0;
							e2[q] = (64 < ext_ld) ? er[2 * ext_chunk] : 0.0;
							t0[q] = (lane < Kp) ? tr[lane] : 0.0;
							t1[q] = (Kp > 32 && lane + 32 < Kp) ? tr[lane + 32] : 0.0;
						}
						if (t + 4 <= total) {
#pragma unroll
							for (int q = 0; q < 4; ++q) {
								double const yd = static_cast<double>(y[q]);
								dt[0] += e0[q];
								dt[1] += e1[q];
								dt[2] += e2[q];
								db[0] = fma(yd, t0[q], db[0]);
								db[1] = fma(yd, t1[q], db[1]);
							}
						} else {
#pragma unroll
							for (int q = 0; q < 4; ++q) {
								double const yd = select_f64(v[q], static_cast<double>(y[q]));
								dt[0] += select_f64(v[q], e0[q]);
								dt[1] += select_f64(v[q], e1[q]);
								dt[2] += select_f64(v[q], e2[q]);
								db[0] = fma(yd, t0[q], db[0]);
								db[1] = fma(yd, t1[q], db[1]);
							}
						}
					}
				};
				// the loads of the next step are issued before this step is examined
				uint2 mv_next = make_uint2(0u, 0u);
				float4 ra_next = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
				float4 rb_next = ra_next;
				if (lane * 8 < n) {
					mv_next = *reinterpret_cast<uint2 const *>(wm + lane * 8);
					ra_next = *reinterpret_cast<float4 const *>(rr + lane * 8);
					rb_next = *reinterpret_cast<float4 const *>(rr + lane * 8 + 4);
				}
				for (int cb0 = 0; cb0 < n; cb0 += 256) {
					int const base = cb0 + lane * 8;
					unsigned bits = 0u;
					uint2 mv = mv_next;
					float4 const ra = ra_next;
					float4 const rb = rb_next;
					if (base + 256 < n) {
						mv_next = *reinterpret_cast<uint2 const *>(wm + base + 256);
						ra_next = *reinterpret_cast<float4 const *>(rr + base + 256);
						rb_next = *reinterpret_cast<float4 const *>(rr + base + 260);
					}
					if (base < n) {
						float const rv[8] = { ra.x, ra.y, ra.z, ra.w, rb.x, rb.y, rb.z, rb.w };
#pragma unroll
						for (int u = 0; u < 8; ++u) {
							unsigned const m = ((u < 4 ? mv.x : mv.y) >> (8 * (u & 3))) & 0xffu;
							// the last channel is never examined
							bool const c = m != 0u && base + u < n - 1 && is_clipped(rv[u], lo, hi);
							bits |= (c ? 1u : 0u) << u;
						}
						if (bits != 0u) {
#pragma unroll
							for (int u = 0; u < 8; ++u) {
								unsigned const clear = ((bits >> u) & 1u) ? ~(0xffu << (8 * (u & 3))) : ~0u;
								if (u < 4) {
									mv.x &= clear;
								} else {
									mv.y &= clear;
								}
							}
							*reinterpret_cast<uint2 *>(wm + base) = mv;
						}
					}
					if (__any_sync(kFull, bits != 0u) == 0) {
						continue;
					}
					int const mine = __popc(bits);
					int incl = mine;
#pragma unroll
					for (int o = 1; o < 32; o <<= 1) {
						int const v = __shfl_up_sync(kFull, incl, o);
						incl += (lane >= o) ? v : 0;
					}
					int const total = __shfl_sync(kFull, incl, 31);
					int at = nlist + incl - mine;
					unsigned b = bits;
					while (b != 0u) {
						list[at++] = base + __ffs(b) - 1;
						b &= b - 1u;
					}
					nlist += total;
					num_clipped += total;
					if (nlist > kClipListCap - 256) { // the next step may add up to 256 entries
						__syncwarp();
						downdate(nlist);
						nlist = 0;
						__syncwarp();
					}
				}
				__syncwarp();
				downdate(nlist);
				__syncwarp();
				if (num_clipped != 0) {
					double *trig = s.trig + r * ext_ld;
					double *bv = s.bvec + r * Kp;
#pragma unroll
					for (int q = 0; q < 3; ++q) {
						if (lane + 32 * q < ext_ld) {
							trig[lane + 32 * q] -= dt[q];
						}
					}
					if (lane < Kp) {
						bv[lane] -= db[0];
					}
					if (Kp > 32 && lane + 32 < Kp) {
						bv[lane + 32] -= db[1];
					}
				}
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
			uint8_t const *wm = wmask + static_cast<size_t>(r) * n;
			uint8_t *fm = p.final_mask + row * p.mask_stride;
			WarpCopy(reinterpret_cast<uint2 *>(fm), reinterpret_cast<uint2 const *>(wm), n / 8, lane);
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
			// (a row that was still active in the last iteration wrote both in place)
			if (p.residual != nullptr && st == kConverged) {
				WarpCopy(reinterpret_cast<float4 *>(p.residual + row * p.data_stride),
						reinterpret_cast<float4 const *>(wres + static_cast<size_t>(r) * n), n / 4, lane);
			}
			if (p.best_fit != nullptr && st == kConverged) {
				WarpCopy(reinterpret_cast<float4 *>(p.best_fit + row * p.data_stride),
						reinterpret_cast<float4 const *>(wbest + static_cast<size_t>(r) * n), n / 4, lane);
			}
			if (p.coeff != nullptr) {
				for (int k = lane; k < Kp; k += 32) {
					if (p.out_pos[k] >= 0) { // slot -> position in the caller's order; no rescale for sinusoids
						p.coeff[row * K + p.out_pos[k]] = s.coef[r * LDC + k];
					}
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

// table_p[i][k] (row stride Kp + 1) and table_a[i][k] (row stride LdA(Kp)), zero in the padding
__global__ void GatherTableKernel(double const *ctx_table, int nb_ctx, int n, int K, int Kp,
		int16_t const *use_idx, double *tables) {
	int const ldp = Kp + 1;
	int const lda = LdA(Kp);
	size_t const np = static_cast<size_t>(n) * ldp;
	size_t const total = np + static_cast<size_t>(n) * lda;
	for (size_t e = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x; e < total;
			e += static_cast<size_t>(gridDim.x) * blockDim.x) {
		size_t i;
		int k;
		if (e < np) {
			i = e / ldp;
			k = static_cast<int>(e - i * ldp);
		} else {
			i = (e - np) / lda;
			k = static_cast<int>((e - np) - i * lda);
		}
		// (slot k of the fit: a column of the context table, or padding)
		tables[e] = (k < Kp && use_idx[k] >= 0) ? ctx_table[i * nb_ctx + use_idx[k]] : 0.0;
	}
}

// Gram index table of a call (see GramPack): entry e = i (i + 1) / 2 + j for bases i >= j
__global__ void GramTableKernel(SinusoidFitParams const p, unsigned *out) {
	int const e = blockIdx.x * blockDim.x + threadIdx.x;
	if (e >= GramEntries(p.Kp)) {
		return;
	}
	int i = static_cast<int>((sqrtf(8.0f * e + 1.0f) - 1.0f) * 0.5f);
	while ((i + 1) * (i + 2) / 2 <= e) {
		++i;
	}
	while (i * (i + 1) / 2 > e) {
		--i;
	}
	out[e] = GramPack(p, i, e - i * (i + 1) / 2, p.next);
}

} // namespace

int SinusoidExtLd(int next) {
	return (next + 31) / 32 * 32;
}

int SinusoidLdA(int Kp) {
	return LdA(Kp);
}

size_t SinusoidTableDoubles(int n, int Kp) {
	return static_cast<size_t>(n) * (Kp + 1) + static_cast<size_t>(n) * LdA(Kp);
}

size_t SinusoidSharedBytes(int Kp, int next) {
	size_t doubles = static_cast<size_t>(RB) * CoefLd(Kp) + static_cast<size_t>(RB) * Kp
			+ static_cast<size_t>(RB) * SinusoidExtLd(next) + static_cast<size_t>(NGROUP) * GwStride(Kp)
			+ static_cast<size_t>(NGROUP) * RB * 3 + RB + static_cast<size_t>(NGROUP) * 4;
	return doubles * sizeof(double) + RB * (2 * sizeof(float) + 2 * sizeof(int)) + 32;
}

size_t SinusoidGramEntries(int Kp) {
	return static_cast<size_t>(GramEntries(Kp));
}

cudaError_t LaunchGatherTable(double const *ctx_table, int nb_ctx, int n, int K, int Kp,
		int16_t const *use_idx_dev, double *tables, SinusoidFitParams const &p, cudaStream_t stream) {
	GatherTableKernel<<<256, 256, 0, stream>>>(ctx_table, nb_ctx, n, K, Kp, use_idx_dev, tables);
	cudaError_t const e = cudaGetLastError();
	if (e != cudaSuccess) {
		return e;
	}
	GramTableKernel<<<(GramEntries(Kp) + 255) / 256, 256, 0, stream>>>(p,
			reinterpret_cast<unsigned *>(tables + SinusoidTableDoubles(n, Kp)));
	return cudaGetLastError();
}

template<int NTE, int NTO>
static cudaError_t LaunchNT(SinusoidFitParams const &p, int grid, cudaStream_t stream) {
	size_t const smem = SinusoidSharedBytes(p.Kp, p.next);
	cudaError_t err = cudaFuncSetAttribute(SinusoidFitKernel<NTE, NTO>,
			cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem));
	if (err != cudaSuccess) {
		return err;
	}
	// two CTAs per SM need (almost) all of the SM's shared memory: ask for the largest carve-out
	err = cudaFuncSetAttribute(SinusoidFitKernel<NTE, NTO>, cudaFuncAttributePreferredSharedMemoryCarveout,
			static_cast<int>(cudaSharedmemCarveoutMaxShared));
	if (err != cudaSuccess) {
		return err;
	}
	if (getenv("SAKURA_B200_DEBUG_OCCUPANCY") != nullptr) {
		int blocks = 0;
		cudaOccupancyMaxActiveBlocksPerMultiprocessor(&blocks, SinusoidFitKernel<NTE, NTO>, kSinThreads, smem);
		fprintf(stderr, "sakura_b200: SinusoidFitKernel<%d, %d> smem %zu B, resident CTAs per SM %d\n", NTE, NTO, smem,
				blocks);
	}
	SinusoidFitKernel<NTE, NTO><<<grid, kSinThreads, smem, stream>>>(p);
	return cudaGetLastError();
}

// The bases of a sinusoid (1, sin w, cos w ...) or Chebyshev fit are balanced between even and odd: these are
// the group counts that occur (the constant alone: no odd group).
bool SinusoidLayoutSupported(int kc_p, int Kp) {
	int const nte = kc_p / 8, nto = (Kp - kc_p) / 8;
	return kc_p % 8 == 0 && Kp % 8 == 0 && Kp <= kSinMaxKp && nte >= 1 && nte <= 3
			&& (nto == nte || nto == nte - 1);
}

cudaError_t LaunchSinusoidFit(SinusoidFitParams const &p, int grid, cudaStream_t stream) {
	if (p.ext_ld != SinusoidExtLd(p.next) || p.ext_ld > 96 || p.ld_p != p.Kp + 1
			|| p.ld_a != LdA(p.Kp) || p.n % 32 != 0 || !SinusoidLayoutSupported(p.kc_p, p.Kp)) {
		return cudaErrorInvalidValue;
	}
	switch (10 * (p.kc_p / 8) + (p.Kp - p.kc_p) / 8) {
	case 10:
		return LaunchNT<1, 0>(p, grid, stream);
	case 11:
		return LaunchNT<1, 1>(p, grid, stream);
	case 21:
		return LaunchNT<2, 1>(p, grid, stream);
	case 22:
		return LaunchNT<2, 2>(p, grid, stream);
	case 32:
		return LaunchNT<3, 2>(p, grid, stream);
	case 33:
		return LaunchNT<3, 3>(p, grid, stream);
	default:
		return cudaErrorInvalidValue;
	}
}

} // namespace sakura_b200
