// Polynomial baseline fit, the north-star hot path (sakura_LSQFitPolynomialFloat semantics,
// libsakura/src/baseline.cc:1115-1224, 1268-1329) as one persistent sm_100a kernel.
//
// One CTA (NT = 128 or 256 threads) owns one spectrum at a time. The row (float data + byte
// mask) is read from HBM exactly once with coalesced 128-bit / 32-bit loads; the data and
// the residual live in shared memory and the mask lives in registers (one bit per channel,
// <= 32 channels per thread) for all clip iterations; every output is written once.
// Algorithmic HBM traffic per spectrum: n*(4+1) in, n*(1+4) out (+ coeff/rms/status).
//
// Per row:
//   prologue  : load; count valid channels; data moments T_k = sum_valid y z^k (FP64, z = channel)
//               power sums over the flagged channels -> S_k = (all-channel constant) - flagged
//   every warp: Hankel normal equations G[j][k] = S[j+k]; in-register LDL^T solve (redundantly
//               in all warps: cheaper than a broadcast through shared memory + a barrier)
//   iteration : pass A  model (FP64 Horner) -> float, residual = y - model (float),
//                       count / sum / sum-of-squares over valid channels (FP64)
//               thresholds exactly as baseline.cc:1200-1207 (rounded to float)
//               pass B  clip predicate of baseline.cc:1088 in float; channel n-1 never examined;
//                       S_k, T_k contributions of the clipped channels only (downdate, the
//                       reference's own strategy: numeric_operation.cc:283-314, 448-479)
//               two block barriers per iteration (statistics, moments)
//
// What differs from the reference's arithmetic (see DESIGN.md "Arithmetic contract"): the
// Gram matrix is assembled from 2K-1 power sums instead of K*K per-channel products, sums are
// tree- instead of sequentially ordered, and the solve is LDL^T in natural order with a
// full-pivot rank-revealing fallback for degenerate rows. All are O(eps)-level perturbations
// of an FP64 computation; masks, rms and float residuals are checked against the oracle.
#include "lsqfit_poly.cuh"
#include "device_common.cuh"
#include "poly_device.cuh"

#include <float.h>
#include <math.h>
#include <stdlib.h>
#include <map>
#include <mutex>

namespace sakura_b200 {

namespace {


template<int K, int NT>
struct Shared {
	static constexpr int NW = NT / 32;
	static constexpr int NV = 3 * K - 1; // 2K-1 power sums + K data moments
	double stat[2][NW][3]; // count,sum,sq partials, double-buffered by iteration parity
	double wtot[2][NW][NV]; // per-warp moment totals, double-buffered
	double st[NW][NV + 1]; // per-warp copy of S (2K-1) and T (K) for the in-register solve
	double cx[K]; // fallback solver result
	double A[K * K]; // fallback solver workspace
	int perm[2 * K];
	int icnt[NW];
	int wclip[2][NW]; // per-warp number of channels clipped in this pass
};

constexpr int kSlotGroups = 8; // float4 groups per thread: 32 channel slots, one mask bit each


// ---- moments over the channels whose bit is set in `bits` (this thread's channel slots) ----
// Accumulates sum x^k (k < 2K-1) and, if WITH_T, sum y x^k (k < K), x = z*h, then reduces the
// per-lane partials inside the warp by transposing them through `scratch` (NV rows of 33
// doubles): lane k adds row k over the lanes that had any bit, in lane order -- deterministic,
// no atomics, no shuffles. Result: wtot[k] for k < NV (zeros when the warp has no bit set).
template<int K, int NT, bool WITH_T>
__device__ __forceinline__ void WarpBitMoments(uint32_t bits, double zbase, double h,
		float const *sdata, int tid, double *scratch, double *wtot) {
	constexpr int NS = 2 * K - 1;
	constexpr int NV = 3 * K - 1;
	int const lane = tid & 31;
	uint32_t const active = __ballot_sync(0xffffffffu, bits != 0u);
	if (active == 0u) {
		if (lane < NV) {
			wtot[lane] = 0.0;
		}
		return;
	}
	if (bits != 0u) {
		double ps[NS];
		double pt[K];
#pragma unroll
		for (int k = 0; k < NS; ++k) {
			ps[k] = 0.0;
		}
#pragma unroll
		for (int k = 0; k < K; ++k) {
			pt[k] = 0.0;
		}
		do {
			int const s = __ffs(bits) - 1;
			bits &= bits - 1u;
			uint32_t const off = static_cast<uint32_t>((s >> 2) * (4 * NT) + (s & 3));
			double const x = (zbase + u32_to_double(off)) * h;
			double yd = 0.0;
			if (WITH_T) {
				yd = static_cast<double>(sdata[4 * tid + off]);
			}
			double pw = 1.0;
#pragma unroll
			for (int k = 0; k < NS; ++k) {
				ps[k] += pw;
				if (WITH_T && k < K) {
					pt[k] = fma(yd, pw, pt[k]);
				}
				pw *= x;
			}
		} while (bits != 0u);
#pragma unroll
		for (int k = 0; k < NS; ++k) {
			scratch[k * 33 + lane] = ps[k];
		}
		if (WITH_T) {
#pragma unroll
			for (int k = 0; k < K; ++k) {
				scratch[(NS + k) * 33 + lane] = pt[k];
			}
		}
	}
	__syncwarp();
	if (lane < NV) {
		double tot = 0.0;
		if (WITH_T || lane < NS) {
			double const *rowp = scratch + lane * 33;
			for (uint32_t m = active; m != 0u; m &= m - 1u) {
				tot += rowp[__ffs(m) - 1];
			}
		}
		wtot[lane] = tot;
	}
	__syncwarp();
}


// Full-pivot LU within one warp on a K x K matrix in shared memory; same algorithm as
// lsq_device.cuh::SolveFullPivLU (Eigen FullPivLU semantics), used for degenerate rows.
template<int K>
__device__ void WarpSolveFullPivLU(double const *S, double const *T, double *A, int *perm,
		double *cx) {
	int const lane = threadIdx.x & 31;
	int *row_tr = perm;
	int *col_tr = perm + K;
	for (int e = lane; e < K * K; e += 32) {
		A[e] = S[e / K + e % K];
	}
	double c = (lane < K) ? T[lane] : 0.0; // lane i holds c[i]
	__syncwarp();
	double maxpivot = 0.0;
	int nonzero_pivots = K;
	for (int k = 0; k < K; ++k) {
		int const m = K - k;
		double best_v = -1.0;
		int best_i = 0x7fffffff;
		for (int idx = lane; idx < m * m; idx += 32) {
			int const j = k + idx / m;
			int const i = k + idx % m;
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
		if (best_v == 0.0) {
			nonzero_pivots = k;
			if (lane >= k && lane < K) {
				row_tr[lane] = lane;
				col_tr[lane] = lane;
			}
			break;
		}
		maxpivot = fmax(maxpivot, best_v);
		int const pcol = k + best_i / m;
		int const prow = k + best_i % m;
		if (lane == 0) {
			row_tr[k] = prow;
			col_tr[k] = pcol;
		}
		if (prow != k) {
			if (lane < K) {
				double const t = A[k * K + lane];
				A[k * K + lane] = A[prow * K + lane];
				A[prow * K + lane] = t;
			}
			double const ck = __shfl_sync(0xffffffffu, c, k);
			double const cp = __shfl_sync(0xffffffffu, c, prow);
			if (lane == k) {
				c = cp;
			} else if (lane == prow) {
				c = ck;
			}
			__syncwarp();
		}
		if (pcol != k) {
			if (lane < K) {
				double const t = A[lane * K + k];
				A[lane * K + k] = A[lane * K + pcol];
				A[lane * K + pcol] = t;
			}
			__syncwarp();
		}
		if (k < K - 1) {
			double const pivot = A[k * K + k];
			double l = 0.0;
			if (lane > k && lane < K) {
				l = A[lane * K + k] / pivot;
				A[lane * K + k] = l;
			}
			double const ck = __shfl_sync(0xffffffffu, c, k);
			if (lane > k && lane < K) {
				c = fma(-l, ck, c);
			}
			__syncwarp();
			int const mm = m - 1;
			for (int idx = lane; idx < mm * mm; idx += 32) {
				int const i = k + 1 + idx / mm;
				int const j = k + 1 + idx % mm;
				A[i * K + j] = fma(-A[i * K + k], A[k * K + j], A[i * K + j]);
			}
			__syncwarp();
		}
	}
	__syncwarp();
	double const threshold = fabs(maxpivot) * (DBL_EPSILON * static_cast<double>(K));
	int rank = 0;
	for (int i = 0; i < nonzero_pivots; ++i) {
		rank += (fabs(A[i * K + i]) > threshold) ? 1 : 0;
	}
	for (int j = rank - 1; j >= 0; --j) {
		double cj = c;
		if (lane == j) {
			cj = c / A[j * K + j];
			c = cj;
		}
		cj = __shfl_sync(0xffffffffu, cj, j);
		if (lane < j) {
			c = fma(-A[lane * K + j], cj, c);
		}
	}
	if (lane < K) {
		cx[lane] = 0.0;
	}
	__syncwarp();
	if (lane == 0) {
		int q[K];
		for (int i = 0; i < K; ++i) {
			q[i] = i;
		}
		for (int k = 0; k < K; ++k) {
			int const t = q[k];
			q[k] = q[col_tr[k]];
			q[col_tr[k]] = t;
		}
		for (int i = 0; i < K; ++i) {
			row_tr[i] = q[i];
		}
	}
	__syncwarp();
	if (lane < rank) {
		cx[row_tr[lane]] = c;
	}
	__syncwarp();
}

template<int K, int NT>
__host__ __device__ constexpr size_t ResidualBytes(int n) {
	// the residual buffer doubles as per-warp reduction scratch: NW x NV x 33 doubles
	size_t const need = static_cast<size_t>(NT / 32) * (3 * K - 1) * 33 * sizeof(double);
	size_t const res = static_cast<size_t>(n) * 4;
	size_t const m = res > need ? res : need;
	return (m + 15) / 16 * 16;
}

template<int K, int NT>
__host__ __device__ constexpr size_t SharedOffset(int n) {
	size_t const off = (static_cast<size_t>(n) * 4 + 15) / 16 * 16 + ResidualBytes<K, NT>(n);
	return (off + 15) / 16 * 16;
}

constexpr int kMinBlocks128 = 6;
constexpr int kMinBlocks256 = 3;

// Every warp keeps the current S (2K-1) and T (K) in its own shared-memory slot, applies the
// block-wide moment totals to it (lane k owns entry k), reads all of it back and solves in
// registers. All warps compute identical bits (same operands, same order), so no barrier and no
// broadcast is needed afterwards. `cx` = solution in x = i/(n-1); `cz` = c_k h^k for Horner in
// the channel index. Returns false when the full-pivot fallback has to be used.
template<int K, int NT, bool FIRST>
__device__ __forceinline__ bool UpdateAndSolve(Shared<K, NT> &sh, PolyConst const &pc, int buf,
		bool complement, int warp, int lane, double *cz) {
	constexpr int NW = NT / 32;
	constexpr int NS = 2 * K - 1;
	constexpr int NV = 3 * K - 1;
	double *mine = sh.st[warp];
	if (lane < NV) {
		double tot = 0.0;
#pragma unroll
		for (int w = 0; w < NW; ++w) {
			tot += sh.wtot[buf][w][lane];
		}
		double v;
		if (FIRST) {
			if (lane < NS) {
				v = complement ? (pc.s_all[lane] - tot) : tot;
			} else {
				v = tot * pc.hpow[lane - NS];
			}
		} else {
			v = mine[lane] - tot;
		}
		mine[lane] = v;
	}
	__syncwarp();
	double S[NS];
	double T[K];
#pragma unroll
	for (int k = 0; k < NS; ++k) {
		S[k] = mine[k];
	}
#pragma unroll
	for (int k = 0; k < K; ++k) {
		T[k] = mine[NS + k];
	}
	double cx[K];
	bool const ok = SolveLDLT<K>(S, T, cx);
#pragma unroll
	for (int k = 0; k < K; ++k) {
		cz[k] = cx[k] * pc.hpow[k];
	}
	if (warp == 0 && lane == 0) {
#pragma unroll
		for (int k = 0; k < K; ++k) {
			sh.cx[k] = cx[k]; // read back only when the row's outputs are written
		}
	}
	return ok;
}

// FULL: n == 4 * NT * kSlotGroups (4096 channels at 128 threads, 8192 at 256), every thread owns
// all 32 channel slots and the per-group bounds checks disappear.
// CAL: position-switch calibration (and NaN/Inf flagging) applied while the row is loaded; only
// instantiated together with FULL, other shapes calibrate in a pre-pass (capi_core.cu).
template<int K, int NT, bool FULL, bool CAL>
__global__ void __launch_bounds__(NT, (NT == 64 ? (K >= 7 ? 8 : 12) : NT == 128 ? (K >= 7 ? 4 : kMinBlocks128) :
		(K >= 7 ? 2 : kMinBlocks256)))
PolyFitKernel(PolyFitParams const p, PolyConst const pc) {
	constexpr int NW = NT / 32;
	constexpr int NS = 2 * K - 1;
	constexpr int NV = 3 * K - 1;
	extern __shared__ __align__(16) unsigned char smem_raw[];
	int const n = p.n;
	int const ng = n / 4;
	// layout: data[n] | residual[n] (also reduction scratch) | Shared
	float *sdata = reinterpret_cast<float *>(smem_raw);
	float *sres = reinterpret_cast<float *>(smem_raw + (static_cast<size_t>(n) * 4 + 15) / 16 * 16);
	Shared<K, NT> &sh = *reinterpret_cast<Shared<K, NT> *>(smem_raw + SharedOffset<K, NT>(n));

	int const tid = threadIdx.x;
	int const lane = tid & 31;
	int const warp = tid >> 5;
	// The residual is dead whenever moments are reduced (before the first fit, and between a
	// clip pass and the next model evaluation), so the two share storage.
	double *scratch = reinterpret_cast<double *>(sres) + warp * (NV * 33);
	double const zbase = static_cast<double>(4 * tid);
	float const clip = p.clip_threshold_sigma;
	// bit s = 4*j + e of a thread's mask word <-> channel 4*(tid + NT*j) + e
	int const nslots = (tid < ng) ? 4 * ((ng - 1 - tid) / NT + 1) : 0;
	uint32_t const slot_mask = (nslots >= 32) ? 0xffffffffu : ((1u << nslots) - 1u);

	// row list (rows handed over by the screening kernel, lsqfit_poly_screen.cu) or all rows
	size_t const num_rows = (p.row_count != nullptr) ? static_cast<size_t>(*p.row_count) : p.num_spectra;
	for (size_t it = blockIdx.x; it < num_rows; it += gridDim.x) {
		size_t const row = (p.row_index != nullptr) ? static_cast<size_t>(p.row_index[it]) : it;
		float4 const *gdata = reinterpret_cast<float4 const *>(p.data + row * p.data_stride);
		uint32_t const *gmask = reinterpret_cast<uint32_t const *>(p.mask + row * p.mask_stride);
		// pull this CTA's next row into L2 while this one is being fitted (one bulk prefetch each
		// for data and mask), so that the next prologue's loads hit L2 instead of HBM
		if (tid == 0 && it + gridDim.x < num_rows) {
			size_t const nrow = (p.row_index != nullptr) ?
					static_cast<size_t>(p.row_index[it + gridDim.x]) : it + gridDim.x;
			asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;"
					:: "l"(p.data + nrow * p.data_stride), "r"(static_cast<unsigned>(n) * 4u) : "memory");
			uint8_t const *nmask = p.mask + nrow * p.mask_stride;
			if (n % 16 == 0 && (reinterpret_cast<uintptr_t>(nmask) & 15u) == 0u) {
				asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;"
						:: "l"(nmask), "r"(static_cast<unsigned>(n)) : "memory");
			}
		}
		__syncthreads(); // the previous row's shared state is dead

		// ---- prologue: load the row, valid bits, data moments in z = channel index ----------
		// A thread's channels are zb + u with zb = 4*tid and u = 4*NT*j + e (j < 8, e < 4): the
		// offsets u and their powers are compile-time constants, so local moments sum y u^k cost
		// one FMA each with an immediate operand; they are shifted to the global variable
		// z = zb + u once per thread (binomial transform, no cancellation: all terms >= 0).
		uint32_t vbits = 0u;
		double t[K];
#pragma unroll
		for (int k = 0; k < K; ++k) {
			t[k] = 0.0;
		}
#pragma unroll
		for (int j = 0; j < kSlotGroups; ++j) {
			int const g = tid + NT * j;
			if (FULL || g < ng) {
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
				reinterpret_cast<float4 *>(sdata)[g] = y;
				uint32_t const nib = (m | (m >> 7) | (m >> 14) | (m >> 21)) & 0xfu;
				vbits |= nib << (4 * j);
				float const ye[4] = { y.x, y.y, y.z, y.w };
#pragma unroll
				for (int e = 0; e < 4; ++e) {
					// select (on the bits), not multiply: a NaN/Inf under the mask must not leak
					double const yd = static_cast<double>(select_or_zero(ye[e], nib & (1u << e)));
					double const u = static_cast<double>(4 * NT * j + e);
					double upow = 1.0;
#pragma unroll
					for (int k = 0; k < K; ++k) {
						t[k] = (k == 0) ? (t[0] + yd) : fma(yd, upow, t[k]);
						upow *= u;
					}
				}
			}
		}
#pragma unroll
		for (int i = 0; i < K - 1; ++i) {
#pragma unroll
			for (int j = K - 1; j > i; --j) {
				t[j] = fma(zbase, t[j - 1], t[j]);
			}
		}
		int const cnt = warp_sum(__popc(vbits));
		if (lane == 0) {
			sh.icnt[warp] = cnt;
		}
		__syncthreads();
		int num_valid = 0;
#pragma unroll
		for (int w = 0; w < NW; ++w) {
			num_valid += sh.icnt[w];
		}

		if (p.num_fitting_max == 0) { // baseline.cc:1295-1296: kOK, nothing written
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

		// ---- first normal equations ------------------------------------------------------------
		// S over the valid channels = constant over all channels - sum over the flagged ones
		// (or the direct sum when more than half of the row is flagged).
		bool const complement = (2 * num_valid >= n);
		WarpBitMoments<K, NT, false>(complement ? (~vbits & slot_mask) : vbits, zbase, pc.h, sdata,
				tid, scratch, sh.wtot[0][warp]);
#pragma unroll
		for (int k = 0; k < K; ++k) {
			t[k] = warp_sum(t[k]);
		}
		__syncwarp(); // the T rows of wtot were zeroed by other lanes just above
		if (lane == 0) {
#pragma unroll
			for (int k = 0; k < K; ++k) {
				sh.wtot[0][warp][NS + k] = t[k];
			}
		}
		__syncthreads();
		// d: model coefficients re-centred on this thread's first channel (see pass A)
		double d[K];
		bool solved;
		{
			double cz[K];
			solved = UpdateAndSolve<K, NT, true>(sh, pc, 0, complement, warp, lane, cz);
			TaylorShift<K>(cz, zbase, d);
		}

		// ---- clip iterations ----------------------------------------------------------------------
		int num_unmasked = n; // baseline.cc:1152 (sic)
		double rms_d = 0.0;
		bool failed = false;
		for (int iter = 1; iter <= p.num_fitting_max; ++iter) {
			if (num_unmasked < K) { // baseline.cc:1162-1165
				failed = true;
				break;
			}
			if (!solved) { // block-uniform: every warp solved the same system
				if (warp == 0) {
					WarpSolveFullPivLU<K>(sh.st[0], sh.st[0] + NS, sh.A, sh.perm, sh.cx);
				}
				__syncthreads();
				double cz[K];
#pragma unroll
				for (int k = 0; k < K; ++k) {
					cz[k] = sh.cx[k] * pc.hpow[k];
				}
				TaylorShift<K>(cz, zbase, d);
			}
			// pass A: model, residual, statistics
			// The model polynomial is re-centred on this thread's first channel (Taylor shift,
			// K(K-1)/2 FMAs per thread) so that the Horner variable of every channel is a
			// compile-time constant: no per-channel index arithmetic or int->double conversion.
			double sum = 0.0, sq = 0.0;
#pragma unroll
			for (int j = 0; j < kSlotGroups; ++j) {
				int const g = tid + NT * j;
				if (FULL || g < ng) {
					float4 const y = reinterpret_cast<float4 const *>(sdata)[g];
					float const ye[4] = { y.x, y.y, y.z, y.w };
					float re[4];
#pragma unroll
					for (int e = 0; e < 4; ++e) {
						double const u = static_cast<double>(4 * NT * j + e);
						double acc = d[K - 1];
#pragma unroll
						for (int k = K - 2; k >= 0; --k) {
							acc = fma(acc, u, d[k]);
						}
						float const mf = static_cast<float>(acc);
						float const r = __fsub_rn(ye[e], mf);
						re[e] = r;
						// flagged channels contribute +0.0: one integer select on the float, then
						// unconditional FP64 accumulation (a predicated DADD/DFMA costs four selects)
						double const rd = static_cast<double>(select_or_zero(r, vbits & (1u << (4 * j + e))));
						sum += rd;
						sq = fma(rd, rd, sq);
					}
					reinterpret_cast<float4 *>(sres)[g] = make_float4(re[0], re[1], re[2], re[3]);
				}
			}
			int c = warp_sum(__popc(vbits));
			sum = warp_sum(sum);
			sq = warp_sum(sq);
			int const buf = iter & 1;
			if (lane == 0) {
				sh.stat[buf][warp][0] = static_cast<double>(c);
				sh.stat[buf][warp][1] = sum;
				sh.stat[buf][warp][2] = sq;
			}
			__syncthreads();
			double cd = 0.0;
			sum = 0.0;
			sq = 0.0;
#pragma unroll
			for (int w = 0; w < NW; ++w) {
				cd += sh.stat[buf][w][0];
				sum += sh.stat[buf][w][1];
				sq += sh.stat[buf][w][2];
			}
			c = static_cast<int>(cd);
			ClipThresholds const th = make_thresholds(c, sum, sq, clip);
			rms_d = th.rms;
			if (iter == p.num_fitting_max) {
				break;
			}
			// pass B: clip. The last channel is never examined (baseline.cc:1081-1096).
			uint32_t clipped = 0u;
			// (r-lo)*(hi-r) < 0 in float (baseline.cc:1088) is the same predicate as
			// (r < lo) || (r > hi) unless the product can underflow to zero; that is excluded for
			// the whole row when neither threshold is tiny and the window is not degenerate:
			// then |r-lo| >= 2^-104 for r != lo and the other factor exceeds hi-lo >= 2^-40.
			bool const by_compare = fabsf(th.lower) >= 0x1p-80f && fabsf(th.upper) >= 0x1p-80f
					&& fabsf(th.lower) <= FLT_MAX && fabsf(th.upper) <= FLT_MAX
					&& __fsub_rn(th.upper, th.lower) >= 0x1p-40f;
			if (by_compare) {
#pragma unroll
				for (int j = 0; j < kSlotGroups; ++j) {
					int const g = tid + NT * j;
					if (FULL || g < ng) {
						float4 const r4 = reinterpret_cast<float4 const *>(sres)[g];
						float const re[4] = { r4.x, r4.y, r4.z, r4.w };
#pragma unroll
						for (int e = 0; e < 4; ++e) {
							bool const c = (re[e] < th.lower) || (re[e] > th.upper);
							clipped |= c ? (1u << (4 * j + e)) : 0u;
						}
					}
				}
			} else {
#pragma unroll
				for (int j = 0; j < kSlotGroups; ++j) {
					int const g = tid + NT * j;
					if (FULL || g < ng) {
						float4 const r4 = reinterpret_cast<float4 const *>(sres)[g];
						float const re[4] = { r4.x, r4.y, r4.z, r4.w };
#pragma unroll
						for (int e = 0; e < 4; ++e) {
							if (is_clipped(re[e], th.lower, th.upper)) {
								clipped |= (1u << (4 * j + e));
							}
						}
					}
				}
			}
			// channel n-1 is never examined
			if (((n - 4) >> 2) % NT == tid) {
				clipped &= ~(8u << (4 * (((n - 4) >> 2) / NT)));
			}
			clipped &= vbits;
			vbits &= ~clipped;
			int const nclip = warp_sum(__popc(clipped));
			// A warp only turns its slice of the residual buffer into scratch when it clipped
			// something, i.e. when another fit (which rewrites the residual) is certain to follow.
			// Other warps may still be reading the residual in pass B: the scratch slices of
			// different warps are disjoint, but they alias residual channels owned by other
			// threads, hence the barrier.
			if (__syncthreads_or(nclip) == 0) {
				break; // nothing clipped anywhere: the fit has converged (baseline.cc:1211-1213)
			}
			WarpBitMoments<K, NT, true>(clipped, zbase, pc.h, sdata, tid, scratch, sh.wtot[buf][warp]);
			if (lane == 0) {
				sh.wclip[buf][warp] = nclip;
			}
			__syncthreads();
			int num_clipped = 0;
#pragma unroll
			for (int w = 0; w < NW; ++w) {
				num_clipped += sh.wclip[buf][w];
			}
			num_unmasked = c - num_clipped;
			if (num_unmasked >= K) {
				double cz[K];
				solved = UpdateAndSolve<K, NT, false>(sh, pc, buf, false, warp, lane, cz);
				TaylorShift<K>(cz, zbase, d);
			}
		}

		// ---- final mask (bits -> bytes) ------------------------------------------------------------
		uint32_t *gfmask = reinterpret_cast<uint32_t *>(p.final_mask + row * p.mask_stride);
		{
#pragma unroll
			for (int j = 0; j < kSlotGroups; ++j) {
				int const g = tid + NT * j;
				if (FULL || g < ng) {
					uint32_t const b = (vbits >> (4 * j)) & 0xfu;
					gfmask[g] = (b & 1u) | ((b & 2u) << 7) | ((b & 4u) << 14) | ((b & 8u) << 21);
				}
			}
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
		// ---- outputs (baseline.cc:1310-1328) ---------------------------------------------------
		// The residual in shared memory is intact here: the reduction scratch that aliases it
		// is only written after a clip pass that clipped something, and then another fit follows.
		if (p.residual != nullptr && p.best_fit == nullptr) {
			float4 *outr = reinterpret_cast<float4 *>(p.residual + row * p.data_stride);
#pragma unroll
			for (int j = 0; j < kSlotGroups; ++j) {
				int const g = tid + NT * j;
				if (FULL || g < ng) {
					outr[g] = reinterpret_cast<float4 const *>(sres)[g];
				}
			}
		} else if (p.best_fit != nullptr) {
			float4 *outr = p.residual ? reinterpret_cast<float4 *>(p.residual + row * p.data_stride) :
					nullptr;
			float4 *outb = reinterpret_cast<float4 *>(p.best_fit + row * p.data_stride);
			// d is the one pass A last used: identical arithmetic, identical model values
#pragma unroll
			for (int j = 0; j < kSlotGroups; ++j) {
				int const g = tid + NT * j;
				if (FULL || g < ng) {
					float4 const y = reinterpret_cast<float4 const *>(sdata)[g];
					float const ye[4] = { y.x, y.y, y.z, y.w };
					float me[4], re[4];
#pragma unroll
					for (int e = 0; e < 4; ++e) {
						double const u = static_cast<double>(4 * NT * j + e);
						double acc = d[K - 1];
#pragma unroll
						for (int k = K - 2; k >= 0; --k) {
							acc = fma(acc, u, d[k]);
						}
						me[e] = static_cast<float>(acc);
						re[e] = __fsub_rn(ye[e], me[e]);
					}
					outb[g] = make_float4(me[0], me[1], me[2], me[3]);
					if (outr != nullptr) {
						outr[g] = make_float4(re[0], re[1], re[2], re[3]);
					}
				}
			}
		}
		if (p.coeff != nullptr && tid < K) {
			if (p.chebyshev) {
				// the fit was done in the (identical) polynomial space; report it in the
				// Chebyshev basis the context was created with
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
			p.rms[row] = static_cast<float>(rms_d);
			p.status[row] = kStatusOK;
			p.lsqfit_status[row] = kLsqOK;
			if (p.flags != nullptr) {
				p.flags[row] = 3;
			}
		}
	}
}

template<int K, int NT>
size_t PolySharedBytes(int n) {
	return SharedOffset<K, NT>(n) + sizeof(Shared<K, NT>) + 16;
}

} // namespace

PolyConst const &GetPolyConst(int n) {
	static std::mutex mu;
	static std::map<int, PolyConst> cache;
	std::lock_guard<std::mutex> lock(mu);
	auto it = cache.find(n);
	if (it != cache.end()) {
		return it->second;
	}
	PolyConst pc;
	double const max_x = static_cast<double>(n - 1);
	pc.h = 1.0 / max_x;
	long double hp = 1.0L;
	for (int k = 0; k < kMaxMoments; ++k) {
		pc.hpow[k] = static_cast<double>(hp);
		hp *= static_cast<long double>(pc.h);
	}
	// sum_i x_i^k with x_i = fl(i*h) exactly as the device forms it, Kahan-accumulated in
	// extended precision so that the constant is correctly rounded to double
	long double acc[kMaxMoments];
	long double comp[kMaxMoments];
	for (int k = 0; k < kMaxMoments; ++k) {
		acc[k] = 0.0L;
		comp[k] = 0.0L;
	}
	for (int i = 0; i < n; ++i) {
		double const x = static_cast<double>(i) * pc.h;
		long double pw = 1.0L;
		for (int k = 0; k < kMaxMoments; ++k) {
			long double const yk = pw - comp[k];
			long double const tk = acc[k] + yk;
			comp[k] = (tk - acc[k]) - yk;
			acc[k] = tk;
			pw *= static_cast<long double>(x);
		}
	}
	for (int k = 0; k < kMaxMoments; ++k) {
		pc.s_all[k] = static_cast<double>(acc[k]);
	}
	{
		// P[j][k]: coefficient of x^k in T_j(2x - 1), by T_{j+1} = 2 (2x - 1) T_j - T_{j-1};
		// cheb = (P^T)^-1, upper triangular, each column from a back substitution
		long double P[kMaxPolyK][kMaxPolyK];
		for (int j = 0; j < kMaxPolyK; ++j) {
			for (int k = 0; k < kMaxPolyK; ++k) {
				P[j][k] = 0.0L;
			}
		}
		P[0][0] = 1.0L;
		P[1][0] = -1.0L;
		P[1][1] = 2.0L;
		for (int j = 1; j + 1 < kMaxPolyK; ++j) {
			for (int k = 0; k <= j + 1; ++k) {
				long double v = -2.0L * P[j][k] - P[j - 1][k];
				if (k > 0) {
					v += 4.0L * P[j][k - 1];
				}
				P[j + 1][k] = v;
			}
		}
		for (int col = 0; col < kMaxPolyK; ++col) {
			// solve sum_{j>=k} P[j][k] a_j = delta(k, col) for a
			long double a[kMaxPolyK];
			for (int k = kMaxPolyK - 1; k >= 0; --k) {
				long double rhs = (k == col) ? 1.0L : 0.0L;
				for (int j = k + 1; j < kMaxPolyK; ++j) {
					rhs -= P[j][k] * a[j];
				}
				a[k] = rhs / P[k][k];
			}
			for (int j = 0; j < kMaxPolyK; ++j) {
				pc.cheb[j][col] = static_cast<double>(a[j]);
			}
		}
	}
	for (int j = 0; j < kMaxPolyK; ++j) {
		for (int k = 0; k < kMaxPolyK; ++k) {
			// binom[j][k] = (-1)^j sum_{i<=k} (-1)^i C(k,i) (i+1)^j (Newton series of u^j at -1, -2, ...)
			long long acc = 0;
			long long c = 1; // C(k, i)
			for (int i = 0; i <= k; ++i) {
				long long pw = 1;
				for (int e = 0; e < j; ++e) {
					pw *= (i + 1);
				}
				acc += ((i & 1) ? -1 : 1) * c * pw;
				c = c * (k - i) / (i + 1);
			}
			pc.binom[j][k] = (k <= j) ? static_cast<double>((j & 1) ? -acc : acc) : 0.0;
		}
	}
	for (int q = 0; q < 32; ++q) {
		pc.node[q] = static_cast<double>(0.5L + 0.5L * cosl((2 * q + 1) * 3.14159265358979323846264338327950288L / 64.0L));
	}
	for (int k = 0; k < kMaxPolyK; ++k) {
		pc.ez[k] = static_cast<double>(1.000001L / cosl(k * 3.14159265358979323846264338327950288L / 64.0L));
	}
	double factor = 1.0;
	for (int k = 0; k < kMaxPolyK; ++k) {
		pc.factor[k] = factor;
		factor *= max_x;
	}
	return cache.emplace(n, pc).first->second;
}

namespace {

template<int K, int NT, bool FULL, bool CAL>
cudaError_t LaunchKNFC(PolyFitParams const &p, int num_sms, cudaStream_t stream) {
	size_t const smem = PolySharedBytes<K, NT>(p.n);
	static int per_sm_cache_n = -1;
	static int per_sm_cache = 0;
	cudaError_t err = cudaFuncSetAttribute(PolyFitKernel<K, NT, FULL, CAL>,
			cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem));
	if (err != cudaSuccess) {
		return err;
	}
	int per_sm = per_sm_cache;
	if (per_sm_cache_n != p.n) {
		err = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, PolyFitKernel<K, NT, FULL, CAL>, NT,
				smem);
		if (err != cudaSuccess) {
			return err;
		}
		per_sm_cache = per_sm;
		per_sm_cache_n = p.n;
	}
	if (per_sm < 1) {
		return cudaErrorInvalidConfiguration;
	}
	size_t grid = static_cast<size_t>(num_sms) * per_sm;
	if (grid > p.num_spectra) {
		grid = p.num_spectra;
	}
	PolyFitKernel<K, NT, FULL, CAL><<<static_cast<unsigned>(grid), NT, smem, stream>>>(p,
			GetPolyConst(p.n));
	return cudaGetLastError();
}

template<int K, int NT>
cudaError_t LaunchKN(PolyFitParams const &p, int num_sms, cudaStream_t stream) {
	if (p.n == 4 * NT * kSlotGroups) {
		if (p.cal.reference != nullptr) {
			return LaunchKNFC<K, NT, true, true>(p, num_sms, stream);
		}
		return LaunchKNFC<K, NT, true, false>(p, num_sms, stream);
	}
	if (p.cal.reference != nullptr) {
		return cudaErrorInvalidValue; // see PolyFusedCalibrationSupported
	}
	return LaunchKNFC<K, NT, false, false>(p, num_sms, stream);
}

template<int K>
cudaError_t LaunchK(PolyFitParams const &p, int num_sms, cudaStream_t stream) {
	if (p.n == 32 * 64) { // 2048 channels (the e2e shape): 64 threads own all of their 32 slots
		return LaunchKN<K, 64>(p, num_sms, stream);
	}
	if (p.n <= 32 * 128) {
		return LaunchKN<K, 128>(p, num_sms, stream);
	}
	return LaunchKN<K, 256>(p, num_sms, stream);
}

} // namespace

bool PolyFastSupported(size_t n, size_t K, size_t data_stride, size_t mask_stride, void const *data,
		void const *mask, void const *best_fit, void const *residual, void const *final_mask) {
	if (K < 1 || K > static_cast<size_t>(kMaxPolyK)) {
		return false;
	}
	if (n < 16 || n % 4 != 0 || n > 32 * 256 || n < K) {
		return false;
	}
	if (data_stride % 4 != 0 || mask_stride % 4 != 0) {
		return false;
	}
	auto a16 = [](void const *q) {return q == nullptr || (reinterpret_cast<uintptr_t>(q) % 16) == 0;};
	auto a4 = [](void const *q) {return (reinterpret_cast<uintptr_t>(q) % 4) == 0;};
	return a16(data) && a16(best_fit) && a16(residual) && a4(mask) && a4(final_mask);
}

bool PolyFusedCalibrationSupported(size_t n) {
	return n == 4u * 64u * kSlotGroups || n == 4u * 128u * kSlotGroups || n == 4u * 256u * kSlotGroups;
}

bool PolyPrefersCalibrationPrePass(size_t n, int num_fitting_max) {
	char const *fused = getenv("SAKURA_B200_CAL_FUSED");
	if (fused != nullptr && fused[0] == '1') {
		return false;
	}
	char const *k = getenv("SAKURA_B200_POLY_KERNEL");
	if (k != nullptr && (k[0] == 'c' || k[0] == 's')) {
		return false; // a forced kernel keeps its fused variant
	}
	if (getenv("SAKURA_B200_POLY_SCREEN") != nullptr) {
		return false;
	}
	return (n == 2048 || n == 4096) && num_fitting_max >= 2;
}

namespace {

cudaError_t LaunchPolyFitKernel(PolyFitParams const &p, int num_sms, cudaStream_t stream);

enum PolyKernelChoice : int {
	kChoiceChannel = 0, kChoiceScreen = 1, kChoiceWarp = 2
};

PolyKernelChoice ChooseKernel(PolyFitParams const &p) {
	// SAKURA_B200_POLY_KERNEL = channel | screen | warp forces one kernel (where the shape allows it);
	// SAKURA_B200_POLY_SCREEN = 0 / 1 is the older switch between the first two. Read per call so that a
	// test can switch within one process (the parity tests compare the kernels with each other).
	// Default: the warp-per-spectrum kernel from the second fit on (4096 / 2048 channels), the screening
	// kernel from the third fit on for 8192 channels, the per-channel kernel otherwise (measured on B200,
	// 4096 channels, order 5: see profiles/README.md).
	bool const can_screen = p.fallback_rows != nullptr && PolyScreenSupported(p);
	bool const can_warp = p.fallback_rows != nullptr && PolyWarpSupported(p);
	char const *k = getenv("SAKURA_B200_POLY_KERNEL");
	if (k != nullptr) {
		if (k[0] == 'c') {
			return kChoiceChannel;
		}
		if (k[0] == 's') {
			return can_screen ? kChoiceScreen : kChoiceChannel;
		}
		if (k[0] == 'w') {
			return can_warp ? kChoiceWarp : kChoiceChannel;
		}
	}
	char const *e = getenv("SAKURA_B200_POLY_SCREEN");
	if (e != nullptr && (e[0] == '0' || e[0] == '1')) {
		return (e[0] == '1' && can_screen) ? kChoiceScreen : kChoiceChannel;
	}
	if (can_warp && p.num_fitting_max >= 2) {
		return kChoiceWarp;
	}
	if (can_screen && p.num_fitting_max >= 3) {
		return kChoiceScreen;
	}
	return kChoiceChannel;
}

} // namespace

cudaError_t LaunchPolyFit(PolyFitParams const &p_in, int num_sms, cudaStream_t stream, int *launches,
		char const **kernel_name) {
	PolyFitParams p = p_in;
	p.row_index = nullptr;
	p.row_count = nullptr;
	PolyKernelChoice const choice = ChooseKernel(p);
	if (kernel_name != nullptr) {
		*kernel_name = (choice == kChoiceWarp) ? "poly_warp" : (choice == kChoiceScreen) ? "poly_screen" : "poly_fast";
	}
	if (choice != kChoiceChannel) {
		// screened clip iterations first; the rows that kernel gives up (undecidable channel,
		// degenerate system) are then fitted by the per-channel kernel from its list
		p.fallback_count = p.fallback_rows;
		p.fallback_rows = p.fallback_rows + 1;
		cudaError_t err = cudaMemsetAsync(p.fallback_count, 0, sizeof(int32_t), stream);
		if (err != cudaSuccess) {
			return err;
		}
		err = (choice == kChoiceWarp) ? LaunchPolyWarpFit(p, num_sms, stream) :
				LaunchPolyScreenFit(p, num_sms, stream);
		*launches += 1;
		if (err != cudaSuccess) {
			return err;
		}
		p.row_index = p.fallback_rows;
		p.row_count = p.fallback_count;
	}
	*launches += 1;
	return LaunchPolyFitKernel(p, num_sms, stream);
}

namespace {

cudaError_t LaunchPolyFitKernel(PolyFitParams const &p, int num_sms, cudaStream_t stream) {
	switch (p.K) {
	case 1:
		return LaunchK<1>(p, num_sms, stream);
	case 2:
		return LaunchK<2>(p, num_sms, stream);
	case 3:
		return LaunchK<3>(p, num_sms, stream);
	case 4:
		return LaunchK<4>(p, num_sms, stream);
	case 5:
		return LaunchK<5>(p, num_sms, stream);
	case 6:
		return LaunchK<6>(p, num_sms, stream);
	case 7:
		return LaunchK<7>(p, num_sms, stream);
	case 8:
		return LaunchK<8>(p, num_sms, stream);
	default:
		return cudaErrorInvalidValue;
	}
}

} // namespace

} // namespace sakura_b200
