// Warp-level solve of a small symmetric positive definite system in shared memory, used by the
// sinusoid kernel (the cubic-spline kernel has its own banded solve since it fits in the B-spline
// basis): blocked LDL^T (8 x 8 blocks) whose trailing updates are FP64
// tensor-core MMAs (mma.sync m8n8k4). The reference solves the same normal equations with Eigen's
// full-pivot LU (numeric_operation.cc:609-622); rows whose pivots are not safely positive are
// reported to the caller, which hands them to the kernel that restates that LU.
#ifndef SAKURA_B200_LDLT_DEVICE_CUH_
#define SAKURA_B200_LDLT_DEVICE_CUH_

#include <float.h>

namespace sakura_b200 {

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

__host__ __device__ constexpr int LdltLd(int Kp) {
	return Kp + 2; // leading dimension of the bordered matrix
}
__host__ __device__ constexpr int LdltDoubles(int Kp) {
	return (Kp + 1) * LdltLd(Kp);
}

// A is the (Kp+1) x LdltLd(Kp) bordered matrix [[G, .], [b^T, .]], Kp = 8 NT8: lower triangle of G
// in rows 0..Kp-1 (rows/columns K..Kp-1 padded with the identity), right-hand side in row Kp.
// After the factorisation the border row holds D^-1 L^-1 b; then L^T x = z gives x[0..Kp).
// Returns false (warp-uniform) when a pivot is not safely positive. All 32 lanes must call.
// PACKED: row i of the lower triangle (and the border row) starts at i (i + 1) / 2 instead of
// i * LdltLd(Kp): half the shared memory, so that twice as many warps can solve at the same time.
// (Reads of a diagonal block's upper triangle then run a few entries into the next row: those values
// never reach a result, as with the stale upper triangle of the unpacked layout.)
__host__ __device__ constexpr int LdltPackedDoubles(int Kp) {
	return (Kp + 1) * (Kp + 2) / 2;
}

template<int NT8, bool PACKED = false>
__device__ bool WarpLdltSolve(double *A, int K, double *xout, int lane) {
	constexpr int Kp = 8 * NT8;
	auto Row = [](int i) {
		return PACKED ? (i * (i + 1)) / 2 : i * LdltLd(8 * NT8);
	};
	constexpr unsigned kFull = 0xffffffffu;
	int const r8 = lane >> 2;
	int const c4 = lane & 3;
	int const l8 = lane & 7;
	double maxd = 0.0;
	double const tol = DBL_EPSILON * K;
	bool ok = true;
#pragma unroll 1
	for (int kb = 0; kb < NT8; ++kb) {
		int const c0 = 8 * kb;
		// diagonal block: lane l8 owns row l8, columns in registers, pivots by shuffle
		double row[8], dinv[8], dneg[8];
#pragma unroll
		for (int j = 0; j < 8; ++j) {
			row[j] = A[Row(c0 + l8) + c0 + j];
		}
#pragma unroll
		for (int k = 0; k < 8; ++k) {
			double const dk = __shfl_sync(kFull, row[k], k);
			maxd = fmax(maxd, fabs(dk));
			ok = ok && (dk > maxd * tol);
			double const rk = fast_rcp_pos(dk);
			dinv[k] = rk;
			dneg[k] = -dk;
			double const lk = row[k] * rk;
#pragma unroll
			for (int j = k + 1; j < 8; ++j) {
				double const ajk = __shfl_sync(kFull, row[k], j);
				row[j] = fma(-lk, ajk, row[j]);
			}
			row[k] = (l8 > k) ? lk : row[k];
		}
		if (lane < 8) {
#pragma unroll
			for (int j = 0; j < 7; ++j) {
				if (j < lane) {
					A[Row(c0 + lane) + c0 + j] = row[j];
				}
			}
		}
		__syncwarp();
		// panel below the block, border row included: w = a - sum_t w_t L_jt, L_ij = w_j / d_j
		for (int i = c0 + 8 + lane; i <= Kp; i += 32) {
			double w[8];
#pragma unroll
			for (int j = 0; j < 8; ++j) {
				w[j] = A[Row(i) + c0 + j];
			}
#pragma unroll
			for (int j = 1; j < 8; ++j) {
#pragma unroll
				for (int t = 0; t < j; ++t) {
					w[j] = fma(-w[t], A[Row(c0 + j) + c0 + t], w[j]);
				}
			}
#pragma unroll
			for (int j = 0; j < 8; ++j) {
				A[Row(i) + c0 + j] = w[j] * dinv[j];
			}
		}
		__syncwarp();
		// trailing blocks (I, J), kb < J <= I: C -= (L_I D) L_J^T
		double const nd0 = c4 == 0 ? dneg[0] : (c4 == 1 ? dneg[1] : (c4 == 2 ? dneg[2] : dneg[3]));
		double const nd1 = c4 == 0 ? dneg[4] : (c4 == 1 ? dneg[5] : (c4 == 2 ? dneg[6] : dneg[7]));
		for (int I = kb + 1; I < NT8; ++I) {
			double const *ai = A + Row(8 * I + r8) + c0 + c4;
			double const a0 = ai[0] * nd0;
			double const a1 = ai[4] * nd1;
			for (int J = kb + 1; J <= I; ++J) {
				double const *bj = A + Row(8 * J + r8) + c0 + c4;
				double *cp = A + Row(8 * I + r8) + 8 * J + 2 * c4;
				double x0 = cp[0];
				double x1 = cp[1];
				dmma(x0, x1, a0, bj[0]);
				dmma(x0, x1, a1, bj[4]);
				if (PACKED && J == I) {
					// the diagonal block's upper triangle does not exist in the packed layout
					if (2 * c4 <= r8) {
						cp[0] = x0;
					}
					if (2 * c4 + 1 <= r8) {
						cp[1] = x1;
					}
				} else {
					cp[0] = x0;
					cp[1] = x1;
				}
			}
		}
		// border row
		for (int j = c0 + 8 + lane; j < Kp; j += 32) {
			double acc = A[Row(Kp) + j];
#pragma unroll
			for (int t = 0; t < 8; ++t) {
				acc = fma(A[Row(Kp) + c0 + t] * dneg[t], A[Row(j) + c0 + t], acc);
			}
			A[Row(Kp) + j] = acc;
		}
		__syncwarp();
	}
	// L^T x = z; lane holds x[lane] and x[lane + 32]
	double xa = (lane < Kp) ? A[Row(Kp) + lane] : 0.0;
	double xb = (Kp > 32 && lane + 32 < Kp) ? A[Row(Kp) + lane + 32] : 0.0;
	for (int k = Kp - 1; k >= 1; --k) {
		double const xk = (k >= 32) ? __shfl_sync(kFull, xb, k - 32) : __shfl_sync(kFull, xa, k);
		double const *Lk = A + Row(k);
		if (lane < k) {
			xa = fma(-Lk[lane], xk, xa);
		}
		if (Kp > 32 && lane + 32 < k) {
			xb = fma(-Lk[lane + 32], xk, xb);
		}
	}
	if (lane < Kp) {
		xout[lane] = xa;
	}
	if (Kp > 32 && lane + 32 < Kp) {
		xout[lane + 32] = xb;
	}
	__syncwarp();
	return ok;
}

} // namespace sakura_b200

#endif
