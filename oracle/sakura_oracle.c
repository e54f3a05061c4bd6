/*
 * TEST INFRASTRUCTURE ONLY (oracle, kind "port"). Never linked into, imported or called by the
 * shipped sakura_b200 library.
 *
 * Plain-C, scalar restatement of the reference's masked least-squares baseline fit with
 * iterative sigma clipping and of the statistics it depends on. Every function cites the
 * reference code it follows (paths relative to tnakazato/sakura, libsakura/src/).
 * Floating-point operations are written in the order the reference's Release build
 * (-O3 -mavx2 -mfma, GCC default -ffp-contract=fast) executes them; where that build fuses a
 * multiply-add, fma() is written explicitly and this file is compiled with -ffp-contract=off.
 *
 * PARITY PINNING: this restatement is checked bit-for-bit (masks, rms, residuals, sums) against
 * oracle/_ref/libsakura_ref.so -- the reference's own object code -- in
 * tests/test_oracle_port_vs_ref.py, and both are checked against the reference's known-answer
 * tests in tests/test_oracle_known_answers.py. The K x K solve restates Eigen 3.3/3.4
 * FullPivLU (Eigen is a third-party dependency absent from /root/reference; see
 * oracle/eigen_shim/Eigen/Core); at bit level that part is "parity unpinned".
 */
#include "sakura_oracle.h"

#include <float.h>
#include <math.h>
#include <stdlib.h>
#include <string.h>

#if defined(_OPENMP)
#include <omp.h>
#endif

enum {
	kPolynomial = 0, kChebyshev = 1, kCubicSpline = 2, kSinusoid = 3
};
enum {
	kStatusOK = 0, kStatusNG = 1, kStatusInvalidArgument = 2
};
enum {
	kLsqOK = 0, kLsqNG = 1, kLsqNotEnoughData = 2
};

/* ------------------------------------------------------------------ basis tables */

/* baseline.cc:251-289 */
static void fill_polynomial(size_t n, size_t nb, double *t) {
	double const max_x = (double) (n - 1);
	size_t idx = 0;
	for (size_t i = 0; i < n; ++i) {
		double const x = (double) i / max_x;
		double val = 1.0;
		for (size_t j = 0; j < nb; ++j) {
			t[idx++] = val;
			val *= x;
		}
	}
}

/* baseline.cc:299-327; the recurrence compiles to vfmsub231sd */
static void fill_chebyshev(size_t n, size_t nb, double *t) {
	size_t idx = 0;
	if (nb == 1) {
		for (size_t i = 0; i < n; ++i) {
			t[idx++] = 1.0;
		}
		return;
	}
	double const max_x = (double) (n - 1);
	for (size_t i = 0; i < n; ++i) {
		t[idx++] = 1.0;
		double const x = 2.0 * (double) i / max_x - 1.0;
		t[idx++] = x;
		double const two_x = 2.0 * x;
		for (size_t j = 2; j < nb; ++j) {
			t[idx] = fma(two_x, t[idx - 1], -t[idx - 2]);
			++idx;
		}
	}
}

/* baseline.cc:343-366 */
static void fill_sinusoid(size_t n, size_t max_nwave, double *t) {
	size_t idx = 0;
	if (n == 1) {
		t[0] = 1.0;
		return;
	}
	double const max_x = (double) (n - 1);
	double const norm = 2.0 * M_PI / max_x;
	for (size_t i = 0; i < n; ++i) {
		t[idx++] = 1.0;
		for (size_t w = 1; w <= max_nwave; ++w) {
			double const x = (double) i * (double) w * norm;
			t[idx++] = sin(x);
			t[idx++] = cos(x);
		}
	}
}

/* ------------------------------------------------------------------ normal equations */

/* numeric_operation.cc:192-225 + 370-396 (templates) / :238-270, 409-435 (generic):
 * sequential over channels, every accumulation an fma */
static int get_lsq_coefficients(size_t n, float const *data, unsigned char const *mask,
		size_t nb_model, double const *basis, size_t K, size_t const *use_idx, double *G,
		double *b) {
	for (size_t i = 0; i < K * K; ++i) {
		G[i] = 0.0;
	}
	size_t num_unmasked = 0;
	for (size_t i = 0; i < n; ++i) {
		if (mask[i]) {
			double const *m = &basis[i * nb_model];
			for (size_t j = 0; j < K; ++j) {
				double const k = m[use_idx[j]];
				double *row = &G[j * K];
				for (size_t l = 0; l < K; ++l) {
					row[l] = fma(k, m[use_idx[l]], row[l]);
				}
			}
			++num_unmasked;
		}
	}
	if (num_unmasked < ((K <= 100) ? K : nb_model)) {
		return kStatusNG; /* :221-224 / :266-269 */
	}
	for (size_t i = 0; i < K; ++i) {
		b[i] = 0.0;
	}
	for (size_t i = 0; i < n; ++i) {
		if (mask[i]) {
			double const *m = &basis[i * nb_model];
			double const d = (double) data[i];
			for (size_t l = 0; l < K; ++l) {
				b[l] = fma(d, m[use_idx[l]], b[l]);
			}
		}
	}
	return kStatusOK;
}

/* numeric_operation.cc:283-314 + 448-479: subtract the clipped channels (fnmadd) */
static void update_lsq_coefficients(float const *data, unsigned char const *mask,
		size_t num_clipped, size_t const *clipped, size_t nb_model, double const *basis, size_t K,
		size_t const *use_idx, double *G, double *b) {
	for (size_t c = 0; c < num_clipped; ++c) {
		if (!mask[clipped[c]]) {
			continue;
		}
		double const *m = &basis[clipped[c] * nb_model];
		for (size_t j = 0; j < K; ++j) {
			double const k = m[use_idx[j]];
			double *row = &G[j * K];
			for (size_t l = 0; l < K; ++l) {
				row[l] = fma(-k, m[use_idx[l]], row[l]);
			}
		}
	}
	for (size_t c = 0; c < num_clipped; ++c) {
		if (!mask[clipped[c]]) {
			continue;
		}
		size_t const ci = clipped[c];
		double const *m = &basis[ci * nb_model];
		double const d = (double) data[ci];
		for (size_t l = 0; l < K; ++l) {
			b[l] = fma(-d, m[use_idx[l]], b[l]);
		}
	}
}

/* numeric_operation.cc:609-622 -> Eigen FullPivLU::computeInPlace / rank / _solve_impl.
 * lu is column-major scratch of n*n doubles; tr holds 2n indices. */
static void solve_full_piv_lu(size_t n, double const *A, double const *rhs, double *x,
		double *lu, double *c, ptrdiff_t *tr) {
	ptrdiff_t const N = (ptrdiff_t) n;
	ptrdiff_t *row_tr = tr, *col_tr = tr + N;
#define AT(i, j) lu[(j) * N + (i)]
	for (ptrdiff_t j = 0; j < N; ++j) {
		for (ptrdiff_t i = 0; i < N; ++i) {
			AT(i, j) = A[i * N + j];
		}
	}
	ptrdiff_t nonzero_pivots = N;
	double maxpivot = 0.0;
	for (ptrdiff_t k = 0; k < N; ++k) {
		ptrdiff_t prow = k, pcol = k;
		double biggest = -1.0;
		for (ptrdiff_t j = k; j < N; ++j) {
			for (ptrdiff_t i = k; i < N; ++i) {
				double const v = fabs(AT(i, j));
				if (v > biggest) {
					biggest = v;
					prow = i;
					pcol = j;
				}
			}
		}
		if (biggest == 0.0) {
			nonzero_pivots = k;
			for (ptrdiff_t i = k; i < N; ++i) {
				row_tr[i] = i;
				col_tr[i] = i;
			}
			break;
		}
		if (biggest > maxpivot) {
			maxpivot = biggest;
		}
		row_tr[k] = prow;
		col_tr[k] = pcol;
		if (prow != k) {
			for (ptrdiff_t j = 0; j < N; ++j) {
				double const t = AT(k, j);
				AT(k, j) = AT(prow, j);
				AT(prow, j) = t;
			}
		}
		if (pcol != k) {
			for (ptrdiff_t i = 0; i < N; ++i) {
				double const t = AT(i, k);
				AT(i, k) = AT(i, pcol);
				AT(i, pcol) = t;
			}
		}
		if (k < N - 1) {
			double const pivot = AT(k, k);
			for (ptrdiff_t i = k + 1; i < N; ++i) {
				AT(i, k) /= pivot;
			}
			for (ptrdiff_t j = k + 1; j < N; ++j) {
				double const ukj = AT(k, j);
				for (ptrdiff_t i = k + 1; i < N; ++i) {
					AT(i, j) = fma(-AT(i, k), ukj, AT(i, j));
				}
			}
		}
	}
	double const threshold = fabs(maxpivot) * (DBL_EPSILON * (double) N);
	ptrdiff_t rank = 0;
	for (ptrdiff_t i = 0; i < nonzero_pivots; ++i) {
		rank += (fabs(AT(i, i)) > threshold) ? 1 : 0;
	}
	for (ptrdiff_t i = 0; i < N; ++i) {
		x[i] = 0.0;
		c[i] = rhs[i];
	}
	if (rank == 0) {
		return;
	}
	for (ptrdiff_t k = 0; k < N; ++k) {
		ptrdiff_t const r = row_tr[k];
		if (r != k) {
			double const t = c[k];
			c[k] = c[r];
			c[r] = t;
		}
	}
	for (ptrdiff_t j = 0; j < N; ++j) {
		double const cj = c[j];
		for (ptrdiff_t i = j + 1; i < N; ++i) {
			c[i] = fma(-AT(i, j), cj, c[i]);
		}
	}
	for (ptrdiff_t j = rank - 1; j >= 0; --j) {
		c[j] /= AT(j, j);
		double const cj = c[j];
		for (ptrdiff_t i = 0; i < j; ++i) {
			c[i] = fma(-AT(i, j), cj, c[i]);
		}
	}
	/* x = Q c: reuse row_tr as the permutation q */
	for (ptrdiff_t i = 0; i < N; ++i) {
		row_tr[i] = i;
	}
	for (ptrdiff_t k = 0; k < N; ++k) {
		ptrdiff_t const t = row_tr[k];
		row_tr[k] = row_tr[col_tr[k]];
		row_tr[col_tr[k]] = t;
	}
	for (ptrdiff_t i = 0; i < rank; ++i) {
		x[row_tr[i]] = c[i];
	}
#undef AT
}

/* ------------------------------------------------------------------ accurate statistics */

typedef struct {
	int count[8];
	double sum[4];
	double sq[4];
	float mn[8], mx[8];
	int imn[8], imx[8];
} SimdAcc; /* statistics.cc:445-464 MiddleLevelAccumulator */

static void acc_clear(SimdAcc *a) {
	for (int l = 0; l < 8; ++l) {
		a->count[l] = 0;
		a->mn[l] = NAN;
		a->mx[l] = NAN;
		a->imn[l] = -1;
		a->imx[l] = -1;
	}
	for (int l = 0; l < 4; ++l) {
		a->sum[l] = 0.0;
		a->sq[l] = 0.0;
	}
}

/* statistics.cc:338-402 StatsBlock: 8 consecutive elements starting at `pos`; `index` is the
 * value recorded for min/max positions (pos for the accurate traverse, block number for the
 * fast variant). Lane l of the double accumulators takes element l, then element l+4. */
static void stats_block(size_t pos, int index, float const *data, unsigned char const *mask,
		SimdAcc *a) {
	float value[8];
	for (int l = 0; l < 8; ++l) {
		int const valid = mask[pos + l] != 0;
		a->count[l] += (int) (signed char) mask[pos + l]; /* _mm256_cvtepi8_epi32 */
		float const v = valid ? data[pos + l] : NAN;
		int const idx = valid ? index : -1;
		if ((v < a->mn[l]) || a->imn[l] < 0) {
			a->mn[l] = v;
			a->imn[l] = idx;
		}
		if ((v > a->mx[l]) || a->imx[l] < 0) {
			a->mx[l] = v;
			a->imx[l] = idx;
		}
		value[l] = valid ? v : 0.0f;
	}
	for (int l = 0; l < 4; ++l) {
		double const v = (double) value[l];
		a->sum[l] = a->sum[l] + v;
		a->sq[l] = fma(v, v, a->sq[l]);
	}
	for (int l = 0; l < 4; ++l) {
		double const v = (double) value[4 + l];
		a->sum[l] = a->sum[l] + v;
		a->sq[l] = fma(v, v, a->sq[l]);
	}
}

/* statistics.cc:529-561 MiddleLevelReduce */
static void acc_merge(SimdAcc const *inc, SimdAcc *a) {
	for (int l = 0; l < 8; ++l) {
		a->count[l] += inc->count[l];
		if ((inc->mn[l] < a->mn[l]) || a->imn[l] < 0) {
			a->mn[l] = inc->mn[l];
			a->imn[l] = inc->imn[l];
		}
		if ((inc->mx[l] > a->mx[l]) || a->imx[l] < 0) {
			a->mx[l] = inc->mx[l];
			a->imx[l] = inc->imx[l];
		}
	}
	for (int l = 0; l < 4; ++l) {
		a->sum[l] += inc->sum[l];
		a->sq[l] += inc->sq[l];
	}
}

/* statistics.cc:574-592 SequentialReduce (+ :232-254 Accumulate); note the recorded index is
 * pos - j, i.e. the start of the tail (a quirk of the reference kept here on purpose) */
static void sequential_reduce(size_t offset, size_t elements, float const *data,
		unsigned char const *mask, SimdAcc *a) {
	for (size_t j = 0; j < elements; ++j) {
		size_t const pos = offset + j;
		if (mask[pos]) {
			int const lf = (int) (j % 8);
			int const ld = (int) (j % 4);
			float const value = data[pos];
			a->count[lf] += 1;
			double const v = (double) value;
			a->sum[ld] += v;
			a->sq[ld] = fma(v, v, a->sq[ld]);
			int const idx = (int) (pos - j);
			if (a->imn[lf] < 0) {
				a->imn[lf] = a->imx[lf] = idx;
				a->mn[lf] = a->mx[lf] = value;
			} else {
				if (value < a->mn[lf]) {
					a->imn[lf] = idx;
					a->mn[lf] = value;
				}
				if (value > a->mx[lf]) {
					a->imx[lf] = idx;
					a->mx[lf] = value;
				}
			}
		}
	}
}

/* statistics.cc:404-409 (hadd then low+high) and :471-527 TopLevelReduce */
static void top_level(SimdAcc const *a, int index_scale, OracleStatistics *r) {
	unsigned int cnt = 0;
	for (int l = 0; l < 8; ++l) {
		cnt += (unsigned int) a->count[l];
	}
	r->count = (size_t) cnt;
	r->sum = (a->sum[0] + a->sum[1]) + (a->sum[2] + a->sum[3]);
	r->square_sum = (a->sq[0] + a->sq[1]) + (a->sq[2] + a->sq[3]);
	{
		float m = a->mn[0];
		long idx = (index_scale == 1) ? (long) a->imn[0] + 0 :
				(a->imn[0] == -1 ? -1 : (long) a->imn[0] * index_scale);
		for (int l = 1; l < 8; ++l) {
			if (!isnan(a->mn[l])) {
				if (isnan(m) || a->mn[l] < m) {
					m = a->mn[l];
					idx = (long) a->imn[l] * index_scale + l;
				}
			}
		}
		r->min = m;
		r->index_of_min = idx < -1 ? -1 : idx;
	}
	{
		float m = a->mx[0];
		long idx = (index_scale == 1) ? (long) a->imx[0] + 0 :
				(a->imx[0] == -1 ? -1 : (long) a->imx[0] * index_scale);
		for (int l = 1; l < 8; ++l) {
			if (!isnan(a->mx[l])) {
				if (isnan(m) || a->mx[l] > m) {
					m = a->mx[l];
					idx = (long) a->imx[l] * index_scale + l;
				}
			}
		}
		r->max = m;
		r->index_of_max = idx < -1 ? -1 : idx;
	}
}

/* statistics.cc:126-167 Traverse<4,8> + :37-100 BlockWiseTraverse, flattened: the blocks are
 * visited leaf by leaf (leaf = 8-element blocks start, start+stride, start+2*stride, ...),
 * leaves are combined four at a time, depth first. */
static void leaf_reduce(size_t start, size_t stride, size_t full_blocks, size_t upper_limit,
		float const *data, unsigned char const *mask, SimdAcc const *residual, SimdAcc *acc) {
	acc_clear(acc);
	if (start == 0) {
		acc_merge(residual, acc);
	}
	size_t delta = start;
	for (size_t i = 0; i < full_blocks; ++i) {
		for (size_t j = 0; j < 4; ++j) {
			stats_block(delta + j * stride, (int) (delta + j * stride), data, mask, acc);
		}
		delta += stride * 4;
	}
	for (size_t j = 0; j < 4; ++j) {
		size_t const index = delta + j * stride;
		if (index >= upper_limit) {
			break;
		}
		stats_block(index, (int) index, data, mask, acc);
	}
}

static void tree_reduce(int depth, size_t *start, size_t stride, size_t full_blocks,
		size_t upper_limit, float const *data, unsigned char const *mask,
		SimdAcc const *residual, SimdAcc *out) {
	if (depth == 0) {
		leaf_reduce(*start, stride, full_blocks, upper_limit, data, mask, residual, out);
		*start += 8;
		return;
	}
	SimdAcc child[4];
	for (int j = 0; j < 4; ++j) {
		tree_reduce(depth - 1, start, stride, full_blocks, upper_limit, data, mask, residual,
				&child[j]);
	}
	*out = child[0];
	for (int j = 1; j < 4; ++j) {
		acc_merge(&child[j], out);
	}
}

void oracle_compute_accurate_statistics(size_t num_data, float const *data,
		unsigned char const *mask, OracleStatistics *result) {
	size_t num_blocks = num_data / 8;
	int levels = 0;
	size_t n = 1;
	size_t stride = 8;
	while (n * 4 <= num_blocks) {
		++levels;
		n *= 4;
		stride *= 4;
	}
	size_t const full_blocks = num_blocks / n;
	SimdAcc residual;
	acc_clear(&residual);
	if (levels == 0) {
		num_blocks = 0;
	}
	sequential_reduce(num_blocks * 8, num_data - num_blocks * 8, data, mask, &residual);
	if (levels > 0) {
		stride /= 4;
		SimdAcc total;
		size_t start = 0;
		tree_reduce(levels - 1, &start, stride, full_blocks, num_blocks * 8, data, mask, &residual,
				&total);
		top_level(&total, 1, result);
	} else {
		top_level(&residual, 1, result);
	}
}

/* statistics.cc:595-700 ComputeStatisticsSimdFloat: linear over 8-element blocks, scalar tail */
void oracle_compute_statistics(size_t num_data, float const *data, unsigned char const *mask,
		OracleStatistics *result) {
	SimdAcc a;
	acc_clear(&a);
	size_t const nblk = num_data / 8;
	for (size_t i = 0; i < nblk; ++i) {
		stats_block(i * 8, (int) i, data, mask, &a);
	}
	top_level(&a, 8, result);
	size_t counted = result->count;
	double total = result->sum;
	double square_total = result->square_sum;
	for (size_t i = nblk * 8; i < num_data; ++i) {
		if (mask[i]) {
			float const f = data[i];
			double const d = (double) f;
			++counted;
			total += d;
			square_total = fma(d, d, square_total);
			if (!isnan(f)) {
				if (isnan(result->min) || f < result->min) {
					result->min = f;
					result->index_of_min = (long) i;
				}
				if (isnan(result->max) || f > result->max) {
					result->max = f;
					result->index_of_max = (long) i;
				}
			}
		}
	}
	result->count = counted;
	result->sum = total;
	result->square_sum = square_total;
}

/* ------------------------------------------------------------------ the fit */

typedef struct {
	int type;
	size_t param; /* order / npiece / nwave of the context */
	size_t n;
	size_t num_bases; /* table columns */
	size_t lsq_max;
	double *basis; /* [n][num_bases] */
	/* scratch (baseline.cc:374-488) */
	double *G, *b, *coeff, *coeff_full, *lu, *cvec, *cspline_basis;
	ptrdiff_t *tr;
	size_t *clipped, *use_idx, *boundary;
	float *best_fit, *residual;
} Ctx;

static void ctx_destroy(Ctx *c) {
	if (c == NULL) {
		return;
	}
	free(c->basis);
	free(c->G);
	free(c->b);
	free(c->coeff);
	free(c->coeff_full);
	free(c->lu);
	free(c->cvec);
	free(c->cspline_basis);
	free(c->tr);
	free(c->clipped);
	free(c->use_idx);
	free(c->boundary);
	free(c->best_fit);
	free(c->residual);
	free(c);
}

/* baseline.cc:584-621, 497-525, 111-182 */
static Ctx *ctx_create(int type, size_t param, size_t n) {
	Ctx *c = (Ctx *) calloc(1, sizeof(Ctx));
	if (c == NULL) {
		return NULL;
	}
	c->type = type;
	c->param = param;
	c->n = n;
	switch (type) {
	case kPolynomial:
	case kChebyshev:
		c->num_bases = param + 1;
		c->lsq_max = param + 1;
		break;
	case kCubicSpline:
		c->num_bases = 4;
		c->lsq_max = param + 3;
		break;
	default:
		c->num_bases = 2 * param + 1;
		c->lsq_max = 2 * param + 1;
		break;
	}
	size_t const K = c->lsq_max;
	size_t const nfull = (type == kCubicSpline) ? 4 * param : K;
	c->basis = (double *) malloc(sizeof(double) * c->num_bases * n);
	c->G = (double *) malloc(sizeof(double) * K * K);
	c->b = (double *) malloc(sizeof(double) * K);
	c->coeff = (double *) malloc(sizeof(double) * K);
	c->coeff_full = (double *) malloc(sizeof(double) * (nfull > 0 ? nfull : 1));
	c->lu = (double *) malloc(sizeof(double) * K * K);
	c->cvec = (double *) malloc(sizeof(double) * K);
	c->tr = (ptrdiff_t *) malloc(sizeof(ptrdiff_t) * 2 * K);
	c->clipped = (size_t *) malloc(sizeof(size_t) * n);
	c->use_idx = (size_t *) malloc(sizeof(size_t) * K);
	c->boundary = (size_t *) malloc(sizeof(size_t) * (param + 2));
	c->best_fit = (float *) malloc(sizeof(float) * n);
	c->residual = (float *) malloc(sizeof(float) * n);
	if (type == kCubicSpline) {
		c->cspline_basis = (double *) malloc(sizeof(double) * K * n);
	}
	if (!c->basis || !c->G || !c->b || !c->coeff || !c->coeff_full || !c->lu || !c->cvec || !c->tr
			|| !c->clipped || !c->use_idx || !c->boundary || !c->best_fit || !c->residual
			|| (type == kCubicSpline && !c->cspline_basis)) {
		ctx_destroy(c);
		return NULL;
	}
	for (size_t i = 0; i < K; ++i) {
		c->use_idx[i] = i; /* baseline.cc:439-443 */
	}
	switch (type) {
	case kPolynomial:
	case kCubicSpline:
		fill_polynomial(n, c->num_bases, c->basis);
		break;
	case kChebyshev:
		fill_chebyshev(n, c->num_bases, c->basis);
		break;
	default:
		fill_sinusoid(n, param, c->basis);
		break;
	}
	return c;
}

/* baseline.cc:808-844. Returns 0, or -1 for "Too few unmasked data" (:822-824). */
static int get_boundaries(size_t n, unsigned char const *mask, size_t num_boundary,
		size_t *boundary) {
	size_t num_unmasked = 0;
	for (size_t i = 0; i < n; ++i) {
		if (mask[i]) {
			++num_unmasked;
		}
	}
	size_t const last = num_boundary - 1;
	if (num_unmasked < last) {
		return -1;
	}
	boundary[0] = 0;
	size_t idx = 1;
	size_t count = 0;
	for (size_t i = 0; i < n; ++i) {
		if (last <= idx) {
			break;
		}
		if (mask[i]) {
			if (num_unmasked * idx <= count * last) {
				boundary[idx] = i;
				++idx;
			}
			++count;
		}
	}
	boundary[last] = n;
	return 0;
}

static double nonneg(double v) {
	return (0.0 < v) ? v : 0.0; /* std::max(0.0, v), baseline.cc:924 */
}

/* baseline.cc:952-980 + 911-940 */
static void set_spline_basis(size_t n, size_t num_boundary, size_t stride, size_t const *boundary,
		double *out) {
	double const max_x = (double) (n - 1);
	for (size_t i = 0; i < n; ++i) {
		double *row = &out[i * stride];
		double const i_d = (double) i;
		double const x = i_d / max_x;
		double val = 1.0;
		for (size_t j = 0; j < 4; ++j) {
			row[j] = val;
			val *= x;
		}
		if (num_boundary > 2) {
			double const v = (i_d - (double) boundary[1]) / max_x;
			double cube_prev = v * v * v;
			row[3] -= nonneg(cube_prev);
			for (size_t j = 2; j < num_boundary; ++j) {
				double const w = (i_d - (double) boundary[j]) / max_x;
				double const cube_cur = w * w * w;
				row[2 + j] = nonneg(cube_prev - nonneg(cube_cur));
				cube_prev = cube_cur;
			}
		}
	}
}

/* baseline.cc:868-893 as compiled (fnmadd / fmadd) */
static void full_spline_coefficients(size_t num_boundary, size_t const *boundary,
		double const *raw, double *cf) {
	for (size_t i = 0; i < 4; ++i) {
		cf[i] = raw[i];
	}
	double const max_x = (double) (boundary[num_boundary - 1] - 1);
	for (size_t i = 1; i < num_boundary - 1; ++i) {
		size_t const j = i + 3;
		double const c = raw[j] - cf[4 * (i - 1) + 3];
		double const b = (double) boundary[i] / max_x;
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

/* baseline.cc:661-715 (fma chain from 0, ascending j, round to float) + :629-642 */
static void model_and_residual(Ctx const *c, size_t K, float const *data, double const *coeff,
		float *best_fit, float *residual) {
	for (size_t i = 0; i < c->n; ++i) {
		double const *row = &c->basis[i * c->num_bases];
		double m = 0.0;
		for (size_t j = 0; j < K; ++j) {
			m = fma(coeff[j], row[c->use_idx[j]], m);
		}
		best_fit[i] = (float) m;
	}
	for (size_t i = 0; i < c->n; ++i) {
		residual[i] = data[i] - best_fit[i];
	}
}

/* baseline.cc:724-758: four products, then left-to-right adds, no fma */
static void spline_model_and_residual(Ctx const *c, size_t num_boundary, size_t const *boundary,
		float const *data, double const *cf, float *best_fit, float *residual) {
	for (size_t p = 0; p + 1 < num_boundary; ++p) {
		for (size_t j = boundary[p]; j < boundary[p + 1]; ++j) {
			double const *bs = &c->basis[4 * j];
			double const p0 = cf[4 * p + 0] * bs[0];
			double const p1 = cf[4 * p + 1] * bs[1];
			double const p2 = cf[4 * p + 2] * bs[2];
			double const p3 = cf[4 * p + 3] * bs[3];
			best_fit[j] = (float) (((p0 + p1) + p2) + p3);
		}
	}
	for (size_t i = 0; i < c->n; ++i) {
		residual[i] = data[i] - best_fit[i];
	}
}

/* baseline.cc:1063-1098 */
static size_t clip_data(size_t num_boundary, size_t const *boundary, float const *resid,
		float lower, float upper, unsigned char *mask, size_t *clipped) {
	size_t num = 0;
	size_t piece_start = 0;
	for (size_t ib = 1; ib < num_boundary; ++ib) {
		size_t const piece_end = boundary[ib] - 1;
		for (size_t i = piece_start; i < piece_end; ++i) {
			if (mask[i]) {
				float const d = resid[i];
				if ((d - lower) * (upper - d) < 0.0) {
					mask[i] = 0;
					clipped[num++] = i;
				}
			}
		}
		piece_start = piece_end + 1;
	}
	return num;
}

/* baseline.cc:1115-1224 DoLSQFit. Returns the sakura_Status; *lsq is the LSQFitStatus. */
static int do_lsq_fit(Ctx *c, float const *data, unsigned char const *mask, size_t nb_model,
		double const *basis, size_t K, size_t num_boundary, size_t *boundary, unsigned nfit,
		float clip_sigma, unsigned char *final_mask, float *rms, int *lsq) {
	size_t const n = c->n;
	size_t num_unmasked = 0;
	for (size_t i = 0; i < n; ++i) {
		if (mask[i]) {
			++num_unmasked;
		}
	}
	if (num_unmasked < K) {
		*lsq = kLsqNotEnoughData;
		return kStatusNG;
	}
	num_unmasked = n;
	size_t num_clipped = n;
	for (size_t i = 0; i < n; ++i) {
		final_mask[i] = mask[i];
	}
	double rms_d = 0.0;
	for (unsigned it = 1; it <= nfit; ++it) {
		if (num_unmasked < K) {
			*lsq = kLsqNotEnoughData;
			return kStatusNG;
		}
		if (num_unmasked <= num_clipped) {
			if (get_lsq_coefficients(n, data, final_mask, nb_model, basis, K, c->use_idx, c->G,
					c->b) != kStatusOK) {
				return kStatusNG;
			}
		} else {
			update_lsq_coefficients(data, mask, num_clipped, c->clipped, nb_model, basis, K,
					c->use_idx, c->G, c->b);
		}
		solve_full_piv_lu(K, c->G, c->b, c->coeff, c->lu, c->cvec, c->tr);
		if (c->type == kCubicSpline) {
			full_spline_coefficients(num_boundary, boundary, c->coeff, c->coeff_full);
			spline_model_and_residual(c, num_boundary, boundary, data, c->coeff_full, c->best_fit,
					c->residual);
		} else {
			model_and_residual(c, K, data, c->coeff, c->best_fit, c->residual);
		}
		OracleStatistics st;
		oracle_compute_accurate_statistics(n, c->residual, final_mask, &st);
		double const cnt = (double) st.count;
		double const mean = st.sum / cnt;
		rms_d = sqrt(fabs(fma(-mean, mean, st.square_sum / cnt))); /* :1200-1202, vfnmadd231sd */
		if (it < nfit) {
			float const thr = (float) ((double) clip_sigma * rms_d);
			float const lower = (float) (mean - (double) thr);
			float const upper = (float) (mean + (double) thr);
			num_clipped = clip_data(num_boundary, boundary, c->residual, lower, upper, final_mask,
					c->clipped);
			if (num_clipped == 0) {
				break;
			}
			num_unmasked = st.count - num_clipped;
		}
	}
	*rms = (float) rms_d;
	return kStatusOK;
}

/* baseline.cc:1239-1254 */
static void set_sinusoid_index(size_t num_nwave, size_t const *nwave, size_t *use_idx) {
	size_t iuse = 0, i = 0;
	if (nwave[0] == 0) {
		use_idx[iuse++] = nwave[i++];
	}
	for (; i < num_nwave; ++i) {
		use_idx[iuse++] = 2 * nwave[i] - 1;
		use_idx[iuse++] = 2 * nwave[i];
	}
}

/* baseline.cc:1268-1329 LSQFit + C entry :1687-1742 / :1835-1893 (argument checks that the
 * row drivers can reach are restated; pointer/alignment checks are the caller's business) */
static int fit_poly_or_sinusoid(Ctx *c, size_t order, size_t num_nwave, size_t const *nwave,
		float const *data, unsigned char const *mask, float clip, unsigned nfit, size_t num_coeff,
		double *coeff, float *best_fit, float *residual, unsigned char *final_mask, float *rms,
		int *lsq) {
	*lsq = kLsqNG;
	if (c->type == kSinusoid) {
		if (num_nwave == 0) {
			return kStatusInvalidArgument;
		}
		for (size_t i = 1; i < num_nwave; ++i) {
			if (nwave[i - 1] >= nwave[i]) {
				return kStatusInvalidArgument;
			}
		}
		if (nwave[num_nwave - 1] > c->param) {
			return kStatusInvalidArgument;
		}
		size_t const need = (nwave[0] == 0) ? (2 * num_nwave - 1) : (2 * num_nwave);
		if (!(need <= num_coeff) || !(num_coeff <= c->num_bases)) {
			return kStatusInvalidArgument;
		}
	} else {
		if (order > c->param) {
			return kStatusInvalidArgument;
		}
		if (coeff != NULL && num_coeff != order + 1) {
			return kStatusInvalidArgument;
		}
	}
	if (!(0.0f < clip)) {
		return kStatusInvalidArgument;
	}
	if (c->type == kSinusoid) {
		set_sinusoid_index(num_nwave, nwave, c->use_idx);
	}
	if (nfit == 0) {
		*lsq = kLsqOK;
		return kStatusOK;
	}
	size_t boundary[2] = { 0, c->n };
	int const st = do_lsq_fit(c, data, mask, c->num_bases, c->basis, num_coeff, 2, boundary, nfit,
			clip, final_mask, rms, lsq);
	if (st != kStatusOK) {
		return st;
	}
	if (coeff != NULL) {
		for (size_t i = 0; i < num_coeff; ++i) {
			coeff[i] = c->coeff[i];
		}
		if (c->type == kPolynomial) {
			double const max_x = (double) (c->n - 1);
			double factor = 1.0;
			for (size_t i = 0; i < num_coeff; ++i) {
				coeff[i] /= factor;
				factor *= max_x;
			}
		}
	}
	if (best_fit != NULL) {
		memcpy(best_fit, c->best_fit, sizeof(float) * c->n);
	}
	if (residual != NULL) {
		memcpy(residual, c->residual, sizeof(float) * c->n);
	}
	*lsq = kLsqOK;
	return kStatusOK;
}

/* baseline.cc:1342-1409 LSQFitCubicSpline + C entry :1761-1816 */
static int fit_spline(Ctx *c, size_t num_pieces, float const *data, unsigned char const *mask,
		float clip, unsigned nfit, double *coeff, float *best_fit, float *residual,
		unsigned char *final_mask, float *rms, size_t *boundary, int *lsq) {
	*lsq = kLsqNG;
	if (num_pieces == 0 || num_pieces > c->param || num_pieces + 3 > c->n || !(0.0f < clip)) {
		return kStatusInvalidArgument;
	}
	size_t const num_boundary = num_pieces + 1;
	if (get_boundaries(c->n, mask, num_boundary, boundary) != 0) {
		return kStatusNG;
	}
	set_spline_basis(c->n, num_boundary, c->lsq_max, boundary, c->cspline_basis);
	if (nfit == 0) {
		*lsq = kLsqOK;
		return kStatusOK;
	}
	size_t const K = num_pieces + 3;
	int const st = do_lsq_fit(c, data, mask, c->lsq_max, c->cspline_basis, K, num_boundary, boundary,
			nfit, clip, final_mask, rms, lsq);
	if (st != kStatusOK) {
		return st;
	}
	if (coeff != NULL) {
		double const max_x = (double) (c->n - 1);
		for (size_t i = 0; i < num_pieces; ++i) {
			double factor = 1.0;
			for (size_t j = 0; j < 4; ++j) {
				coeff[4 * i + j] = c->coeff_full[4 * i + j] / factor;
				factor *= max_x;
			}
		}
	}
	if (best_fit != NULL) {
		memcpy(best_fit, c->best_fit, sizeof(float) * c->n);
	}
	if (residual != NULL) {
		memcpy(residual, c->residual, sizeof(float) * c->n);
	}
	*lsq = kLsqOK;
	return kStatusOK;
}

/* ------------------------------------------------------------------ row drivers (same ABI as
 * oracle/ref_driver.cc so that oracle/refbind.py binds either library) */

static int pick_threads(int nthreads) {
#if defined(_OPENMP)
	return nthreads <= 0 ? omp_get_max_threads() : nthreads;
#else
	(void) nthreads;
	return 1;
#endif
}

int ref_init(void) {
	return 0;
}

int ref_max_threads(void) {
#if defined(_OPENMP)
	return omp_get_max_threads();
#else
	return 1;
#endif
}

int ref_lsqfit_polynomial_batch(int lsqfit_type, uint16_t context_order, uint16_t order,
		size_t num_spectra, size_t num_data, float const *data, size_t data_stride,
		unsigned char const *mask, size_t mask_stride, float clip_threshold_sigma,
		uint16_t num_fitting_max, size_t num_coeff, double *coeff, float *best_fit, float *residual,
		unsigned char *final_mask, float *rms, int32_t *status, int32_t *lsqfit_status,
		int nthreads) {
	if ((size_t) context_order + 1 > num_data) {
		return -1;
	}
	int failed = 0;
	nthreads = pick_threads(nthreads);
#pragma omp parallel num_threads(nthreads)
	{
		Ctx *c = ctx_create(lsqfit_type == 1 ? kChebyshev : kPolynomial, context_order, num_data);
		if (c == NULL) {
#pragma omp atomic write
			failed = 1;
		}
#pragma omp for schedule(static)
		for (ptrdiff_t r = 0; r < (ptrdiff_t) num_spectra; ++r) {
			if (c == NULL) {
				continue;
			}
			int lsq = kLsqNG;
			float rms_row = rms[r];
			status[r] = fit_poly_or_sinusoid(c, order, 0, NULL, data + r * data_stride,
					mask + r * mask_stride, clip_threshold_sigma, num_fitting_max, num_coeff,
					coeff ? coeff + r * num_coeff : NULL, best_fit ? best_fit + r * data_stride : NULL,
					residual ? residual + r * data_stride : NULL, final_mask + r * mask_stride,
					&rms_row, &lsq);
			rms[r] = rms_row;
			lsqfit_status[r] = lsq;
		}
		ctx_destroy(c);
	}
	return failed ? -1 : 0;
}

int ref_lsqfit_cubicspline_batch(uint16_t context_npiece, size_t num_pieces, size_t num_spectra,
		size_t num_data, float const *data, size_t data_stride, unsigned char const *mask,
		size_t mask_stride, float clip_threshold_sigma, uint16_t num_fitting_max, double *coeff,
		float *best_fit, float *residual, unsigned char *final_mask, float *rms, size_t *boundary,
		int32_t *status, int32_t *lsqfit_status, int nthreads) {
	if (context_npiece == 0 || (size_t) context_npiece + 3 > num_data) {
		return -1;
	}
	int failed = 0;
	nthreads = pick_threads(nthreads);
#pragma omp parallel num_threads(nthreads)
	{
		Ctx *c = ctx_create(kCubicSpline, context_npiece, num_data);
		if (c == NULL) {
#pragma omp atomic write
			failed = 1;
		}
		size_t *bd = (size_t *) malloc(sizeof(size_t) * (num_pieces + 2));
#pragma omp for schedule(static)
		for (ptrdiff_t r = 0; r < (ptrdiff_t) num_spectra; ++r) {
			if (c == NULL || bd == NULL) {
				continue;
			}
			int lsq = kLsqNG;
			float rms_row = rms[r];
			size_t *bout = boundary + r * (num_pieces + 1);
			memcpy(bd, bout, sizeof(size_t) * (num_pieces + 1));
			status[r] = fit_spline(c, num_pieces, data + r * data_stride, mask + r * mask_stride,
					clip_threshold_sigma, num_fitting_max, coeff ? coeff + r * num_pieces * 4 : NULL,
					best_fit ? best_fit + r * data_stride : NULL,
					residual ? residual + r * data_stride : NULL, final_mask + r * mask_stride,
					&rms_row, bd, &lsq);
			memcpy(bout, bd, sizeof(size_t) * (num_pieces + 1));
			rms[r] = rms_row;
			lsqfit_status[r] = lsq;
		}
		free(bd);
		ctx_destroy(c);
	}
	return failed ? -1 : 0;
}

int ref_lsqfit_sinusoid_batch(uint16_t context_nwave, size_t num_nwave, size_t const *nwave,
		size_t num_spectra, size_t num_data, float const *data, size_t data_stride,
		unsigned char const *mask, size_t mask_stride, float clip_threshold_sigma,
		uint16_t num_fitting_max, size_t num_coeff, double *coeff, float *best_fit, float *residual,
		unsigned char *final_mask, float *rms, int32_t *status, int32_t *lsqfit_status,
		int nthreads) {
	if (2 * (size_t) context_nwave + 2 > num_data) {
		return -1;
	}
	int failed = 0;
	nthreads = pick_threads(nthreads);
#pragma omp parallel num_threads(nthreads)
	{
		Ctx *c = ctx_create(kSinusoid, context_nwave, num_data);
		if (c == NULL) {
#pragma omp atomic write
			failed = 1;
		}
#pragma omp for schedule(static)
		for (ptrdiff_t r = 0; r < (ptrdiff_t) num_spectra; ++r) {
			if (c == NULL) {
				continue;
			}
			int lsq = kLsqNG;
			float rms_row = rms[r];
			status[r] = fit_poly_or_sinusoid(c, 0, num_nwave, nwave, data + r * data_stride,
					mask + r * mask_stride, clip_threshold_sigma, num_fitting_max, num_coeff,
					coeff ? coeff + r * num_coeff : NULL, best_fit ? best_fit + r * data_stride : NULL,
					residual ? residual + r * data_stride : NULL, final_mask + r * mask_stride,
					&rms_row, &lsq);
			rms[r] = rms_row;
			lsqfit_status[r] = lsq;
		}
		ctx_destroy(c);
	}
	return failed ? -1 : 0;
}

int ref_statistics_batch(int accurate, size_t num_spectra, size_t num_data, float const *data,
		size_t data_stride, unsigned char const *mask, size_t mask_stride,
		OracleStatistics *result, int32_t *status, int nthreads) {
	nthreads = pick_threads(nthreads);
#pragma omp parallel for schedule(static) num_threads(nthreads)
	for (ptrdiff_t r = 0; r < (ptrdiff_t) num_spectra; ++r) {
		if (num_data > (size_t) INT32_MAX) {
			status[r] = kStatusInvalidArgument; /* statistics.cc:855 */
			continue;
		}
		if (accurate) {
			oracle_compute_accurate_statistics(num_data, data + r * data_stride,
					mask + r * mask_stride, &result[r]);
		} else {
			oracle_compute_statistics(num_data, data + r * data_stride, mask + r * mask_stride,
					&result[r]);
		}
		status[r] = kStatusOK;
	}
	return 0;
}

/* Position-switch calibration, normalization.cc:128-139 (in place, AVX: mul(factor, sub(result,
 * reference)) then div) and :147-176 (out of place, scalar expression factor * (target -
 * reference) / reference): float subtract, multiply, divide in that order in both. Argument
 * rules of :224-243 / :253-271: num_data == 0 is kOK before anything else, then NULL pointers and
 * unaligned arrays (32 bytes) are kInvalidArgument. */
int ref_calibrate_batch(size_t num_spectra, size_t num_data, float const *scaling,
		size_t scaling_stride, float const_scaling, float const *data, size_t data_stride,
		float const *reference, size_t reference_stride, float *result, size_t result_stride,
		int32_t *status) {
	for (size_t r = 0; r < num_spectra; ++r) {
		float const *s = scaling ? scaling + r * scaling_stride : NULL;
		float const *t = data ? data + r * data_stride : NULL;
		float const *ref = reference ? reference + r * reference_stride : NULL;
		float *out = result ? result + r * result_stride : NULL;
		if (num_data == 0) {
			status[r] = kStatusOK;
			continue;
		}
		if (t == NULL || ref == NULL || out == NULL || ((uintptr_t) t % 32) != 0
				|| ((uintptr_t) ref % 32) != 0 || ((uintptr_t) out % 32) != 0
				|| (s != NULL && ((uintptr_t) s % 32) != 0)) {
			status[r] = kStatusInvalidArgument;
			continue;
		}
		for (size_t i = 0; i < num_data; ++i) {
			float const factor = s ? s[i] : const_scaling;
			float const diff = t[i] - ref[i];
			float const prod = factor * diff;
			out[i] = prod / ref[i];
		}
		status[r] = kStatusOK;
	}
	return 0;
}

/* ------------------------------------------------------------------ boolean filters
 * bool_filter_collection.cc. Ranges: the reference's 8-wide SIMD body evaluates
 * "lower <= x && x <= upper" (:180-200, :360-380), the scalar remainder of num_data % 8 elements
 * and the generic form for more than 16 conditions evaluate "(x - lower) * (upper - x) >= 0"
 * (:213-220, :330-345; int products wrap). Argument rules :535-600: data, result and both bound
 * arrays non-NULL and 32-byte aligned, also for zero conditions. */
static int in_range_float(int inclusive, int product, float x, float lo, float hi) {
	if (product) {
		float const v = (x - lo) * (hi - x);
		return inclusive ? (v >= 0.0f) : (v > 0.0f);
	}
	/* the AVX float comparisons are the unordered ones (_CMP_NGT_UQ / _CMP_NGE_UQ,
	 * libsakura/packed_operation.h:1423-1436): a NaN element compares "in range" */
	return inclusive ? (!(lo > x) && !(x > hi)) : (!(lo >= x) && !(x >= hi));
}

static int in_range_int(int inclusive, int product, int x, int lo, int hi) {
	if (product) {
		int const v = (int) (((unsigned) x - (unsigned) lo) * ((unsigned) hi - (unsigned) x));
		return inclusive ? (v >= 0) : (v > 0);
	}
	return inclusive ? (lo <= x && x <= hi) : (lo < x && x < hi);
}

int ref_bool_filter(int op, int type, size_t num_data, void const *data, size_t num_condition,
		void const *lower, void const *upper, uint32_t threshold_bits, unsigned char *result) {
	if (data == NULL || result == NULL || ((uintptr_t) data % 32) != 0 || ((uintptr_t) result % 32) != 0) {
		return kStatusInvalidArgument;
	}
	if (op <= 1 && (lower == NULL || upper == NULL || ((uintptr_t) lower % 32) != 0
			|| ((uintptr_t) upper % 32) != 0)) {
		return kStatusInvalidArgument;
	}
	float tf;
	int ti;
	memcpy(&tf, &threshold_bits, 4);
	memcpy(&ti, &threshold_bits, 4);
	size_t const compare_end = (num_condition <= 16) ? (num_data / 8) * 8 : 0;
	for (size_t i = 0; i < num_data; ++i) {
		int r = 0;
		if (op <= 1) {
			int const product = i >= compare_end;
			for (size_t j = 0; j < num_condition; ++j) {
				if (type == 0) {
					r |= in_range_float(op == 0, product, ((float const *) data)[i],
							((float const *) lower)[j], ((float const *) upper)[j]);
				} else {
					r |= in_range_int(op == 0, product, ((int const *) data)[i], ((int const *) lower)[j],
							((int const *) upper)[j]);
				}
			}
		} else if (op <= 5) {
			if (type == 0) {
				float const x = ((float const *) data)[i];
				r = op == 2 ? x > tf : op == 3 ? x >= tf : op == 4 ? x < tf : x <= tf;
			} else {
				int const x = ((int const *) data)[i];
				r = op == 2 ? x > ti : op == 3 ? x >= ti : op == 4 ? x < ti : x <= ti;
			}
		} else if (op == 6) {
			uint32_t bits;
			memcpy(&bits, (float const *) data + i, 4);
			r = (bits & 0x7F800000u) != 0x7F800000u; /* :838-842 */
		} else if (op == 7) {
			r = type == 2 ? ((unsigned char const *) data)[i] != 0 : ((uint32_t const *) data)[i] != 0;
		} else {
			result[i] = (unsigned char) (((unsigned char const *) data)[i] ^ 1u); /* :424 */
			continue;
		}
		result[i] = (unsigned char) r;
	}
	return kStatusOK;
}

/* ------------------------------------------------------------------ bit operations
 * bit_operation.cc:60-190, the expressions as written there: x = (T) mask - 1 from the edit-mask
 * byte; arguments NULL or not 32-byte aligned -> kInvalidArgument (:41-52). */
int ref_bit_operation(int op, int wide, uint32_t bit_mask, size_t num_data, void const *data,
		unsigned char const *edit_mask, void *result) {
	if (data == NULL || result == NULL || edit_mask == NULL || ((uintptr_t) data % 32) != 0
			|| ((uintptr_t) result % 32) != 0 || ((uintptr_t) edit_mask % 32) != 0) {
		return kStatusInvalidArgument;
	}
	for (size_t i = 0; i < num_data; ++i) {
		uint32_t const d = wide ? ((uint32_t const *) data)[i] : ((unsigned char const *) data)[i];
		uint32_t const x = (uint32_t) edit_mask[i] - 1u;
		uint32_t const nx = ~x;
		uint32_t r;
		switch (op) {
		case 0: r = d & (bit_mask | x); break;
		case 1: r = (d ^ nx) & (bit_mask | x); break;
		case 2: r = (d ^ nx) | (bit_mask & nx); break;
		case 3: r = d ^ nx; break;
		case 4: r = d | (bit_mask & nx); break;
		default: r = d ^ (bit_mask & nx); break;
		}
		if (wide) {
			((uint32_t *) result)[i] = r;
		} else {
			((unsigned char *) result)[i] = (unsigned char) r;
		}
	}
	return kStatusOK;
}

/* ------------------------------------------------------------------ interpolation
 * interpolation.cc, nearest and linear (polynomial of order 0 is nearest, :1419-1422). For every
 * array: the valid base points are gathered in ascending order of position (:905-922); no valid
 * point -> mask false, data untouched (:1182-1185); one -> constant (:1186-1190); otherwise points
 * below the first / at or above the last valid position take that point's value (:1207-1225) and a
 * point with b_j <= x < b_{j+1} is interpolated between j and j+1 (Locate, :164-232): nearest
 * :625-637, linear :691-704 with the final multiply-add fused as the Release build compiles it.
 * The result for a point depends on its position only, so the reference's handling of descending
 * interpolated positions (reverse, then write back to front) needs no restatement. */
int ref_interpolate(int y_axis, int method, int polynomial_order, size_t num_base,
		double const *base_position, size_t num_array, float const *base_data,
		unsigned char const *base_mask, size_t num_interpolated, double const *interpolated_position,
		float *interpolated_data, unsigned char *interpolated_mask) {
	if (num_base == 0) { /* :1305-1309 */
		return kStatusInvalidArgument;
	}
	if (num_interpolated == 0 || num_array == 0) { /* :1311-1316 */
		return kStatusOK;
	}
	void const *arrays[6] = { base_position, base_data, interpolated_position, interpolated_data,
			base_mask, interpolated_mask };
	for (int k = 0; k < 6; ++k) { /* :1318-1335 */
		if (arrays[k] == NULL || ((uintptr_t) arrays[k] % 32) != 0) {
			return kStatusInvalidArgument;
		}
	}
	int linear;
	if (method == 0 || (method == 2 && polynomial_order == 0)) {
		linear = 0;
	} else if (method == 1) {
		linear = 1;
	} else {
		return -1; /* polynomial / spline interpolation: not restated */
	}
	int const ascending = base_position[num_base - 1] > base_position[0];
	double *vx = (double *) malloc(num_base * sizeof(double));
	float *vy = (float *) malloc(num_base * sizeof(float));
	for (size_t a = 0; a < num_array; ++a) {
		size_t m = 0;
		for (size_t k = 0; k < num_base; ++k) {
			size_t const kb = ascending ? k : num_base - 1 - k;
			size_t const e = y_axis ? kb * num_array + a : a * num_base + kb;
			if (base_mask[e]) {
				vx[m] = base_position[kb];
				vy[m] = base_data[e];
				++m;
			}
		}
		for (size_t i = 0; i < num_interpolated; ++i) {
			size_t const e = y_axis ? i * num_array + a : a * num_interpolated + i;
			interpolated_mask[e] = (m != 0);
			if (m == 0) {
				continue;
			}
			double const x = interpolated_position[i];
			float v;
			if (m == 1 || x < vx[0]) {
				v = vy[0];
			} else if (x >= vx[m - 1]) {
				v = vy[m - 1];
			} else {
				size_t j = 0;
				while (vx[j + 1] <= x) {
					++j;
				}
				if (linear) {
					double const dydx = (double) (vy[j + 1] - vy[j]) / (vx[j + 1] - vx[j]);
					v = (float) fma(dydx, x - vx[j], (double) vy[j]);
				} else {
					double const midpoint = (vx[j + 1] + vx[j]) / 2.0;
					v = (x > midpoint) ? vy[j + 1] : vy[j];
				}
			}
			interpolated_data[e] = v;
		}
	}
	free(vx);
	free(vy);
	return kStatusOK;
}

/* standalone building blocks for the known-answer tests (numeric_operation.cc:867-968) */
int oracle_get_lsq_coefficients(size_t num_data, float const *data, unsigned char const *mask,
		size_t num_model_bases, double const *basis, size_t num_lsq_bases, size_t const *use_idx,
		double *lsq_matrix, double *lsq_vector) {
	return get_lsq_coefficients(num_data, data, mask, num_model_bases, basis, num_lsq_bases, use_idx,
			lsq_matrix, lsq_vector);
}

int oracle_update_lsq_coefficients(size_t num_data, float const *data, unsigned char const *mask,
		size_t num_exclude, size_t const *exclude, size_t num_model_bases, double const *basis,
		size_t num_lsq_bases, size_t const *use_idx, double *lsq_matrix, double *lsq_vector) {
	(void) num_data;
	update_lsq_coefficients(data, mask, num_exclude, exclude, num_model_bases, basis, num_lsq_bases,
			use_idx, lsq_matrix, lsq_vector);
	return kStatusOK;
}

int oracle_solve_by_lu(size_t n, double const *A, double const *b, double *x) {
	double *lu = (double *) malloc(sizeof(double) * n * n);
	double *c = (double *) malloc(sizeof(double) * n);
	ptrdiff_t *tr = (ptrdiff_t *) malloc(sizeof(ptrdiff_t) * 2 * n);
	if (!lu || !c || !tr) {
		free(lu);
		free(c);
		free(tr);
		return 3;
	}
	solve_full_piv_lu(n, A, b, x, lu, c, tr);
	free(lu);
	free(c);
	free(tr);
	return kStatusOK;
}
