// TEST INFRASTRUCTURE ONLY (oracle). Never linked into, imported or called by
// the shipped sakura_b200 library.
//
// OpenMP-over-rows driver around the *reference's own object code*
// (baseline.cc, numeric_operation.cc, statistics.cc, gen_util.cc compiled from
// /root/reference where they lie; see oracle/Makefile). The reference library
// is single-threaded per call and a sakura_LSQFitContextFloat "can not be
// shared between threads" (libsakura/src/libsakura/sakura.h:1613-1615), so
// rows fan out with one context per thread -- the same way the reference's own
// callers parallelise (bench/e2eReduce/src/main.cc:520-608,
// test/bench-lsqfit-casacore.cc).
//
// Every entry point takes plain pointers/sizes so that tests and bench.py can
// bind it with ctypes. Rows whose start address is not LIBSAKURA-aligned are
// staged through per-thread aligned scratch (the reference rejects unaligned
// pointers: baseline.cc:1704-1722).
#include <libsakura/sakura.h>

#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <vector>

#if defined(_OPENMP)
#include <omp.h>
#endif

namespace {

struct AlignedBuf {
	void *p = nullptr;
	size_t cap = 0;
	~AlignedBuf() {
		free(p);
	}
	void *Get(size_t bytes) {
		if (bytes > cap) {
			free(p);
			p = nullptr;
			size_t const a = sakura_GetAlignment();
			if (posix_memalign(&p, a < 64 ? 64 : a, bytes < 64 ? 64 : bytes) != 0) {
				p = nullptr;
			}
			cap = bytes;
		}
		return p;
	}
};

struct RowIO {
	AlignedBuf data, mask, coeff, best_fit, residual, final_mask, boundary;
};

template<typename T>
T const *StageIn(T const *src, size_t n, AlignedBuf *buf) {
	if (sakura_IsAligned(src)) {
		return src;
	}
	T *dst = static_cast<T *>(buf->Get(n * sizeof(T)));
	memcpy(dst, src, n * sizeof(T));
	return dst;
}

template<typename T>
T *StageOut(T *dst, size_t n, AlignedBuf *buf) {
	if (dst == nullptr || sakura_IsAligned(dst)) {
		return dst;
	}
	T *tmp = static_cast<T *>(buf->Get(n * sizeof(T)));
	// keep "untouched on failure" semantics observable through the staging copy
	memcpy(tmp, dst, n * sizeof(T));
	return tmp;
}

template<typename T>
void Commit(T *dst, T const *staged, size_t n) {
	if (dst != nullptr && staged != dst) {
		memcpy(dst, staged, n * sizeof(T));
	}
}

int SetThreads(int nthreads) {
#if defined(_OPENMP)
	if (nthreads <= 0) {
		nthreads = omp_get_max_threads();
	}
	return nthreads;
#else
	(void) nthreads;
	return 1;
#endif
}

} // namespace

extern "C" {

int ref_init(void) {
	return static_cast<int>(sakura_Initialize(nullptr, nullptr));
}

int ref_max_threads(void) {
#if defined(_OPENMP)
	return omp_get_max_threads();
#else
	return 1;
#endif
}

/**
 * Rows through sakura_LSQFitPolynomialFloat (baseline.cc:1687).
 * lsqfit_type: 0 polynomial, 1 Chebyshev (sakura.h:1591-1603).
 * Strides are in elements. Any of coeff/best_fit/residual may be NULL.
 * status[r] receives the sakura_Status, lsqfit_status[r] the sakura_LSQFitStatus.
 * Returns 0, or -1 if a context could not be created.
 */
int ref_lsqfit_polynomial_batch(int lsqfit_type, uint16_t context_order, uint16_t order,
		size_t num_spectra, size_t num_data, float const *data, size_t data_stride,
		bool const *mask, size_t mask_stride, float clip_threshold_sigma,
		uint16_t num_fitting_max, size_t num_coeff, double *coeff, float *best_fit,
		float *residual, bool *final_mask, float *rms, int32_t *status,
		int32_t *lsqfit_status, int nthreads) {
	int failed = 0;
	nthreads = SetThreads(nthreads);
#pragma omp parallel num_threads(nthreads)
	{
		sakura_LSQFitContextFloat *ctx = nullptr;
		sakura_Status cs = sakura_CreateLSQFitContextPolynomialFloat(
				static_cast<sakura_LSQFitType>(lsqfit_type), context_order, num_data, &ctx);
		if (cs != sakura_Status_kOK) {
#pragma omp atomic write
			failed = 1;
		}
		RowIO io;
#pragma omp for schedule(static)
		for (ptrdiff_t r = 0; r < static_cast<ptrdiff_t>(num_spectra); ++r) {
			if (ctx == nullptr) {
				continue;
			}
			float const *d = StageIn(data + r * data_stride, num_data, &io.data);
			bool const *m = StageIn(mask + r * mask_stride, num_data, &io.mask);
			double *c_dst = coeff ? coeff + r * num_coeff : nullptr;
			float *b_dst = best_fit ? best_fit + r * data_stride : nullptr;
			float *r_dst = residual ? residual + r * data_stride : nullptr;
			bool *f_dst = final_mask + r * mask_stride;
			double *c = StageOut(c_dst, num_coeff, &io.coeff);
			float *b = StageOut(b_dst, num_data, &io.best_fit);
			float *rs = StageOut(r_dst, num_data, &io.residual);
			bool *f = StageOut(f_dst, num_data, &io.final_mask);
			sakura_LSQFitStatus ls = sakura_LSQFitStatus_kNG;
			float rms_row = rms[r];
			sakura_Status st = sakura_LSQFitPolynomialFloat(ctx, order, num_data, d, m,
					clip_threshold_sigma, num_fitting_max, num_coeff, c, b, rs, f, &rms_row,
					&ls);
			Commit(c_dst, c, num_coeff);
			Commit(b_dst, b, num_data);
			Commit(r_dst, rs, num_data);
			Commit(f_dst, f, num_data);
			rms[r] = rms_row;
			status[r] = static_cast<int32_t>(st);
			lsqfit_status[r] = static_cast<int32_t>(ls);
		}
		if (ctx != nullptr) {
			(void) sakura_DestroyLSQFitContextFloat(ctx);
		}
	}
	return failed ? -1 : 0;
}

/**
 * Rows through sakura_LSQFitCubicSplineFloat (baseline.cc:1761).
 * coeff is [num_spectra][num_pieces][4], boundary is [num_spectra][num_pieces+1].
 */
int ref_lsqfit_cubicspline_batch(uint16_t context_npiece, size_t num_pieces,
		size_t num_spectra, size_t num_data, float const *data, size_t data_stride,
		bool const *mask, size_t mask_stride, float clip_threshold_sigma,
		uint16_t num_fitting_max, double *coeff, float *best_fit, float *residual,
		bool *final_mask, float *rms, size_t *boundary, int32_t *status,
		int32_t *lsqfit_status, int nthreads) {
	int failed = 0;
	nthreads = SetThreads(nthreads);
#pragma omp parallel num_threads(nthreads)
	{
		sakura_LSQFitContextFloat *ctx = nullptr;
		sakura_Status cs = sakura_CreateLSQFitContextCubicSplineFloat(context_npiece,
				num_data, &ctx);
		if (cs != sakura_Status_kOK) {
#pragma omp atomic write
			failed = 1;
		}
		RowIO io;
		size_t const ncoeff = num_pieces * 4;
		size_t const nb = num_pieces + 1;
#pragma omp for schedule(static)
		for (ptrdiff_t r = 0; r < static_cast<ptrdiff_t>(num_spectra); ++r) {
			if (ctx == nullptr) {
				continue;
			}
			float const *d = StageIn(data + r * data_stride, num_data, &io.data);
			bool const *m = StageIn(mask + r * mask_stride, num_data, &io.mask);
			double *c_dst = coeff ? coeff + r * ncoeff : nullptr;
			float *b_dst = best_fit ? best_fit + r * data_stride : nullptr;
			float *r_dst = residual ? residual + r * data_stride : nullptr;
			bool *f_dst = final_mask + r * mask_stride;
			size_t *bd_dst = boundary + r * nb;
			double *c = StageOut(c_dst, ncoeff, &io.coeff);
			float *b = StageOut(b_dst, num_data, &io.best_fit);
			float *rs = StageOut(r_dst, num_data, &io.residual);
			bool *f = StageOut(f_dst, num_data, &io.final_mask);
			size_t *bd = StageOut(bd_dst, nb, &io.boundary);
			sakura_LSQFitStatus ls = sakura_LSQFitStatus_kNG;
			float rms_row = rms[r];
			sakura_Status st = sakura_LSQFitCubicSplineFloat(ctx, num_pieces, num_data, d, m,
					clip_threshold_sigma, num_fitting_max,
					reinterpret_cast<double (*)[4]>(c), b, rs, f, &rms_row, bd, &ls);
			Commit(c_dst, c, ncoeff);
			Commit(b_dst, b, num_data);
			Commit(r_dst, rs, num_data);
			Commit(f_dst, f, num_data);
			Commit(bd_dst, bd, nb);
			rms[r] = rms_row;
			status[r] = static_cast<int32_t>(st);
			lsqfit_status[r] = static_cast<int32_t>(ls);
		}
		if (ctx != nullptr) {
			(void) sakura_DestroyLSQFitContextFloat(ctx);
		}
	}
	return failed ? -1 : 0;
}

/**
 * Rows through sakura_LSQFitSinusoidFloat (baseline.cc:1835).
 */
int ref_lsqfit_sinusoid_batch(uint16_t context_nwave, size_t num_nwave, size_t const *nwave,
		size_t num_spectra, size_t num_data, float const *data, size_t data_stride,
		bool const *mask, size_t mask_stride, float clip_threshold_sigma,
		uint16_t num_fitting_max, size_t num_coeff, double *coeff, float *best_fit,
		float *residual, bool *final_mask, float *rms, int32_t *status,
		int32_t *lsqfit_status, int nthreads) {
	int failed = 0;
	nthreads = SetThreads(nthreads);
#pragma omp parallel num_threads(nthreads)
	{
		sakura_LSQFitContextFloat *ctx = nullptr;
		sakura_Status cs = sakura_CreateLSQFitContextSinusoidFloat(context_nwave, num_data,
				&ctx);
		if (cs != sakura_Status_kOK) {
#pragma omp atomic write
			failed = 1;
		}
		RowIO io;
#pragma omp for schedule(static)
		for (ptrdiff_t r = 0; r < static_cast<ptrdiff_t>(num_spectra); ++r) {
			if (ctx == nullptr) {
				continue;
			}
			float const *d = StageIn(data + r * data_stride, num_data, &io.data);
			bool const *m = StageIn(mask + r * mask_stride, num_data, &io.mask);
			double *c_dst = coeff ? coeff + r * num_coeff : nullptr;
			float *b_dst = best_fit ? best_fit + r * data_stride : nullptr;
			float *r_dst = residual ? residual + r * data_stride : nullptr;
			bool *f_dst = final_mask + r * mask_stride;
			double *c = StageOut(c_dst, num_coeff, &io.coeff);
			float *b = StageOut(b_dst, num_data, &io.best_fit);
			float *rs = StageOut(r_dst, num_data, &io.residual);
			bool *f = StageOut(f_dst, num_data, &io.final_mask);
			sakura_LSQFitStatus ls = sakura_LSQFitStatus_kNG;
			float rms_row = rms[r];
			sakura_Status st = sakura_LSQFitSinusoidFloat(ctx, num_nwave, nwave, num_data, d,
					m, clip_threshold_sigma, num_fitting_max, num_coeff, c, b, rs, f,
					&rms_row, &ls);
			Commit(c_dst, c, num_coeff);
			Commit(b_dst, b, num_data);
			Commit(r_dst, rs, num_data);
			Commit(f_dst, f, num_data);
			rms[r] = rms_row;
			status[r] = static_cast<int32_t>(st);
			lsqfit_status[r] = static_cast<int32_t>(ls);
		}
		if (ctx != nullptr) {
			(void) sakura_DestroyLSQFitContextFloat(ctx);
		}
	}
	return failed ? -1 : 0;
}

/**
 * Rows through sakura_ComputeStatisticsFloat (accurate == 0, statistics.cc:871)
 * or sakura_ComputeAccurateStatisticsFloat (accurate != 0, statistics.cc:879).
 * result is an array of sakura_StatisticsResultFloat (48 bytes each).
 */
int ref_statistics_batch(int accurate, size_t num_spectra, size_t num_data, float const *data,
		size_t data_stride, bool const *mask, size_t mask_stride,
		sakura_StatisticsResultFloat *result, int32_t *status, int nthreads) {
	nthreads = SetThreads(nthreads);
#pragma omp parallel num_threads(nthreads)
	{
		RowIO io;
#pragma omp for schedule(static)
		for (ptrdiff_t r = 0; r < static_cast<ptrdiff_t>(num_spectra); ++r) {
			float const *d = StageIn(data + r * data_stride, num_data, &io.data);
			bool const *m = StageIn(mask + r * mask_stride, num_data, &io.mask);
			sakura_Status st =
					accurate ?
							sakura_ComputeAccurateStatisticsFloat(num_data, d, m, &result[r]) :
							sakura_ComputeStatisticsFloat(num_data, d, m, &result[r]);
			status[r] = static_cast<int32_t>(st);
		}
	}
	return 0;
}

/**
 * sakura_CalibrateDataWith{Array,Const}ScalingFloat (normalization.cc:206-282) row by row. A row
 * stride of 0 shares one reference / scaling array; scaling == NULL selects const_scaling.
 * Rows must be aligned (sakura_GetAlignment()); the reference rejects unaligned arrays.
 */
int ref_calibrate_batch(size_t num_spectra, size_t num_data, float const *scaling,
		size_t scaling_stride, float const_scaling, float const *data, size_t data_stride,
		float const *reference, size_t reference_stride, float *result, size_t result_stride,
		int32_t *status) {
	for (size_t r = 0; r < num_spectra; ++r) {
		sakura_Status st;
		if (scaling != nullptr) {
			st = sakura_CalibrateDataWithArrayScalingFloat(num_data, scaling + r * scaling_stride,
					data + r * data_stride, reference + r * reference_stride,
					result + r * result_stride);
		} else {
			st = sakura_CalibrateDataWithConstScalingFloat(const_scaling, num_data,
					data + r * data_stride, reference + r * reference_stride,
					result + r * result_stride);
		}
		status[r] = static_cast<int32_t>(st);
	}
	return 0;
}

/*
 * Boolean-filter builders (bool_filter_collection.cc), one call of the reference's own function.
 * op: 0 ranges inclusive, 1 ranges exclusive, 2 >, 3 >=, 4 <, 5 <=, 6 finite (SetFalseIfNanOrInf),
 * 7 non-zero (Uint8ToBool / Uint32ToBool), 8 InvertBool. type: 0 float, 1 int, 2 uint8, 3 uint32.
 */
int ref_bool_filter(int op, int type, size_t num_data, void const *data, size_t num_condition,
		void const *lower, void const *upper, uint32_t threshold_bits, bool *result) {
	float tf;
	int ti;
	memcpy(&tf, &threshold_bits, 4);
	memcpy(&ti, &threshold_bits, 4);
	float const *df = static_cast<float const *>(data);
	int const *di = static_cast<int const *>(data);
	float const *lf = static_cast<float const *>(lower), *uf = static_cast<float const *>(upper);
	int const *li = static_cast<int const *>(lower), *ui = static_cast<int const *>(upper);
	switch (op) {
	case 0:
		return type == 0 ? sakura_SetTrueIfInRangesInclusiveFloat(num_data, df, num_condition, lf, uf, result)
				: sakura_SetTrueIfInRangesInclusiveInt(num_data, di, num_condition, li, ui, result);
	case 1:
		return type == 0 ? sakura_SetTrueIfInRangesExclusiveFloat(num_data, df, num_condition, lf, uf, result)
				: sakura_SetTrueIfInRangesExclusiveInt(num_data, di, num_condition, li, ui, result);
	case 2:
		return type == 0 ? sakura_SetTrueIfGreaterThanFloat(num_data, df, tf, result)
				: sakura_SetTrueIfGreaterThanInt(num_data, di, ti, result);
	case 3:
		return type == 0 ? sakura_SetTrueIfGreaterThanOrEqualsFloat(num_data, df, tf, result)
				: sakura_SetTrueIfGreaterThanOrEqualsInt(num_data, di, ti, result);
	case 4:
		return type == 0 ? sakura_SetTrueIfLessThanFloat(num_data, df, tf, result)
				: sakura_SetTrueIfLessThanInt(num_data, di, ti, result);
	case 5:
		return type == 0 ? sakura_SetTrueIfLessThanOrEqualsFloat(num_data, df, tf, result)
				: sakura_SetTrueIfLessThanOrEqualsInt(num_data, di, ti, result);
	case 6:
		return sakura_SetFalseIfNanOrInfFloat(num_data, df, result);
	case 7:
		return type == 2 ? sakura_Uint8ToBool(num_data, static_cast<uint8_t const *>(data), result)
				: sakura_Uint32ToBool(num_data, static_cast<uint32_t const *>(data), result);
	case 8:
		return sakura_InvertBool(num_data, static_cast<bool const *>(data), result);
	default:
		return -1;
	}
}

/* bit_operation.cc: op 0 And, 1 ConverseNonImplication, 2 Implication, 3 Not, 4 Or, 5 Xor */
int ref_bit_operation(int op, int wide, uint32_t bit_mask, size_t num_data, void const *data,
		bool const *edit_mask, void *result) {
	uint8_t const *d8 = static_cast<uint8_t const *>(data);
	uint32_t const *d32 = static_cast<uint32_t const *>(data);
	uint8_t *r8 = static_cast<uint8_t *>(result);
	uint32_t *r32 = static_cast<uint32_t *>(result);
	uint8_t const m8 = static_cast<uint8_t>(bit_mask);
	switch (op) {
	case 0:
		return wide ? sakura_OperateBitwiseAndUint32(bit_mask, num_data, d32, edit_mask, r32)
				: sakura_OperateBitwiseAndUint8(m8, num_data, d8, edit_mask, r8);
	case 1:
		return wide ? sakura_OperateBitwiseConverseNonImplicationUint32(bit_mask, num_data, d32, edit_mask, r32)
				: sakura_OperateBitwiseConverseNonImplicationUint8(m8, num_data, d8, edit_mask, r8);
	case 2:
		return wide ? sakura_OperateBitwiseImplicationUint32(bit_mask, num_data, d32, edit_mask, r32)
				: sakura_OperateBitwiseImplicationUint8(m8, num_data, d8, edit_mask, r8);
	case 3:
		return wide ? sakura_OperateBitwiseNotUint32(num_data, d32, edit_mask, r32)
				: sakura_OperateBitwiseNotUint8(num_data, d8, edit_mask, r8);
	case 4:
		return wide ? sakura_OperateBitwiseOrUint32(bit_mask, num_data, d32, edit_mask, r32)
				: sakura_OperateBitwiseOrUint8(m8, num_data, d8, edit_mask, r8);
	case 5:
		return wide ? sakura_OperateBitwiseXorUint32(bit_mask, num_data, d32, edit_mask, r32)
				: sakura_OperateBitwiseXorUint8(m8, num_data, d8, edit_mask, r8);
	default:
		return -1;
	}
}

/* sakura_InterpolateXAxisFloat (y_axis == 0) / sakura_InterpolateYAxisFloat (y_axis != 0). */
int ref_interpolate(int y_axis, int method, int polynomial_order, size_t num_base,
		double const *base_position, size_t num_array, float const *base_data, bool const *base_mask,
		size_t num_interpolated, double const *interpolated_position, float *interpolated_data,
		bool *interpolated_mask) {
	sakura_InterpolationMethod const m = static_cast<sakura_InterpolationMethod>(method);
	uint8_t const order = static_cast<uint8_t>(polynomial_order);
	return y_axis ? sakura_InterpolateYAxisFloat(m, order, num_base, base_position, num_array, base_data,
					base_mask, num_interpolated, interpolated_position, interpolated_data,
					interpolated_mask)
			: sakura_InterpolateXAxisFloat(m, order, num_base, base_position, num_array, base_data,
					base_mask, num_interpolated, interpolated_position, interpolated_data,
					interpolated_mask);
}

} // extern "C"
