/*
 * sakura_b200.h -- C ABI of libsakura_b200.so
 *
 * B200-native (sm_100a CUDA) drop-in for ONE hot path of libsakura 5.3: masked
 * least-squares baseline fitting of single-dish spectra with iterative
 * sigma-clipping, plus the residual statistics it depends on.
 *
 * Every `sakura_*` symbol below keeps the name, argument order, argument
 * meaning, alignment contract and status codes of the reference declaration it
 * replaces (cited as `sakura.h:<line>` = libsakura/src/libsakura/sakura.h and
 * `baseline.cc:<line>` etc. of tnakazato/sakura).  A caller written against
 * <libsakura/sakura.h> for this path links against this library unchanged.
 * The `...Batch` entry points are new: many spectra per call, row r of every
 * array belonging to spectrum r; results equal a loop over the single-row call.
 *
 * Pointers are HOST pointers unless the function name ends in `Device`.
 * No torch / C++ types cross this boundary.  The header is valid C99.
 */
#ifndef SAKURA_B200_H_
#define SAKURA_B200_H_

#include <stddef.h>
#include <stdint.h>
#include <stdbool.h>
#include <sys/types.h>

#ifdef __cplusplus
extern "C" {
# define SAKURA_B200_NOEXCEPT noexcept
#else
# define SAKURA_B200_NOEXCEPT
#endif

/* ---- status codes ------------------------------------------------------ */

/** Replaces sakura_Status, sakura.h:64-82. */
typedef enum {
	sakura_Status_kOK = 0,
	sakura_Status_kNG = 1,
	sakura_Status_kInvalidArgument = 2,
	sakura_Status_kNoMemory = 3,
	sakura_Status_kUnknownError = 99
} sakura_Status;

/** Replaces sakura_LSQFitStatus, sakura.h:1570-1586. */
typedef enum {
	sakura_LSQFitStatus_kOK = 0,
	sakura_LSQFitStatus_kNG = 1,
	sakura_LSQFitStatus_kNotEnoughData = 2,
	sakura_LSQFitStatus_kNumElements
} sakura_LSQFitStatus;

/** Replaces sakura_LSQFitType, sakura.h:1591-1603. */
typedef enum {
	sakura_LSQFitType_kPolynomial,
	sakura_LSQFitType_kChebyshev,
	sakura_LSQFitType_kNumElements
} sakura_LSQFitType;

/** Replaces sakura_StatisticsResultFloat, sakura.h:208-237 (48 bytes). */
typedef struct {
	size_t count;
	double sum;
	double square_sum;
	float min;
	float max;
	ssize_t index_of_min;
	ssize_t index_of_max;
} sakura_StatisticsResultFloat;

/** Opaque, as in sakura.h:1608. One context per thread (sakura.h:1613-1615). */
struct sakura_LSQFitContextFloat;

/* ---- lifecycle, allocator hooks, alignment ----------------------------- */

typedef void *(*sakura_UserAllocator)(size_t size); /* sakura.h:98 */
typedef void (*sakura_UserDeallocator)(void *pointer); /* sakura.h:111 */

/** Replaces sakura_Initialize, sakura.h:125 (gen_util.cc:95-104).
 *  Host-side context memory goes through the hooks; device memory does not. */
sakura_Status sakura_Initialize(sakura_UserAllocator allocator,
		sakura_UserDeallocator deallocator) SAKURA_B200_NOEXCEPT;
/** Replaces sakura_CleanUp, sakura.h:137. Releases cached device buffers. */
void sakura_CleanUp(void) SAKURA_B200_NOEXCEPT;
/** Replaces sakura_GetAlignment, sakura.h:160. Returns 32 (the AVX build's value,
 *  localdef.h:75-86) so that callers' buffers are laid out identically. */
size_t sakura_GetAlignment(void) SAKURA_B200_NOEXCEPT;
/** Replaces sakura_IsAligned, sakura.h:151. */
bool sakura_IsAligned(void const *ptr) SAKURA_B200_NOEXCEPT;
/** Replace sakura_AlignAny / AlignFloat / AlignDouble, sakura.h:170-197. */
void *sakura_AlignAny(size_t size_of_arena, void *arena, size_t size_required)
		SAKURA_B200_NOEXCEPT;
float *sakura_AlignFloat(size_t elements_in_arena, float *arena, size_t elements_required)
		SAKURA_B200_NOEXCEPT;
double *sakura_AlignDouble(size_t elements_in_arena, double *arena,
		size_t elements_required) SAKURA_B200_NOEXCEPT;

/* ---- statistics --------------------------------------------------------- */

/** Replaces sakura_ComputeStatisticsFloat, sakura.h:254 (statistics.cc:871). */
sakura_Status sakura_ComputeStatisticsFloat(size_t num_data, float const data[],
		bool const is_valid[], sakura_StatisticsResultFloat *result) SAKURA_B200_NOEXCEPT;
/** Replaces sakura_ComputeAccurateStatisticsFloat, sakura.h:264 (statistics.cc:879). */
sakura_Status sakura_ComputeAccurateStatisticsFloat(size_t num_data, float const data[],
		bool const is_valid[], sakura_StatisticsResultFloat *result) SAKURA_B200_NOEXCEPT;
/** NEW. Row r: data + r*data_stride, is_valid + r*mask_stride (strides in elements). */
sakura_Status sakura_ComputeStatisticsFloatBatch(size_t num_spectra, size_t num_data,
		float const data[], size_t data_stride, bool const is_valid[], size_t mask_stride,
		sakura_StatisticsResultFloat result[/*num_spectra*/]) SAKURA_B200_NOEXCEPT;
/** NEW. Same with device pointers; asynchronous on `cuda_stream` (a cudaStream_t, may be 0). */
sakura_Status sakura_ComputeStatisticsFloatBatchDevice(size_t num_spectra, size_t num_data,
		float const d_data[], size_t data_stride, bool const d_is_valid[], size_t mask_stride,
		sakura_StatisticsResultFloat d_result[/*num_spectra*/], void *cuda_stream)
				SAKURA_B200_NOEXCEPT;

/* ---- normal-equation building blocks ------------------------------------ */

/** Replaces sakura_GetLSQCoefficientsDouble, sakura.h:1403 (numeric_operation.cc:867). */
sakura_Status sakura_GetLSQCoefficientsDouble(size_t const num_data, float const data[],
		bool const mask[], size_t const num_model_bases, double const basis_data[],
		size_t const num_lsq_bases, size_t const use_bases_idx[], double lsq_matrix[],
		double lsq_vector[]) SAKURA_B200_NOEXCEPT;
/** Replaces sakura_UpdateLSQCoefficientsDouble, sakura.h:1482 (numeric_operation.cc:905). */
sakura_Status sakura_UpdateLSQCoefficientsDouble(size_t const num_data, float const data[],
		bool const mask[], size_t const num_exclude_indices, size_t const exclude_indices[],
		size_t const num_model_bases, double const basis_data[], size_t const num_lsq_bases,
		size_t const use_bases_idx[], double lsq_matrix[], double lsq_vector[])
				SAKURA_B200_NOEXCEPT;
/** Replaces sakura_SolveSimultaneousEquationsByLUDouble, sakura.h:1520
 *  (numeric_operation.cc:948): full-pivot LU, rank-revealing, never fails. */
sakura_Status sakura_SolveSimultaneousEquationsByLUDouble(size_t num_equations,
		double const in_matrix[], double const in_vector[], double out[])
				SAKURA_B200_NOEXCEPT;

/* ---- fit contexts --------------------------------------------------------- */

/** Replaces sakura_CreateLSQFitContextPolynomialFloat, sakura.h:1635 (baseline.cc:1515). */
sakura_Status sakura_CreateLSQFitContextPolynomialFloat(sakura_LSQFitType const lsqfit_type,
		uint16_t order, size_t num_data, struct sakura_LSQFitContextFloat **context)
				SAKURA_B200_NOEXCEPT;
/** Replaces sakura_CreateLSQFitContextCubicSplineFloat, sakura.h:1660 (baseline.cc:1555). */
sakura_Status sakura_CreateLSQFitContextCubicSplineFloat(uint16_t npiece, size_t num_data,
		struct sakura_LSQFitContextFloat **context) SAKURA_B200_NOEXCEPT;
/** Replaces sakura_CreateLSQFitContextSinusoidFloat, sakura.h:1686 (baseline.cc:1591). */
sakura_Status sakura_CreateLSQFitContextSinusoidFloat(uint16_t nwave, size_t num_data,
		struct sakura_LSQFitContextFloat **context) SAKURA_B200_NOEXCEPT;
/** Replaces sakura_DestroyLSQFitContextFloat, sakura.h:1701 (baseline.cc:1627). */
sakura_Status sakura_DestroyLSQFitContextFloat(struct sakura_LSQFitContextFloat *context)
		SAKURA_B200_NOEXCEPT;
/** Replaces sakura_GetNumberOfCoefficientsFloat, sakura.h:2084 (baseline.cc:1646). */
sakura_Status sakura_GetNumberOfCoefficientsFloat(
		struct sakura_LSQFitContextFloat const *context, uint16_t order, size_t *num_coeff)
				SAKURA_B200_NOEXCEPT;

/* ---- single-spectrum fits (reference signatures) -------------------------- */

/** Replaces sakura_LSQFitPolynomialFloat, sakura.h:1773-1781 (baseline.cc:1687). */
sakura_Status sakura_LSQFitPolynomialFloat(struct sakura_LSQFitContextFloat const *context,
		uint16_t order, size_t num_data, float const data[], bool const mask[],
		float clip_threshold_sigma, uint16_t num_fitting_max, size_t num_coeff,
		double coeff[], float best_fit[], float residual[], bool final_mask[], float *rms,
		sakura_LSQFitStatus *lsqfit_status) SAKURA_B200_NOEXCEPT;
/** Replaces sakura_LSQFitCubicSplineFloat, sakura.h:1859-1868 (baseline.cc:1761). */
sakura_Status sakura_LSQFitCubicSplineFloat(struct sakura_LSQFitContextFloat const *context,
		size_t num_pieces, size_t num_data, float const data[], bool const mask[],
		float clip_threshold_sigma, uint16_t num_fitting_max, double coeff[][4],
		float best_fit[], float residual[], bool final_mask[], float *rms, size_t boundary[],
		sakura_LSQFitStatus *lsqfit_status) SAKURA_B200_NOEXCEPT;
/** Replaces sakura_LSQFitSinusoidFloat, sakura.h:1947-1955 (baseline.cc:1835). */
sakura_Status sakura_LSQFitSinusoidFloat(struct sakura_LSQFitContextFloat const *context,
		size_t num_nwave, size_t const nwave[], size_t num_data, float const data[],
		bool const mask[], float clip_threshold_sigma, uint16_t num_fitting_max,
		size_t num_coeff, double coeff[], float best_fit[], float residual[],
		bool final_mask[], float *rms, sakura_LSQFitStatus *lsqfit_status)
				SAKURA_B200_NOEXCEPT;
/** Replaces sakura_SubtractPolynomialFloat, sakura.h:1981 (baseline.cc:1902). */
sakura_Status sakura_SubtractPolynomialFloat(struct sakura_LSQFitContextFloat const *context,
		size_t num_data, float const data[], size_t num_coeff, double const coeff[],
		float out[]) SAKURA_B200_NOEXCEPT;
/** Replaces sakura_SubtractCubicSplineFloat, sakura.h:2018 (baseline.cc:1941). */
sakura_Status sakura_SubtractCubicSplineFloat(struct sakura_LSQFitContextFloat const *context,
		size_t num_data, float const data[], size_t num_pieces, double const coeff[][4],
		size_t const boundary[], float out[]) SAKURA_B200_NOEXCEPT;
/** Replaces sakura_SubtractSinusoidFloat, sakura.h:2059 (baseline.cc:1983). */
sakura_Status sakura_SubtractSinusoidFloat(struct sakura_LSQFitContextFloat const *context,
		size_t num_data, float const data[], size_t num_nwave, size_t const nwave[],
		size_t num_coeff, double const coeff[], float out[]) SAKURA_B200_NOEXCEPT;

/* ---- NEW: batched many-spectra fits ---------------------------------------
 *
 * Row r of `data`/`best_fit`/`residual` starts at element r*data_stride, of
 * `mask`/`final_mask` at r*mask_stride, of `coeff` at r*num_coeff (spline:
 * r*num_pieces*4), of `boundary` at r*(num_pieces+1).  `rms`, `status`,
 * `lsqfit_status` have one entry per row.  Per-row semantics are exactly those of
 * the single-row call (baseline.cc:1115-1224, 1268-1409):
 *   status[r] = the sakura_Status the single-row call would have returned
 *               (kOK, or kNG when the row has too few usable channels),
 *   lsqfit_status[r] = kOK / kNG / kNotEnoughData,
 *   a failing row leaves its coeff/best_fit/residual/rms untouched, and its
 *   final_mask untouched unless clipping had already started (SURVEY A-13).
 * The return value is kInvalidArgument when the *call* is malformed (NULL /
 * unaligned base pointers, strides that break row alignment, order/num_coeff
 * out of range), kNoMemory when device memory cannot be had, kUnknownError on a
 * CUDA failure, else kOK -- also when some rows failed: inspect status[].
 * coeff, best_fit, residual may be NULL.  Host variant: base pointers must be
 * sakura-aligned and the strides multiples that keep every row aligned.
 * Device variant: device pointers, 16-byte aligned rows, asynchronous on
 * `cuda_stream` (cudaStream_t; 0 = default stream); in-place final_mask==mask is
 * allowed. residual (or best_fit, not both) may be the data array itself for the
 * polynomial call where its row-resident kernels apply (num_data % 4 == 0, <= 8192,
 * order <= 7): every row is read completely before it is written (the reference
 * documents in-place use, sakura.h:1755-1766); any other overlap with data is
 * kInvalidArgument. The context carries device scratch (hand-over lists, staging for
 * the general kernel): successive device calls on ONE context must be issued on the
 * same stream, or be separated by a synchronisation of the earlier one.
 */
sakura_Status sakura_LSQFitPolynomialFloatBatch(
		struct sakura_LSQFitContextFloat const *context, uint16_t order, size_t num_spectra,
		size_t num_data, float const data[], size_t data_stride, bool const mask[],
		size_t mask_stride, float clip_threshold_sigma, uint16_t num_fitting_max,
		size_t num_coeff, double coeff[], float best_fit[], float residual[],
		bool final_mask[], float rms[], int32_t status[], int32_t lsqfit_status[])
				SAKURA_B200_NOEXCEPT;
sakura_Status sakura_LSQFitPolynomialFloatBatchDevice(
		struct sakura_LSQFitContextFloat const *context, uint16_t order, size_t num_spectra,
		size_t num_data, float const d_data[], size_t data_stride, bool const d_mask[],
		size_t mask_stride, float clip_threshold_sigma, uint16_t num_fitting_max,
		size_t num_coeff, double d_coeff[], float d_best_fit[], float d_residual[],
		bool d_final_mask[], float d_rms[], int32_t d_status[], int32_t d_lsqfit_status[],
		void *cuda_stream) SAKURA_B200_NOEXCEPT;

sakura_Status sakura_LSQFitCubicSplineFloatBatch(
		struct sakura_LSQFitContextFloat const *context, size_t num_pieces, size_t num_spectra,
		size_t num_data, float const data[], size_t data_stride, bool const mask[],
		size_t mask_stride, float clip_threshold_sigma, uint16_t num_fitting_max,
		double coeff[/*num_spectra*num_pieces*4*/], float best_fit[], float residual[],
		bool final_mask[], float rms[], size_t boundary[/*num_spectra*(num_pieces+1)*/],
		int32_t status[], int32_t lsqfit_status[]) SAKURA_B200_NOEXCEPT;
sakura_Status sakura_LSQFitCubicSplineFloatBatchDevice(
		struct sakura_LSQFitContextFloat const *context, size_t num_pieces, size_t num_spectra,
		size_t num_data, float const d_data[], size_t data_stride, bool const d_mask[],
		size_t mask_stride, float clip_threshold_sigma, uint16_t num_fitting_max,
		double d_coeff[], float d_best_fit[], float d_residual[], bool d_final_mask[],
		float d_rms[], size_t d_boundary[], int32_t d_status[], int32_t d_lsqfit_status[],
		void *cuda_stream) SAKURA_B200_NOEXCEPT;

sakura_Status sakura_LSQFitSinusoidFloatBatch(
		struct sakura_LSQFitContextFloat const *context, size_t num_nwave,
		size_t const nwave[], size_t num_spectra, size_t num_data, float const data[],
		size_t data_stride, bool const mask[], size_t mask_stride, float clip_threshold_sigma,
		uint16_t num_fitting_max, size_t num_coeff, double coeff[], float best_fit[],
		float residual[], bool final_mask[], float rms[], int32_t status[],
		int32_t lsqfit_status[]) SAKURA_B200_NOEXCEPT;
sakura_Status sakura_LSQFitSinusoidFloatBatchDevice(
		struct sakura_LSQFitContextFloat const *context, size_t num_nwave,
		size_t const nwave[/*host*/], size_t num_spectra, size_t num_data,
		float const d_data[], size_t data_stride, bool const d_mask[], size_t mask_stride,
		float clip_threshold_sigma, uint16_t num_fitting_max, size_t num_coeff,
		double d_coeff[], float d_best_fit[], float d_residual[], bool d_final_mask[],
		float d_rms[], int32_t d_status[], int32_t d_lsqfit_status[], void *cuda_stream)
				SAKURA_B200_NOEXCEPT;

/* ---- NEW: one call, several GPUs ---------------------------------------------------
 *
 * The host-pointer batch calls above drive the CURRENT device. The ...MultiDevice variants take the
 * same arguments plus a device list and split the rows into one contiguous block per device
 * (block g = rows [R*g/G, R*(g+1)/G)): what the reference's callers do with worker threads
 * (bench/e2eReduce/src/main.cc:290, 520-608 fans blocks of rows out to a thread pool) happens inside
 * the library, with one host thread, one replica of the context and one copy/compute pipeline per
 * GPU. Rows are independent: there is no exchange between the devices.
 *   num_devices == 0 or device_ids == NULL : every visible CUDA device (the first num_devices of
 *                                            them when num_devices > 0);
 *   otherwise device_ids[0..num_devices)   : distinct CUDA device ordinals (kInvalidArgument if not).
 * Results, per-row status semantics and argument rules are those of the ...Batch call; the calling
 * thread's current device is restored. Page-locked host buffers (cudaHostAlloc / cudaHostRegister,
 * portable) are needed for the copies of the devices to overlap. */
sakura_Status sakura_LSQFitPolynomialFloatBatchMultiDevice(
		struct sakura_LSQFitContextFloat const *context, uint16_t order, size_t num_spectra,
		size_t num_data, float const data[], size_t data_stride, bool const mask[],
		size_t mask_stride, float clip_threshold_sigma, uint16_t num_fitting_max,
		size_t num_coeff, double coeff[], float best_fit[], float residual[],
		bool final_mask[], float rms[], int32_t status[], int32_t lsqfit_status[],
		size_t num_devices, int const device_ids[]) SAKURA_B200_NOEXCEPT;
sakura_Status sakura_LSQFitCubicSplineFloatBatchMultiDevice(
		struct sakura_LSQFitContextFloat const *context, size_t num_pieces, size_t num_spectra,
		size_t num_data, float const data[], size_t data_stride, bool const mask[],
		size_t mask_stride, float clip_threshold_sigma, uint16_t num_fitting_max,
		double coeff[/*num_spectra*num_pieces*4*/], float best_fit[], float residual[],
		bool final_mask[], float rms[], size_t boundary[/*num_spectra*(num_pieces+1)*/],
		int32_t status[], int32_t lsqfit_status[], size_t num_devices, int const device_ids[])
				SAKURA_B200_NOEXCEPT;
sakura_Status sakura_LSQFitSinusoidFloatBatchMultiDevice(
		struct sakura_LSQFitContextFloat const *context, size_t num_nwave,
		size_t const nwave[], size_t num_spectra, size_t num_data, float const data[],
		size_t data_stride, bool const mask[], size_t mask_stride, float clip_threshold_sigma,
		uint16_t num_fitting_max, size_t num_coeff, double coeff[], float best_fit[],
		float residual[], bool final_mask[], float rms[], int32_t status[],
		int32_t lsqfit_status[], size_t num_devices, int const device_ids[]) SAKURA_B200_NOEXCEPT;

/* ---- position-switch calibration ---------------------------
 * result = scaling_factor * (data - reference) / reference in float: subtract, multiply, true
 * divide (normalization.cc:128-139). */

/** Replaces sakura_CalibrateDataWithArrayScalingFloat, sakura.h:1137 (normalization.cc:206). */
sakura_Status sakura_CalibrateDataWithArrayScalingFloat(size_t num_data,
		float const scaling_factor[/*num_data*/], float const data[/*num_data*/],
		float const reference[/*num_data*/], float result[/*num_data*/]) SAKURA_B200_NOEXCEPT;
/** Replaces sakura_CalibrateDataWithConstScalingFloat, sakura.h:1178 (normalization.cc:245). */
sakura_Status sakura_CalibrateDataWithConstScalingFloat(float scaling_factor, size_t num_data,
		float const data[/*num_data*/], float const reference[/*num_data*/],
		float result[/*num_data*/]) SAKURA_B200_NOEXCEPT;
/** NEW. Rows of device arrays; strides in elements, a stride of 0 shares one reference /
 *  scaling array between all rows; d_scaling_factor NULL selects const_scaling_factor.
 *  d_result may alias d_data or d_reference. Asynchronous on `cuda_stream`. */
sakura_Status sakura_CalibrateDataFloatBatchDevice(size_t num_spectra, size_t num_data,
		float const d_scaling_factor[], size_t scaling_stride, float const_scaling_factor,
		float const d_data[], size_t data_stride, float const d_reference[],
		size_t reference_stride, float d_result[], size_t result_stride, void *cuda_stream)
				SAKURA_B200_NOEXCEPT;
/** NEW. sakura_LSQFitPolynomialFloatBatchDevice on the calibrated spectra
 *  scaling * (d_data - d_reference) / d_reference, which are never written to memory: the fit
 *  kernel calibrates each row while loading it. With flag_nan_or_inf the mask the fit uses is
 *  d_mask AND "calibrated value is finite" (sakura_SetFalseIfNanOrInfFloat, sakura.h:636, merged in).
 *  residual / best_fit refer to the calibrated spectrum. Results are identical to calling the
 *  calibration, the NaN/Inf filter and the fit one after the other. */
sakura_Status sakura_CalibrateAndLSQFitPolynomialFloatBatchDevice(
		struct sakura_LSQFitContextFloat const *context, uint16_t order, size_t num_spectra,
		size_t num_data, float const d_scaling_factor[], size_t scaling_stride,
		float const_scaling_factor, float const d_data[], size_t data_stride,
		float const d_reference[], size_t reference_stride, bool flag_nan_or_inf,
		bool const d_mask[], size_t mask_stride, float clip_threshold_sigma,
		uint16_t num_fitting_max, size_t num_coeff, double d_coeff[], float d_best_fit[],
		float d_residual[], bool d_final_mask[], float d_rms[], int32_t d_status[],
		int32_t d_lsqfit_status[], void *cuda_stream) SAKURA_B200_NOEXCEPT;

/* ---- boolean-filter builders: the masks the fit consumes ---------------
 * Same names, signatures, argument checks (NULL / unaligned data, result or bound arrays =>
 * kInvalidArgument, also with zero conditions) and element semantics as the reference's
 * sakura.h:418-682 / bool_filter_collection.cc: host arrays in, host arrays out (staged through
 * the GPU). The *Device variants take device arrays (any alignment; 16-byte aligned arrays take
 * the 128-bit path), host bound arrays, run asynchronously on `cuda_stream`, and can AND the
 * filter into an existing mask in the same pass (`d_and_with`, may be NULL or alias `d_result`):
 * the merge every caller does next (bench/e2eReduce/src/main.cc:150-229). */

/** Replaces sakura_SetTrueIfInRangesInclusiveFloat, sakura.h:418 (bool_filter_collection.cc:666). */
sakura_Status sakura_SetTrueIfInRangesInclusiveFloat(size_t num_data, float const data[],
		size_t num_condition, float const lower_bounds[], float const upper_bounds[],
		bool result[]) SAKURA_B200_NOEXCEPT;
/** Replaces sakura_SetTrueIfInRangesInclusiveInt, sakura.h:427. */
sakura_Status sakura_SetTrueIfInRangesInclusiveInt(size_t num_data, int const data[],
		size_t num_condition, int const lower_bounds[], int const upper_bounds[],
		bool result[]) SAKURA_B200_NOEXCEPT;
/** Replaces sakura_SetTrueIfInRangesExclusiveFloat, sakura.h:467. */
sakura_Status sakura_SetTrueIfInRangesExclusiveFloat(size_t num_data, float const data[],
		size_t num_condition, float const lower_bounds[], float const upper_bounds[],
		bool result[]) SAKURA_B200_NOEXCEPT;
/** Replaces sakura_SetTrueIfInRangesExclusiveInt, sakura.h:476. */
sakura_Status sakura_SetTrueIfInRangesExclusiveInt(size_t num_data, int const data[],
		size_t num_condition, int const lower_bounds[], int const upper_bounds[],
		bool result[]) SAKURA_B200_NOEXCEPT;
/** Replace sakura_SetTrueIf{GreaterThan,GreaterThanOrEquals,LessThan,LessThanOrEquals}{Float,Int},
 *  sakura.h:504-636. */
sakura_Status sakura_SetTrueIfGreaterThanFloat(size_t num_data, float const data[], float threshold,
		bool result[]) SAKURA_B200_NOEXCEPT;
sakura_Status sakura_SetTrueIfGreaterThanInt(size_t num_data, int const data[], int threshold,
		bool result[]) SAKURA_B200_NOEXCEPT;
sakura_Status sakura_SetTrueIfGreaterThanOrEqualsFloat(size_t num_data, float const data[],
		float threshold, bool result[]) SAKURA_B200_NOEXCEPT;
sakura_Status sakura_SetTrueIfGreaterThanOrEqualsInt(size_t num_data, int const data[], int threshold,
		bool result[]) SAKURA_B200_NOEXCEPT;
sakura_Status sakura_SetTrueIfLessThanFloat(size_t num_data, float const data[], float threshold,
		bool result[]) SAKURA_B200_NOEXCEPT;
sakura_Status sakura_SetTrueIfLessThanInt(size_t num_data, int const data[], int threshold,
		bool result[]) SAKURA_B200_NOEXCEPT;
sakura_Status sakura_SetTrueIfLessThanOrEqualsFloat(size_t num_data, float const data[],
		float threshold, bool result[]) SAKURA_B200_NOEXCEPT;
sakura_Status sakura_SetTrueIfLessThanOrEqualsInt(size_t num_data, int const data[], int threshold,
		bool result[]) SAKURA_B200_NOEXCEPT;
/** Replaces sakura_SetFalseIfNanOrInfFloat, sakura.h:651. */
sakura_Status sakura_SetFalseIfNanOrInfFloat(size_t num_data, float const data[],
		bool result[]) SAKURA_B200_NOEXCEPT;
/** Replace sakura_Uint8ToBool / sakura_Uint32ToBool / sakura_InvertBool, sakura.h:667-691
 *  (InvertBool may run in place). */
sakura_Status sakura_Uint8ToBool(size_t num_data, uint8_t const data[], bool result[]) SAKURA_B200_NOEXCEPT;
sakura_Status sakura_Uint32ToBool(size_t num_data, uint32_t const data[], bool result[]) SAKURA_B200_NOEXCEPT;
sakura_Status sakura_InvertBool(size_t num_data, bool const data[], bool result[]) SAKURA_B200_NOEXCEPT;

/** NEW. Device-array variants. `inclusive` selects <= / <; bounds are host arrays. */
sakura_Status sakura_SetTrueIfInRangesFloatDevice(bool inclusive, size_t num_data, float const d_data[],
		size_t num_condition, float const lower_bounds[], float const upper_bounds[],
		bool const d_and_with[], bool d_result[], void *cuda_stream) SAKURA_B200_NOEXCEPT;
sakura_Status sakura_SetTrueIfInRangesIntDevice(bool inclusive, size_t num_data, int const d_data[],
		size_t num_condition, int const lower_bounds[], int const upper_bounds[],
		bool const d_and_with[], bool d_result[], void *cuda_stream) SAKURA_B200_NOEXCEPT;
/** NEW. `comparison`: 0 x > t, 1 x >= t, 2 x < t, 3 x <= t. */
sakura_Status sakura_SetTrueIfComparesFloatDevice(int comparison, size_t num_data, float const d_data[],
		float threshold, bool const d_and_with[], bool d_result[], void *cuda_stream) SAKURA_B200_NOEXCEPT;
sakura_Status sakura_SetTrueIfComparesIntDevice(int comparison, size_t num_data, int const d_data[],
		int threshold, bool const d_and_with[], bool d_result[], void *cuda_stream) SAKURA_B200_NOEXCEPT;
sakura_Status sakura_SetFalseIfNanOrInfFloatDevice(size_t num_data, float const d_data[],
		bool const d_and_with[], bool d_result[], void *cuda_stream) SAKURA_B200_NOEXCEPT;
sakura_Status sakura_Uint8ToBoolDevice(size_t num_data, uint8_t const d_data[], bool const d_and_with[],
		bool d_result[], void *cuda_stream) SAKURA_B200_NOEXCEPT;
sakura_Status sakura_Uint32ToBoolDevice(size_t num_data, uint32_t const d_data[],
		bool const d_and_with[], bool d_result[], void *cuda_stream) SAKURA_B200_NOEXCEPT;
sakura_Status sakura_InvertBoolDevice(size_t num_data, bool const d_data[], bool const d_and_with[],
		bool d_result[], void *cuda_stream) SAKURA_B200_NOEXCEPT;

/* ---- bit operations with an edit mask: from a bool mask back to flag bytes ----------------
 * sakura.h:693-956 / bit_operation.cc: result[i] = edit_mask[i] ? (data[i] OP bit_mask) : data[i];
 * same names, signatures and argument rules (NULL / unaligned => kInvalidArgument); `result` may be
 * `data` (bench/e2eReduce/src/main.cc:59-78 ORs bit 7 into the flag bytes where the mask says so). */
sakura_Status sakura_OperateBitwiseAndUint8(uint8_t bit_mask, size_t num_data, uint8_t const data[],
		bool const edit_mask[], uint8_t result[]) SAKURA_B200_NOEXCEPT;
sakura_Status sakura_OperateBitwiseAndUint32(uint32_t bit_mask, size_t num_data, uint32_t const data[],
		bool const edit_mask[], uint32_t result[]) SAKURA_B200_NOEXCEPT;
sakura_Status sakura_OperateBitwiseConverseNonImplicationUint8(uint8_t bit_mask, size_t num_data,
		uint8_t const data[], bool const edit_mask[], uint8_t result[]) SAKURA_B200_NOEXCEPT;
sakura_Status sakura_OperateBitwiseConverseNonImplicationUint32(uint32_t bit_mask, size_t num_data,
		uint32_t const data[], bool const edit_mask[], uint32_t result[]) SAKURA_B200_NOEXCEPT;
sakura_Status sakura_OperateBitwiseImplicationUint8(uint8_t bit_mask, size_t num_data, uint8_t const data[],
		bool const edit_mask[], uint8_t result[]) SAKURA_B200_NOEXCEPT;
sakura_Status sakura_OperateBitwiseImplicationUint32(uint32_t bit_mask, size_t num_data,
		uint32_t const data[], bool const edit_mask[], uint32_t result[]) SAKURA_B200_NOEXCEPT;
sakura_Status sakura_OperateBitwiseNotUint8(size_t num_data, uint8_t const data[], bool const edit_mask[],
		uint8_t result[]) SAKURA_B200_NOEXCEPT;
sakura_Status sakura_OperateBitwiseNotUint32(size_t num_data, uint32_t const data[], bool const edit_mask[],
		uint32_t result[]) SAKURA_B200_NOEXCEPT;
sakura_Status sakura_OperateBitwiseOrUint8(uint8_t bit_mask, size_t num_data, uint8_t const data[],
		bool const edit_mask[], uint8_t result[]) SAKURA_B200_NOEXCEPT;
sakura_Status sakura_OperateBitwiseOrUint32(uint32_t bit_mask, size_t num_data, uint32_t const data[],
		bool const edit_mask[], uint32_t result[]) SAKURA_B200_NOEXCEPT;
sakura_Status sakura_OperateBitwiseXorUint8(uint8_t bit_mask, size_t num_data, uint8_t const data[],
		bool const edit_mask[], uint8_t result[]) SAKURA_B200_NOEXCEPT;
sakura_Status sakura_OperateBitwiseXorUint32(uint32_t bit_mask, size_t num_data, uint32_t const data[],
		bool const edit_mask[], uint32_t result[]) SAKURA_B200_NOEXCEPT;
/** NEW. Device arrays; `operation`: 0 And, 1 ConverseNonImplication, 2 Implication, 3 Not, 4 Or, 5 Xor. */
sakura_Status sakura_OperateBitwiseUint8Device(int operation, uint8_t bit_mask, size_t num_data,
		uint8_t const d_data[], bool const d_edit_mask[], uint8_t d_result[], void *cuda_stream)
		SAKURA_B200_NOEXCEPT;
sakura_Status sakura_OperateBitwiseUint32Device(int operation, uint32_t bit_mask, size_t num_data,
		uint32_t const d_data[], bool const d_edit_mask[], uint32_t d_result[], void *cuda_stream)
		SAKURA_B200_NOEXCEPT;

/* ---- interpolation of many arrays at once ------------------------------
 * sakura.h:961-1096 / interpolation.cc. Nearest and linear interpolation (and polynomial of order
 * 0, which the reference maps to nearest) with masked base data, no extrapolation, ascending or
 * descending positions; same argument rules and results as the reference. Polynomial (order > 0)
 * and natural-spline interpolation are not implemented: kInvalidArgument, never approximated. */
typedef enum {
	sakura_InterpolationMethod_kNearest = 0,
	sakura_InterpolationMethod_kLinear = 1,
	sakura_InterpolationMethod_kPolynomial = 2,
	sakura_InterpolationMethod_kSpline = 3,
	sakura_InterpolationMethod_kNumElements = 4
} sakura_InterpolationMethod;

/** Replaces sakura_InterpolateXAxisFloat, sakura.h:1073: data laid out [num_array][num_base]. */
sakura_Status sakura_InterpolateXAxisFloat(sakura_InterpolationMethod interpolation_method,
		uint8_t polynomial_order, size_t num_base, double const base_position[/*num_base*/],
		size_t num_array, float const base_data[/*num_base*num_array*/],
		bool const base_mask[/*num_base*num_array*/], size_t num_interpolated,
		double const interpolated_position[/*num_interpolated*/],
		float interpolated_data[/*num_interpolated*num_array*/],
		bool interpolated_mask[/*num_interpolated*num_array*/]) SAKURA_B200_NOEXCEPT;
/** Replaces sakura_InterpolateYAxisFloat, sakura.h:1087: data laid out [num_base][num_array]
 *  (e.g. Tsys[time][channel] -> one row per spectrum for the calibration). */
sakura_Status sakura_InterpolateYAxisFloat(sakura_InterpolationMethod interpolation_method,
		uint8_t polynomial_order, size_t num_base, double const base_position[/*num_base*/],
		size_t num_array, float const base_data[/*num_base*num_array*/],
		bool const base_mask[/*num_base*num_array*/], size_t num_interpolated,
		double const interpolated_position[/*num_interpolated*/],
		float interpolated_data[/*num_interpolated*num_array*/],
		bool interpolated_mask[/*num_interpolated*num_array*/]) SAKURA_B200_NOEXCEPT;
/** NEW. All arrays on the device (positions too); asynchronous on `cuda_stream`. */
sakura_Status sakura_InterpolateXAxisFloatDevice(sakura_InterpolationMethod interpolation_method,
		uint8_t polynomial_order, size_t num_base, double const d_base_position[], size_t num_array,
		float const d_base_data[], bool const d_base_mask[], size_t num_interpolated,
		double const d_interpolated_position[], float d_interpolated_data[], bool d_interpolated_mask[],
		void *cuda_stream) SAKURA_B200_NOEXCEPT;
sakura_Status sakura_InterpolateYAxisFloatDevice(sakura_InterpolationMethod interpolation_method,
		uint8_t polynomial_order, size_t num_base, double const d_base_position[], size_t num_array,
		float const d_base_data[], bool const d_base_mask[], size_t num_interpolated,
		double const d_interpolated_position[], float d_interpolated_data[], bool d_interpolated_mask[],
		void *cuda_stream) SAKURA_B200_NOEXCEPT;

/* ---- NEW: library introspection (measurement support) --------------------- */

/** Number of CUDA kernels this library has launched in this process so far. */
uint64_t sakura_b200_GetKernelLaunchCount(void) SAKURA_B200_NOEXCEPT;
/** Name of the kernel variant the last fit call dispatched to ("poly_fast", "general", ...). */
char const *sakura_b200_GetLastKernelName(void) SAKURA_B200_NOEXCEPT;
/** Rows of the last polynomial fit on this context that the screening kernel ("poly_screen") did
 *  not decide itself and handed to the per-channel kernel; -1 if the last call did not screen.
 *  Synchronises the device. */
int64_t sakura_b200_GetLastHandOverCount(struct sakura_LSQFitContextFloat const *context)
		SAKURA_B200_NOEXCEPT;
/** Text of the last CUDA error seen by the library ("" if none). */
char const *sakura_b200_GetLastErrorMessage(void) SAKURA_B200_NOEXCEPT;
/** "sakura_b200 <version> (sm_100a)". */
char const *sakura_b200_GetVersionString(void) SAKURA_B200_NOEXCEPT;
/** Tuning of the host-pointer batch paths (process-wide; a value outside its range keeps the current
 *  setting): `chunk_mib` in [1, 4096] = MiB of spectra per chunk of the copy/compute pipeline
 *  (default 96, environment SAKURA_B200_CHUNK_MIB); `packed_kib` in [0, 2^20] = batches that stage
 *  less than this many KiB use one pinned arena with a single upload and download (default 2048,
 *  environment SAKURA_B200_PACKED_KIB; 0 disables that path). Results do not depend on either. */
void sakura_b200_SetHostPipelineTuning(int64_t chunk_mib, int64_t packed_kib) SAKURA_B200_NOEXCEPT;
/** The host-side codec of the pipelined paths (no GPU involved): boolean masks (one byte per channel, any
 *  non-zero byte is true, as `bool const mask[]` of sakura.h:1773-1781) to one bit per channel and back.
 *  The batch entry points move masks over PCIe in this form (a fifth of the bytes of a spectrum in either
 *  direction; SAKURA_B200_PACK_MASKS=0 turns that off); exported so that the codec can be tested and used
 *  on its own. `num_data` must be a multiple of 32; `bits` holds num_rows * num_data / 32 words, bit b of
 *  word w of a row = channel 32 w + b. Unpacking writes bytes 0 / 1 and only the rows r with
 *  (row_flags[r] & flag_bit) != 0 when `row_flags` is not NULL. */
sakura_Status sakura_b200_PackBoolMaskRows(size_t num_rows, size_t num_data, bool const mask[],
		size_t mask_stride, uint32_t bits[]) SAKURA_B200_NOEXCEPT;
sakura_Status sakura_b200_UnpackBoolMaskRows(size_t num_rows, size_t num_data, uint32_t const bits[],
		bool mask[], size_t mask_stride, int32_t const row_flags[], int32_t flag_bit) SAKURA_B200_NOEXCEPT;

#ifdef __cplusplus
}
#endif

#endif /* SAKURA_B200_H_ */
