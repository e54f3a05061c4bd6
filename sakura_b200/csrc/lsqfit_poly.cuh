// Register-resident polynomial LSQ-fit kernel (the north-star hot path); see lsqfit_poly.cu.
#ifndef SAKURA_B200_LSQFIT_POLY_CUH_
#define SAKURA_B200_LSQFIT_POLY_CUH_

#include <stddef.h>
#include <stdint.h>
#include <cuda_runtime.h>

#include "calibration.cuh"

namespace sakura_b200 {

struct PolyFitParams {
	size_t num_spectra;
	int n; // channels per spectrum
	int K; // fitted bases = order + 1
	float const *data;
	size_t data_stride;
	uint8_t const *mask;
	size_t mask_stride;
	float clip_threshold_sigma;
	int num_fitting_max;
	double *coeff; // may be NULL; [num_spectra][K]
	float *best_fit; // may be NULL
	float *residual; // may be NULL
	uint8_t *final_mask;
	float *rms;
	int32_t *status;
	int32_t *lsqfit_status;
	int32_t *flags; // may be NULL; bit0 results published, bit1 final_mask written
	// optional position-switch calibration (and NaN/Inf flagging) applied while the row is loaded:
	// the fit then sees scaling * (data - reference) / reference (SURVEY 8(f) N1 fused in front)
	CalibrationSpec cal;
	int chebyshev; // report the coefficients in the Chebyshev basis T_j(2 i/(n-1) - 1)
	// PolyFitKernel only: fit the rows row_index[0 .. *row_count) instead of all rows (device
	// pointers; the list is produced by the screening kernel)
	int32_t const *row_index;
	int32_t const *row_count;
	// rows the screening kernel does not decide: LaunchPolyFit takes a buffer of num_spectra + 1
	// int32 in fallback_rows (NULL: no screening) and splits it into count and list itself
	int32_t *fallback_rows;
	int32_t *fallback_count;
};

/** True when the shape/alignment can be served by the register-resident kernel. */
bool PolyFastSupported(size_t n, size_t K, size_t data_stride, size_t mask_stride, void const *data,
		void const *mask, void const *best_fit, void const *residual, void const *final_mask);

/** True when the kernel variant that calibrates while loading exists for this channel count
 *  (4096 and 8192: every thread owns all of its 32 channel slots). */
bool PolyFusedCalibrationSupported(size_t n);

/** True when a separate calibration pass followed by the plain fit is the faster way for this shape
 *  (the warp-per-spectrum kernel would run the fit; SAKURA_B200_CAL_FUSED=1 keeps the fused variants). */
bool PolyPrefersCalibrationPrePass(size_t n, int num_fitting_max);

cudaError_t LaunchPolyFit(PolyFitParams const &p, int num_sms, cudaStream_t stream, int *launches,
		char const **kernel_name);

/** Screened clip iterations (lsqfit_poly_screen.cu): n = 4096 or 8192 channels, K <= 8. */
bool PolyScreenSupported(PolyFitParams const &p);
cudaError_t LaunchPolyScreenFit(PolyFitParams const &p, int num_sms, cudaStream_t stream);

/** One warp per spectrum, no block barrier (lsqfit_poly_warp.cu): n = 2048 or 4096 channels, K <= 8. */
bool PolyWarpSupported(PolyFitParams const &p);
cudaError_t LaunchPolyWarpFit(PolyFitParams const &p, int num_sms, cudaStream_t stream);

} // namespace sakura_b200

#endif
