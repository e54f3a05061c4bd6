// General (any basis, any number of bases <= kMaxBases) batched LSQ fit kernel.
// One CTA per spectrum; grid-stride over spectra.
#ifndef SAKURA_B200_LSQFIT_GENERAL_CUH_
#define SAKURA_B200_LSQFIT_GENERAL_CUH_

#include <stddef.h>
#include <stdint.h>
#include <cuda_runtime.h>

namespace sakura_b200 {

constexpr int kMaxBases = 128; // largest number of simultaneous LSQ equations supported
constexpr int kGeneralThreads = 256;

struct GeneralFitParams {
	int fit_type; // FitType
	size_t num_spectra;
	int n; // channels per spectrum
	int nb_ctx; // row stride of the context basis table
	int K; // number of fitted bases (num_coeff of DoLSQFit)
	int nb_row; // row stride of the per-row spline table (num_lsq_bases_max)
	int npiece; // spline pieces actually fitted (0 otherwise)
	double const *basis; // context table [n][nb_ctx]
	float const *data;
	size_t data_stride;
	uint8_t const *mask;
	size_t mask_stride;
	float clip_threshold_sigma;
	int num_fitting_max;
	double *coeff; // may be NULL; row stride coeff_stride
	size_t coeff_stride;
	float *best_fit; // may be NULL (row stride data_stride)
	float *residual; // may be NULL (row stride data_stride)
	uint8_t *final_mask; // never NULL
	float *rms;
	int32_t *status;
	int32_t *lsqfit_status;
	int32_t *flags; // may be NULL; bit0 results published, bit1 final_mask written, bit2 boundary written
	unsigned long long *boundary; // spline: [num_spectra][npiece+1] (size_t), else NULL
	float *ws_residual; // [gridDim.x][n] per-CTA staging of the residual
	float *ws_best_fit; // [gridDim.x][n] per-CTA staging of the model (when best_fit != NULL)
	double *ws_spline_basis; // [gridDim.x][n][nb_row] scratch (spline only)
	int16_t use_idx[kMaxBases]; // column of the context table for each fitted basis
	int32_t const *row_index; // optional: rows to fit (device), else NULL = all rows
	int32_t const *row_count; // optional: number of entries of row_index (device)
};

size_t GeneralFitSharedBytes(int K, int npiece, int threads);

/** Launches the kernel on `stream` with `grid` CTAs. Returns the CUDA error code. */
cudaError_t LaunchGeneralFit(GeneralFitParams const &p, int grid, cudaStream_t stream);

struct SubtractParams {
	int fit_type;
	size_t num_spectra;
	int n;
	int nb_ctx;
	int K;
	int npiece;
	double const *basis;
	float const *data;
	size_t data_stride;
	double const *coeff; // device, already multiplied by (n-1)^k where the reference does
	size_t coeff_stride;
	unsigned long long const *boundary; // spline: device [num_spectra][npiece+1]
	float *out;
	int16_t use_idx[kMaxBases];
};
cudaError_t LaunchSubtract(SubtractParams const &p, int grid, cudaStream_t stream);

} // namespace sakura_b200

#endif
