// Batched sinusoid baseline fit with FP64 tensor-core (DMMA) contractions; see lsqfit_sinusoid.cu.
#ifndef SAKURA_B200_LSQFIT_SINUSOID_CUH_
#define SAKURA_B200_LSQFIT_SINUSOID_CUH_

#include <stddef.h>
#include <stdint.h>
#include <cuda_runtime.h>

namespace sakura_b200 {

constexpr int kSinRowsPerTile = 32; // spectra fitted together by one CTA (4 DMMA row tiles)
constexpr int kSinMaxKp = 48; // fitted bases, padded to a multiple of 8
constexpr int kSinThreads = 512;

struct SinusoidFitParams {
	size_t num_spectra;
	int n; // channels (multiple of 8)
	int K; // fitted bases
	int Kp; // K rounded up to a multiple of 8
	int next; // columns of the extended trigonometric table: 2*mmax + 1
	double const *table; // [n][Kp] selected basis columns (zero padded), built per call
	int ext_ld; // row stride of ext: SinusoidExtLd(next), the padding columns are zero
	double const *ext; // [n][ext_ld]: 1, then (sin, cos)(m theta_i) for m = 1..mmax
	// basis k is: kind[k] = 0 constant, 1 sin, 2 cos; wave[k] its wave number
	int8_t kind[kSinMaxKp];
	int16_t wave[kSinMaxKp];
	float const *data;
	size_t data_stride;
	uint8_t const *mask;
	size_t mask_stride;
	float clip_threshold_sigma;
	int num_fitting_max;
	double *coeff; // may be NULL, [num_spectra][K]
	float *best_fit; // may be NULL
	float *residual; // may be NULL
	uint8_t *final_mask;
	float *rms;
	int32_t *status;
	int32_t *lsqfit_status;
	int32_t *flags; // may be NULL
	float *ws_residual; // [gridDim.x][32][n]
	float *ws_best_fit; // [gridDim.x][32][n] (when best_fit != NULL)
	uint8_t *ws_mask; // [gridDim.x][32][n] working copy of the mask
	int32_t *fallback_rows; // rows whose normal equations are (nearly) singular
	int32_t *fallback_count;
};

size_t SinusoidSharedBytes(int Kp, int next);
int SinusoidExtLd(int next);

/** Gathers the selected columns of the context table into the compact [n][Kp] table. */
cudaError_t LaunchGatherTable(double const *ctx_table, int nb_ctx, int n, int K, int Kp,
		int16_t const *use_idx_host, double *table, cudaStream_t stream);

cudaError_t LaunchSinusoidFit(SinusoidFitParams const &p, int grid, cudaStream_t stream);

} // namespace sakura_b200

#endif
