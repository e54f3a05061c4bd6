// Batched cubic-spline baseline fit from per-piece local moments; see lsqfit_spline.cu.
#ifndef SAKURA_B200_LSQFIT_SPLINE_CUH_
#define SAKURA_B200_LSQFIT_SPLINE_CUH_

#include <stddef.h>
#include <stdint.h>
#include <cuda_runtime.h>

namespace sakura_b200 {

constexpr int kSplineThreads = 256;
constexpr int kSplineMaxPieces = 29; // npiece + 3 bases <= 32
constexpr int kSplineMaxChannels = 16384;

struct SplineFitParams {
	size_t num_spectra;
	int n; // channels (multiple of 4)
	int npiece;
	float const *data;
	size_t data_stride;
	uint8_t const *mask;
	size_t mask_stride;
	float clip_threshold_sigma;
	int num_fitting_max;
	double *coeff; // may be NULL; [num_spectra][coeff_stride], 4 per piece
	size_t coeff_stride;
	float *best_fit; // may be NULL
	float *residual; // may be NULL
	uint8_t *final_mask;
	float *rms;
	int32_t *status;
	int32_t *lsqfit_status;
	int32_t *flags; // may be NULL; bit0 results published, bit1 final_mask written, bit2 boundary written
	unsigned long long *boundary; // may be NULL; [num_spectra][npiece + 1]
	int32_t *fallback_rows; // rows whose normal equations are (nearly) singular
	int32_t *fallback_count;
};

bool SplineFastSupported(int n, int npiece);
size_t SplineSharedBytes(int n, int npiece);
cudaError_t LaunchSplineFit(SplineFitParams const &p, int grid, cudaStream_t stream);
/** Largest grid worth launching: resident CTAs per SM for this shape times the SM count. */
int SplineResidentCtas(int n, int npiece, int num_sms);

} // namespace sakura_b200

#endif
