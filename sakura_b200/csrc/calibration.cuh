// Position-switch calibration result = scaling * (data - reference) / reference
// (sakura_CalibrateDataWith{Array,Const}ScalingFloat, libsakura/src/normalization.cc:128-139,
// 206-222): float subtract, multiply, true IEEE divide in that order. SURVEY 8(f) N1.
#ifndef SAKURA_B200_CALIBRATION_CUH_
#define SAKURA_B200_CALIBRATION_CUH_

#include <stddef.h>
#include <stdint.h>
#include <cuda_runtime.h>

namespace sakura_b200 {

/** How a fit kernel (or the standalone kernel) obtains its input spectrum. Strides are in
 *  elements; a stride of 0 shares one array between all rows. */
struct CalibrationSpec {
	float const *reference; // NULL: no calibration, the data is used as is
	size_t reference_stride;
	float const *scaling; // NULL: const_scaling
	size_t scaling_stride;
	float const_scaling;
	int flag_nonfinite; // also clear the mask where the calibrated value is NaN or +-Inf
};

__device__ __forceinline__ float CalibrateOne(float s, float t, float r) {
	return __fdiv_rn(__fmul_rn(s, __fsub_rn(t, r)), r);
}

struct CalibrateParams {
	size_t num_spectra;
	size_t num_data;
	float const *data;
	size_t data_stride;
	CalibrationSpec spec;
	float *result;
	size_t result_stride;
	uint8_t const *mask; // optional: mask_out = mask && finite(result) when spec.flag_nonfinite
	size_t mask_stride;
	uint8_t *mask_out;
};

cudaError_t LaunchCalibrate(CalibrateParams const &p, int num_sms, cudaStream_t stream);

} // namespace sakura_b200

#endif
