// One-dimensional interpolation of many arrays at once (sakura_InterpolateXAxisFloat /
// sakura_InterpolateYAxisFloat, libsakura/src/interpolation.cc): nearest and linear, masked base
// data, no extrapolation. In the e2e chain this turns the few Tsys / sky measurements of a scan
// into per-spectrum rows for the calibration (bench/e2eReduce/src/main.cc:86-104). SURVEY 8(f) N4.
#ifndef SAKURA_B200_INTERPOLATION_CUH_
#define SAKURA_B200_INTERPOLATION_CUH_

#include <stddef.h>
#include <stdint.h>
#include <cuda_runtime.h>

namespace sakura_b200 {

struct InterpolateParams {
	int y_axis; // 1: layout [num_base][num_array] (YAxis), 0: [num_array][num_base] (XAxis)
	int linear; // 1: linear, 0: nearest
	size_t num_base;
	size_t num_array;
	size_t num_interpolated;
	double const *base_position; // sorted, ascending or descending, no duplicates
	float const *base_data;
	uint8_t const *base_mask;
	double const *interpolated_position; // any order: every point is interpolated on its own
	float *interpolated_data;
	uint8_t *interpolated_mask;
};

cudaError_t LaunchInterpolate(InterpolateParams const &p, int num_sms, cudaStream_t stream);

} // namespace sakura_b200

#endif
