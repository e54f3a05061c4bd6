// Batched masked statistics launcher (see statistics.cu).
#ifndef SAKURA_B200_STATISTICS_CUH_
#define SAKURA_B200_STATISTICS_CUH_

#include <stddef.h>
#include <stdint.h>
#include <cuda_runtime.h>

namespace sakura_b200 {

/** Same layout as sakura_StatisticsResultFloat (sakura.h:208-237). */
struct StatResult {
	unsigned long long count;
	double sum;
	double square_sum;
	float min;
	float max;
	long long index_of_min;
	long long index_of_max;
};

/** Bytes of device scratch LaunchStats needs for this shape (0 when rows are short). */
size_t StatsPartialBytes(size_t num_rows, size_t n, size_t *chunks_out, size_t *chunk_len_out);

cudaError_t LaunchStats(size_t num_rows, size_t n, float const *data, size_t data_stride,
		uint8_t const *mask, size_t mask_stride, StatResult *result, void *partial_ws,
		int num_sms, cudaStream_t stream, int *launches);

} // namespace sakura_b200

#endif
