// Launchers of the standalone normal-equation building blocks (lsq_blocks.cu).
#ifndef SAKURA_B200_LSQ_BLOCKS_CUH_
#define SAKURA_B200_LSQ_BLOCKS_CUH_

#include <stdint.h>
#include <cuda_runtime.h>

namespace sakura_b200 {

cudaError_t LaunchNormalEquations(int n, int K, double const *table, int stride, int const *use_idx,
		uint8_t const *mask, float const *data, double *G, double *b, int *num_unmasked,
		cudaStream_t stream);
cudaError_t LaunchUpdateNormalEquations(int K, double const *table, int stride, int const *use_idx,
		uint8_t const *mask, float const *data, int num_exclude, unsigned long long const *exclude,
		double *G, double *b, cudaStream_t stream);
cudaError_t LaunchSolve(int K, double *A, double *bvec, double *x, int *tr, cudaStream_t stream);

} // namespace sakura_b200

#endif
