// Device side of the bit-packed mask transfer (host_bits.h): bits <-> the one-byte-per-channel masks the
// fit kernels read and write.
#ifndef SAKURA_B200_MASK_BITS_CUH_
#define SAKURA_B200_MASK_BITS_CUH_

#include <stddef.h>
#include <stdint.h>
#include <cuda_runtime.h>

namespace sakura_b200 {

/** bytes[r * byte_stride + 32 w + b] = bit b of bits[r * words + w] (0 / 1); byte rows 16-byte aligned. */
cudaError_t LaunchMaskBitsToBytes(uint32_t const *bits, size_t words, size_t rows, uint8_t *bytes,
		size_t byte_stride, cudaStream_t stream);
/** The inverse: any non-zero byte is "true". */
cudaError_t LaunchMaskBytesToBits(uint8_t const *bytes, size_t byte_stride, size_t words, size_t rows,
		uint32_t *bits, cudaStream_t stream);

} // namespace sakura_b200

#endif
