// Host side of the bit-packed mask transfer of the pipelined host path (capi_fit.cu): the boolean masks
// of the C ABI are one byte per channel; over PCIe they travel as one bit per channel (a fifth of the
// bytes a spectrum moves in either direction), packed / unpacked here by a small pool of worker threads
// while the copy engines move the previous chunk.
#ifndef SAKURA_B200_HOST_BITS_H_
#define SAKURA_B200_HOST_BITS_H_

#include <stddef.h>
#include <stdint.h>
#include <functional>

namespace sakura_b200 {

/** bits[r * n / 32 + w] bit b = (bytes[r * byte_stride + 32 w + b] != 0); n % 32 == 0. */
void PackMaskRows(uint8_t const *bytes, size_t byte_stride, size_t n, size_t rows, uint32_t *bits);

/** The inverse (bytes 0 / 1). `row_flags` (may be NULL) selects the rows to write: those with
 *  (row_flags[r] & flag_bit) != 0; the others keep the caller's bytes. */
void UnpackMaskRows(uint32_t const *bits, size_t n, size_t rows, uint8_t *bytes, size_t byte_stride,
		int32_t const *row_flags, int32_t flag_bit);

/** Runs fn(begin, end) over [0, count) split between the calling thread and the pool's workers. */
void HostParallelFor(size_t count, size_t min_per_task, std::function<void(size_t, size_t)> const &fn);

} // namespace sakura_b200

#endif
