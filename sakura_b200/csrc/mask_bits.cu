// See mask_bits.cuh: streaming kernels, one 32-channel word per thread (4 B <-> 32 B), HBM-bound and
// small against the fit (4.5 of the 41 KB a 4096-channel spectrum moves on the device).
#include "mask_bits.cuh"

namespace sakura_b200 {

namespace {

__global__ void __launch_bounds__(256)
MaskBitsToBytesKernel(uint32_t const *__restrict__ bits, size_t words, size_t total, uint8_t *__restrict__ bytes,
		size_t byte_stride) {
	size_t const idx = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x;
	if (idx >= total) {
		return;
	}
	size_t const r = idx / words;
	size_t const w = idx - r * words;
	uint32_t const b = bits[idx];
	uint32_t o[8];
#pragma unroll
	for (int q = 0; q < 8; ++q) {
		o[q] = (((b >> (4 * q)) & 0xfu) * 0x00204081u) & 0x01010101u; // nibble -> four 0 / 1 bytes
	}
	uint4 *dst = reinterpret_cast<uint4 *>(bytes + r * byte_stride + 32 * w);
	dst[0] = make_uint4(o[0], o[1], o[2], o[3]);
	dst[1] = make_uint4(o[4], o[5], o[6], o[7]);
}

__device__ __forceinline__ uint32_t Bits16(uint4 const m) {
	auto nib = [](uint32_t w) {
		// bit 7 of every non-zero byte, then the four flags gathered into the top nibble by a product
		uint32_t const t = (((w & 0x7f7f7f7fu) + 0x7f7f7f7fu) | w) & 0x80808080u;
		return (t * 0x00204081u) >> 28;
	};
	return nib(m.x) | (nib(m.y) << 4) | (nib(m.z) << 8) | (nib(m.w) << 12);
}

__global__ void __launch_bounds__(256)
MaskBytesToBitsKernel(uint8_t const *__restrict__ bytes, size_t byte_stride, size_t words, size_t total,
		uint32_t *__restrict__ bits) {
	size_t const idx = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x;
	if (idx >= total) {
		return;
	}
	size_t const r = idx / words;
	size_t const w = idx - r * words;
	uint4 const *src = reinterpret_cast<uint4 const *>(bytes + r * byte_stride + 32 * w);
	bits[idx] = Bits16(src[0]) | (Bits16(src[1]) << 16);
}

} // namespace

cudaError_t LaunchMaskBitsToBytes(uint32_t const *bits, size_t words, size_t rows, uint8_t *bytes,
		size_t byte_stride, cudaStream_t stream) {
	size_t const total = words * rows;
	if (total == 0) {
		return cudaSuccess;
	}
	MaskBitsToBytesKernel<<<static_cast<unsigned>((total + 255) / 256), 256, 0, stream>>>(bits, words, total,
			bytes, byte_stride);
	return cudaGetLastError();
}

cudaError_t LaunchMaskBytesToBits(uint8_t const *bytes, size_t byte_stride, size_t words, size_t rows,
		uint32_t *bits, cudaStream_t stream) {
	size_t const total = words * rows;
	if (total == 0) {
		return cudaSuccess;
	}
	MaskBytesToBitsKernel<<<static_cast<unsigned>((total + 255) / 256), 256, 0, stream>>>(bytes, byte_stride,
			words, total, bits);
	return cudaGetLastError();
}

} // namespace sakura_b200
