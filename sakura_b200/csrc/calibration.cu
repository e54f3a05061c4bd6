// Standalone batched calibration kernel (HBM-bound streaming: 12 B in, 4 B out per channel with
// an array scaling factor). The fit kernels apply the same expression while loading a row
// (lsqfit_poly.cu), which saves this kernel's write and the fit's re-read of the calibrated data.
#include "calibration.cuh"

namespace sakura_b200 {

namespace {

__global__ void __launch_bounds__(256)
CalibrateKernel(CalibrateParams const p) {
	size_t const n = p.num_data;
	for (size_t row = blockIdx.y; row < p.num_spectra; row += gridDim.y) {
		float const *t = p.data + row * p.data_stride;
		float const *r = p.spec.reference + row * p.spec.reference_stride;
		float const *s = p.spec.scaling ? p.spec.scaling + row * p.spec.scaling_stride : nullptr;
		float *out = p.result + row * p.result_stride;
		uint8_t const *m = p.mask_out ? p.mask + row * p.mask_stride : nullptr;
		uint8_t *mo = p.mask_out ? p.mask_out + row * p.mask_stride : nullptr;
		for (size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < n;
				i += static_cast<size_t>(gridDim.x) * blockDim.x) {
			float const v = CalibrateOne(s ? s[i] : p.spec.const_scaling, t[i], r[i]);
			out[i] = v;
			if (mo != nullptr) {
				bool const finite = (__float_as_uint(v) & 0x7f800000u) != 0x7f800000u;
				mo[i] = (m[i] != 0 && (finite || !p.spec.flag_nonfinite)) ? 1 : 0;
			}
		}
	}
}

// four channels per thread with 128-bit accesses; MASK: mask_out = mask && finite(result) (32-bit accesses)
template<bool MASK>
__global__ void __launch_bounds__(256)
CalibrateKernel4(CalibrateParams const p) {
	size_t const n4 = p.num_data / 4;
	for (size_t row = blockIdx.y; row < p.num_spectra; row += gridDim.y) {
		float4 const *t = reinterpret_cast<float4 const *>(p.data + row * p.data_stride);
		float4 const *r = reinterpret_cast<float4 const *>(p.spec.reference + row * p.spec.reference_stride);
		float4 const *s = p.spec.scaling ?
				reinterpret_cast<float4 const *>(p.spec.scaling + row * p.spec.scaling_stride) : nullptr;
		float4 *out = reinterpret_cast<float4 *>(p.result + row * p.result_stride);
		uint32_t const *m = MASK ? reinterpret_cast<uint32_t const *>(p.mask + row * p.mask_stride) : nullptr;
		uint32_t *mo = MASK ? reinterpret_cast<uint32_t *>(p.mask_out + row * p.mask_stride) : nullptr;
		for (size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < n4;
				i += static_cast<size_t>(gridDim.x) * blockDim.x) {
			float4 const tv = t[i];
			float4 const rv = r[i];
			float4 sv = make_float4(p.spec.const_scaling, p.spec.const_scaling, p.spec.const_scaling,
					p.spec.const_scaling);
			if (s != nullptr) {
				sv = s[i];
			}
			float4 const v = make_float4(CalibrateOne(sv.x, tv.x, rv.x), CalibrateOne(sv.y, tv.y, rv.y),
					CalibrateOne(sv.z, tv.z, rv.z), CalibrateOne(sv.w, tv.w, rv.w));
			out[i] = v;
			if (MASK) {
				uint32_t w = __vcmpne4(m[i], 0u) & 0x01010101u; // any non-zero byte is "true"
				if (p.spec.flag_nonfinite) {
					auto finite = [](float x) {
						return (__float_as_uint(x) & 0x7f800000u) != 0x7f800000u;
					};
					w &= (finite(v.x) ? 0x1u : 0u) | (finite(v.y) ? 0x100u : 0u) | (finite(v.z) ? 0x10000u : 0u)
							| (finite(v.w) ? 0x1000000u : 0u);
				}
				mo[i] = w;
			}
		}
	}
}

bool Aligned16(void const *ptr, size_t stride_elems) {
	return (reinterpret_cast<uintptr_t>(ptr) % 16) == 0 && stride_elems % 4 == 0;
}

} // namespace

cudaError_t LaunchCalibrate(CalibrateParams const &p, int num_sms, cudaStream_t stream) {
	if (p.num_spectra == 0 || p.num_data == 0) {
		return cudaSuccess;
	}
	bool const mask4 = p.mask_out == nullptr
			|| ((reinterpret_cast<uintptr_t>(p.mask) % 4) == 0 && (reinterpret_cast<uintptr_t>(p.mask_out) % 4) == 0
					&& p.mask_stride % 4 == 0);
	if (mask4 && p.num_data % 4 == 0 && Aligned16(p.data, p.data_stride)
			&& Aligned16(p.spec.reference, p.spec.reference_stride)
			&& Aligned16(p.spec.scaling, p.spec.scaling_stride) && Aligned16(p.result, p.result_stride)) {
		unsigned bx4 = static_cast<unsigned>((p.num_data / 4 + 255) / 256);
		if (bx4 > 16u) {
			bx4 = 16u;
		}
		size_t by4 = static_cast<size_t>(num_sms) * 16 / bx4 + 1;
		if (by4 > p.num_spectra) {
			by4 = p.num_spectra;
		}
		if (by4 > 65535) {
			by4 = 65535;
		}
		if (p.mask_out != nullptr) {
			CalibrateKernel4<true><<<dim3(bx4, static_cast<unsigned>(by4)), 256, 0, stream>>>(p);
		} else {
			CalibrateKernel4<false><<<dim3(bx4, static_cast<unsigned>(by4)), 256, 0, stream>>>(p);
		}
		return cudaGetLastError();
	}
	unsigned bx = static_cast<unsigned>((p.num_data + 255) / 256);
	if (bx > 64u) {
		bx = 64u;
	}
	size_t by = static_cast<size_t>(num_sms) * 8 / bx + 1;
	if (by > p.num_spectra) {
		by = p.num_spectra;
	}
	if (by > 65535) {
		by = 65535;
	}
	CalibrateKernel<<<dim3(bx, static_cast<unsigned>(by)), 256, 0, stream>>>(p);
	return cudaGetLastError();
}

} // namespace sakura_b200
