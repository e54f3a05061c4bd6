// Boolean-filter builders (libsakura/src/bool_filter_collection.cc) as one streaming kernel:
// 16 elements per thread, 128-bit loads of the data and one 128-bit store of 16 result bytes.
// HBM-bound: 5 B per element for float/int32 input, 2 B for byte input (+1 B with the fused AND).
#include "bool_filters.cuh"

namespace sakura_b200 {

namespace {

// The reference's AVX float comparisons are the UNORDERED predicates (_CMP_NGT_UQ / _CMP_NGE_UQ,
// packed_operation.h:1423-1436): a NaN element is "in range" in the SIMD body (and not in the scalar
// remainder, where the product is NaN). Reproduced as the negated opposite comparison.
template<bool INCLUSIVE>
__device__ __forceinline__ bool InRangeCompare(float x, float lo, float hi) {
	return INCLUSIVE ? (!(lo > x) && !(x > hi)) : (!(lo >= x) && !(x >= hi));
}

template<bool INCLUSIVE>
__device__ __forceinline__ bool InRangeCompare(int x, int lo, int hi) {
	return INCLUSIVE ? (lo <= x && x <= hi) : (lo < x && x < hi);
}

// (x - lower) * (upper - x) >= 0 (> 0), float arithmetic as written (bool_filter_collection.cc:52)
template<bool INCLUSIVE>
__device__ __forceinline__ bool InRangeProduct(float x, float lo, float hi) {
	float const v = __fmul_rn(__fsub_rn(x, lo), __fsub_rn(hi, x));
	return INCLUSIVE ? (v >= 0.0f) : (v > 0.0f);
}

// the int form overflows for far-apart values (undefined in C++; two's-complement wrap-around here)
template<bool INCLUSIVE>
__device__ __forceinline__ bool InRangeProduct(int x, int lo, int hi) {
	unsigned const a = static_cast<unsigned>(x) - static_cast<unsigned>(lo);
	unsigned const b = static_cast<unsigned>(hi) - static_cast<unsigned>(x);
	int const v = static_cast<int>(a * b);
	return INCLUSIVE ? (v >= 0) : (v > 0);
}

__device__ __forceinline__ float AsValue(uint32_t raw, float) {
	return __uint_as_float(raw);
}

__device__ __forceinline__ int AsValue(uint32_t raw, int) {
	return static_cast<int>(raw);
}

template<typename T, bool INCLUSIVE>
__device__ __forceinline__ bool Ranges(BoolFilterParams const &p, T x, bool compare) {
	bool in_range = false;
	if (p.num_condition <= static_cast<size_t>(kFilterInlineBounds)) {
		int const nc = static_cast<int>(p.num_condition);
#pragma unroll 4
		for (int j = 0; j < nc; ++j) {
			T const lo = AsValue(p.lower[j], T());
			T const hi = AsValue(p.upper[j], T());
			in_range |= compare ? InRangeCompare<INCLUSIVE>(x, lo, hi) : InRangeProduct<INCLUSIVE>(x, lo, hi);
		}
	} else {
		// generic form (bool_filter_collection.cc:330-345): product predicate, first hit wins
		for (size_t j = 0; j < p.num_condition && !in_range; ++j) {
			T const lo = AsValue(p.bounds[j], T());
			T const hi = AsValue(p.bounds[p.num_condition + j], T());
			in_range = InRangeProduct<INCLUSIVE>(x, lo, hi);
		}
	}
	return in_range;
}

template<typename T, int OP>
__device__ __forceinline__ bool Evaluate(BoolFilterParams const &p, uint32_t raw, bool compare) {
	T const x = AsValue(raw, T());
	T const thr = AsValue(p.threshold, T());
	switch (OP) {
	case kFilterRangesInclusive:
		return Ranges<T, true>(p, x, compare);
	case kFilterRangesExclusive:
		return Ranges<T, false>(p, x, compare);
	case kFilterGreaterThan:
		return x > thr;
	case kFilterGreaterThanOrEquals:
		return x >= thr;
	case kFilterLessThan:
		return x < thr;
	case kFilterLessThanOrEquals:
		return x <= thr;
	case kFilterFiniteFloat:
		return (raw & 0x7f800000u) != 0x7f800000u;
	default: // kFilterNonZero (uint32)
		return raw != 0u;
	}
}

// 32-bit element types: 16 elements per thread
template<typename T, int OP>
__global__ void __launch_bounds__(256)
BoolFilterKernel32(BoolFilterParams const p) {
	size_t const n16 = p.num_data / 16;
	uint4 const *in = reinterpret_cast<uint4 const *>(p.data);
	uint4 *out = reinterpret_cast<uint4 *>(p.result);
	uint4 const *andw = reinterpret_cast<uint4 const *>(p.and_with);
	size_t const stride = static_cast<size_t>(gridDim.x) * blockDim.x;
	for (size_t g = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x; g < n16; g += stride) {
		uint4 v[4];
#pragma unroll
		for (int q = 0; q < 4; ++q) {
			v[q] = __ldg(in + 4 * g + q);
		}
		uint32_t w[4];
		constexpr bool kRanges = OP == kFilterRangesInclusive || OP == kFilterRangesExclusive;
		if (kRanges && 16 * g + 16 <= p.compare_end) {
			// all 16 elements in the reference's SIMD body: conditions outermost, bounds read once
			T x[16];
#pragma unroll
			for (int q = 0; q < 4; ++q) {
				x[4 * q] = AsValue(v[q].x, T());
				x[4 * q + 1] = AsValue(v[q].y, T());
				x[4 * q + 2] = AsValue(v[q].z, T());
				x[4 * q + 3] = AsValue(v[q].w, T());
			}
			uint32_t bits = 0u;
			int const nc = static_cast<int>(p.num_condition);
			for (int j = 0; j < nc; ++j) {
				T const lo = AsValue(p.lower[j], T());
				T const hi = AsValue(p.upper[j], T());
#pragma unroll
				for (int e = 0; e < 16; ++e) {
					bits |= InRangeCompare<OP == kFilterRangesInclusive>(x[e], lo, hi) ? (1u << e) : 0u;
				}
			}
#pragma unroll
			for (int q = 0; q < 4; ++q) {
				w[q] = (((bits >> (4 * q)) & 0xfu) * 0x00204081u) & 0x01010101u;
			}
		} else
#pragma unroll
		for (int q = 0; q < 4; ++q) {
			// compare_end is a multiple of 8: the four elements of a float4 are on one side of it
			bool const cmp = 16 * g + 4 * q < p.compare_end;
			w[q] = (Evaluate<T, OP>(p, v[q].x, cmp) ? 0x1u : 0u) | (Evaluate<T, OP>(p, v[q].y, cmp) ? 0x100u : 0u)
					| (Evaluate<T, OP>(p, v[q].z, cmp) ? 0x10000u : 0u)
					| (Evaluate<T, OP>(p, v[q].w, cmp) ? 0x1000000u : 0u);
		}
		uint4 r = make_uint4(w[0], w[1], w[2], w[3]);
		if (andw != nullptr) {
			uint4 const a = andw[g];
			r.x &= a.x;
			r.y &= a.y;
			r.z &= a.z;
			r.w &= a.w;
		}
		out[g] = r;
	}
	// remainder (< 16 elements) and nothing else: one thread each
	size_t const tail = n16 * 16;
	size_t const i = tail + static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x;
	if (i < p.num_data) {
		uint8_t r = Evaluate<T, OP>(p, reinterpret_cast<uint32_t const *>(p.data)[i], i < p.compare_end) ? 1 : 0;
		if (p.and_with != nullptr) {
			r &= p.and_with[i];
		}
		p.result[i] = r;
	}
}

// any alignment: one element per thread
template<typename T, int OP>
__global__ void __launch_bounds__(256)
BoolFilterKernel32Scalar(BoolFilterParams const p) {
	size_t const stride = static_cast<size_t>(gridDim.x) * blockDim.x;
	for (size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < p.num_data; i += stride) {
		uint8_t r = Evaluate<T, OP>(p, reinterpret_cast<uint32_t const *>(p.data)[i], i < p.compare_end) ? 1 : 0;
		if (p.and_with != nullptr) {
			r &= p.and_with[i];
		}
		p.result[i] = r;
	}
}

// byte element types (uint8 -> bool, bool inversion): 16 per thread when aligned
__global__ void __launch_bounds__(256)
BoolFilterKernel8(BoolFilterParams const p, bool vector) {
	size_t const stride = static_cast<size_t>(gridDim.x) * blockDim.x;
	size_t const n16 = vector ? p.num_data / 16 : 0;
	bool const invert = p.op == kFilterInvertBool;
	auto word = [invert](uint32_t v) -> uint32_t {
		// invert: byte ^ 1 (bool_filter_collection.cc:424); else byte != 0
		return invert ? (v ^ 0x01010101u) : (__vcmpne4(v, 0u) & 0x01010101u);
	};
	for (size_t g = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x; g < n16; g += stride) {
		uint4 const v = __ldg(reinterpret_cast<uint4 const *>(p.data) + g);
		uint4 r = make_uint4(word(v.x), word(v.y), word(v.z), word(v.w));
		if (p.and_with != nullptr) {
			uint4 const a = reinterpret_cast<uint4 const *>(p.and_with)[g];
			r.x &= a.x;
			r.y &= a.y;
			r.z &= a.z;
			r.w &= a.w;
		}
		reinterpret_cast<uint4 *>(p.result)[g] = r;
	}
	for (size_t i = n16 * 16 + static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < p.num_data;
			i += stride) {
		uint8_t const v = reinterpret_cast<uint8_t const *>(p.data)[i];
		uint8_t r = invert ? static_cast<uint8_t>(v ^ 1u) : static_cast<uint8_t>(v != 0);
		if (p.and_with != nullptr) {
			r &= p.and_with[i];
		}
		p.result[i] = r;
	}
}

bool Aligned16(void const *q) {
	return q == nullptr || (reinterpret_cast<uintptr_t>(q) % 16) == 0;
}

template<typename T, int OP>
void Launch32(BoolFilterParams const &p, bool vector, unsigned grid, cudaStream_t stream) {
	if (vector) {
		BoolFilterKernel32<T, OP><<<grid, 256, 0, stream>>>(p);
	} else {
		BoolFilterKernel32Scalar<T, OP><<<grid, 256, 0, stream>>>(p);
	}
}

template<typename T>
cudaError_t Dispatch32(BoolFilterParams const &p, bool vector, unsigned grid, cudaStream_t stream) {
	switch (p.op) {
	case kFilterRangesInclusive:
		Launch32<T, kFilterRangesInclusive>(p, vector, grid, stream);
		break;
	case kFilterRangesExclusive:
		Launch32<T, kFilterRangesExclusive>(p, vector, grid, stream);
		break;
	case kFilterGreaterThan:
		Launch32<T, kFilterGreaterThan>(p, vector, grid, stream);
		break;
	case kFilterGreaterThanOrEquals:
		Launch32<T, kFilterGreaterThanOrEquals>(p, vector, grid, stream);
		break;
	case kFilterLessThan:
		Launch32<T, kFilterLessThan>(p, vector, grid, stream);
		break;
	case kFilterLessThanOrEquals:
		Launch32<T, kFilterLessThanOrEquals>(p, vector, grid, stream);
		break;
	case kFilterFiniteFloat:
		Launch32<T, kFilterFiniteFloat>(p, vector, grid, stream);
		break;
	case kFilterNonZero:
		Launch32<T, kFilterNonZero>(p, vector, grid, stream);
		break;
	default:
		return cudaErrorInvalidValue;
	}
	return cudaGetLastError();
}

} // namespace

// bit_operation.cc:60-190, the expressions as written there (x derived from the edit-mask byte)
template<typename T>
__device__ __forceinline__ T BitOperate(int op, T d, T bm, uint8_t mask) {
	T const x = static_cast<T>(static_cast<T>(mask) - static_cast<T>(1));
	T const nx = static_cast<T>(~x);
	switch (op) {
	case kBitAnd:
		return static_cast<T>(d & (bm | x));
	case kBitConverseNonImplication:
		return static_cast<T>((d ^ nx) & (bm | x));
	case kBitImplication:
		return static_cast<T>((d ^ nx) | (bm & nx));
	case kBitNot:
		return static_cast<T>(d ^ nx);
	case kBitOr:
		return static_cast<T>(d | (bm & nx));
	default:
		return static_cast<T>(d ^ (bm & nx));
	}
}

template<typename T>
__global__ void __launch_bounds__(256)
BitOperationKernel(BitOperationParams const p) {
	size_t const stride = static_cast<size_t>(gridDim.x) * blockDim.x;
	T const *in = static_cast<T const *>(p.data);
	T *out = static_cast<T *>(p.result);
	T const bm = static_cast<T>(p.bit_mask);
	int const op = p.op;
	// four elements per thread when everything is 4-element aligned (uint8: one word each of data,
	// mask and result; uint32: 128-bit data accesses)
	bool const vec = (reinterpret_cast<uintptr_t>(p.data) % (4 * sizeof(T))) == 0
			&& (reinterpret_cast<uintptr_t>(p.result) % (4 * sizeof(T))) == 0
			&& (reinterpret_cast<uintptr_t>(p.edit_mask) % 4) == 0;
	size_t const n4 = vec ? p.num_data / 4 : 0;
	for (size_t g = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x; g < n4; g += stride) {
		uint32_t const m = reinterpret_cast<uint32_t const *>(p.edit_mask)[g];
		T v[4];
		if (sizeof(T) == 1) {
			uint32_t const w = reinterpret_cast<uint32_t const *>(in)[g];
			uint32_t r = 0u;
#pragma unroll
			for (int e = 0; e < 4; ++e) {
				r |= static_cast<uint32_t>(BitOperate<T>(op, static_cast<T>(w >> (8 * e)), bm,
						static_cast<uint8_t>(m >> (8 * e)))) << (8 * e);
			}
			reinterpret_cast<uint32_t *>(out)[g] = r;
		} else {
			uint4 const w = reinterpret_cast<uint4 const *>(in)[g];
			v[0] = static_cast<T>(w.x);
			v[1] = static_cast<T>(w.y);
			v[2] = static_cast<T>(w.z);
			v[3] = static_cast<T>(w.w);
#pragma unroll
			for (int e = 0; e < 4; ++e) {
				v[e] = BitOperate<T>(op, v[e], bm, static_cast<uint8_t>(m >> (8 * e)));
			}
			reinterpret_cast<uint4 *>(out)[g] = make_uint4(v[0], v[1], v[2], v[3]);
		}
	}
	for (size_t i = n4 * 4 + static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < p.num_data;
			i += stride) {
		out[i] = BitOperate<T>(op, in[i], bm, p.edit_mask[i]);
	}
}

cudaError_t LaunchBitOperation(BitOperationParams const &p, int num_sms, cudaStream_t stream) {
	if (p.num_data == 0) {
		return cudaSuccess;
	}
	size_t blocks = (p.num_data / 4 + 255) / 256 + 1;
	size_t const cap = static_cast<size_t>(num_sms) * 16;
	if (blocks > cap) {
		blocks = cap;
	}
	if (p.wide) {
		BitOperationKernel<uint32_t><<<static_cast<unsigned>(blocks), 256, 0, stream>>>(p);
	} else {
		BitOperationKernel<uint8_t><<<static_cast<unsigned>(blocks), 256, 0, stream>>>(p);
	}
	return cudaGetLastError();
}

cudaError_t LaunchBoolFilter(BoolFilterParams const &p, int num_sms, cudaStream_t stream) {
	if (p.num_data == 0) {
		return cudaSuccess;
	}
	bool const vector = Aligned16(p.data) && Aligned16(p.result) && Aligned16(p.and_with);
	size_t const per_thread = vector ? 16 : 1;
	size_t blocks = (p.num_data / per_thread + 255) / 256 + 1;
	size_t const cap = static_cast<size_t>(num_sms) * 16;
	if (blocks > cap) {
		blocks = cap;
	}
	unsigned const grid = static_cast<unsigned>(blocks);
	if (p.type == kFilterUint8) {
		BoolFilterKernel8<<<grid, 256, 0, stream>>>(p, vector);
	} else if (p.type == kFilterFloat) {
		return Dispatch32<float>(p, vector, grid, stream);
	} else { // int32 and uint32 share the integer instantiation (uint32 only uses kFilterNonZero)
		return Dispatch32<int>(p, vector, grid, stream);
	}
	return cudaGetLastError();
}

} // namespace sakura_b200
