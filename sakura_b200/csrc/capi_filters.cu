// C ABI of libsakura_b200.so, part 3: the callers either side of the fit: mask
// builders, bit operations, interpolation, calibration.
#include "capi_internal.cuh"

using namespace sakura_b200;
using namespace sakura_b200::capi;

// ---------------------------------------------------------------- boolean filters
namespace {

sakura_Status LaunchFilter(BoolFilterParams &p, size_t num_condition, void const *lower,
		void const *upper, DeviceBuffer *bounds_buf, cudaStream_t stream) {
	DeviceInfo info;
	sakura_Status const st = CurrentDevice(&info);
	if (st != sakura_Status_kOK) {
		return st;
	}
	p.num_condition = num_condition;
	// elements the reference's 8-wide SIMD body covers; the scalar remainder (and the generic
	// form for more than 16 conditions) uses the product predicate (bool_filter_collection.cc)
	p.compare_end = (num_condition <= 16) ? (p.num_data / 8) * 8 : 0;
	if (p.op == kFilterRangesInclusive || p.op == kFilterRangesExclusive) {
		if (num_condition <= static_cast<size_t>(kFilterInlineBounds)) {
			memcpy(p.lower, lower, num_condition * 4);
			memcpy(p.upper, upper, num_condition * 4);
		} else {
			CUDA_TRY(bounds_buf->Reserve(2 * num_condition * 4));
			CUDA_TRY(cudaMemcpyAsync(bounds_buf->ptr, lower, num_condition * 4, cudaMemcpyHostToDevice,
					stream));
			CUDA_TRY(cudaMemcpyAsync(static_cast<char *>(bounds_buf->ptr) + num_condition * 4, upper,
					num_condition * 4, cudaMemcpyHostToDevice, stream));
			// the bound arrays are the caller's: they must be consumed before this call returns
			CUDA_TRY(cudaStreamSynchronize(stream));
			p.bounds = bounds_buf->As<uint32_t>();
		}
	}
	cudaError_t const e = LaunchBoolFilter(p, info.num_sms, stream);
	g_launch_count += 1;
	g_last_kernel = "bool_filter";
	if (e != cudaSuccess) {
		SetError("LaunchBoolFilter", e);
		return sakura_Status_kUnknownError;
	}
	return sakura_Status_kOK;
}

bool ValidHostArray(void const *q) { // bool_filter_collection.cc:535-543
	return q != nullptr && sakura_IsAligned(q);
}

// host arrays: stage through the device (result may alias data: InvertBool in place)
sakura_Status FilterHost(int op, int type, size_t elem_size, size_t num_data, void const *data,
		size_t num_condition, void const *lower, void const *upper, uint32_t threshold, bool *result) {
	CHECK_ARGS(ValidHostArray(data) && ValidHostArray(result)); // :545-551
	if (op == kFilterRangesInclusive || op == kFilterRangesExclusive) {
		CHECK_ARGS(ValidHostArray(lower) && ValidHostArray(upper)); // :555-572, also for 0 conditions
	}
	if (num_data == 0) {
		return sakura_Status_kOK;
	}
	static thread_local DeviceBuffer d_in, d_out, d_bounds;
	CUDA_TRY(d_in.Reserve(num_data * elem_size));
	CUDA_TRY(d_out.Reserve(num_data));
	cudaStream_t const stream = 0;
	CUDA_TRY(cudaMemcpyAsync(d_in.ptr, data, num_data * elem_size, cudaMemcpyHostToDevice, stream));
	BoolFilterParams p;
	memset(&p, 0, sizeof(p));
	p.op = op;
	p.type = type;
	p.num_data = num_data;
	p.data = d_in.ptr;
	p.result = d_out.As<uint8_t>();
	p.threshold = threshold;
	sakura_Status const st = LaunchFilter(p, num_condition, lower, upper, &d_bounds, stream);
	if (st != sakura_Status_kOK) {
		return st;
	}
	CUDA_TRY(cudaMemcpyAsync(result, d_out.ptr, num_data, cudaMemcpyDeviceToHost, stream));
	CUDA_TRY(cudaStreamSynchronize(stream));
	return sakura_Status_kOK;
}

sakura_Status FilterDevice(int op, int type, size_t num_data, void const *d_data, size_t num_condition,
		void const *lower, void const *upper, uint32_t threshold, bool const *d_and_with, bool *d_result,
		void *cuda_stream) {
	CHECK_ARGS(d_data != nullptr && d_result != nullptr);
	if (op == kFilterRangesInclusive || op == kFilterRangesExclusive) {
		CHECK_ARGS(num_condition == 0 || (lower != nullptr && upper != nullptr));
	}
	if (num_data == 0) {
		return sakura_Status_kOK;
	}
	static thread_local DeviceBuffer d_bounds;
	BoolFilterParams p;
	memset(&p, 0, sizeof(p));
	p.op = op;
	p.type = type;
	p.num_data = num_data;
	p.data = d_data;
	p.result = reinterpret_cast<uint8_t *>(d_result);
	p.and_with = reinterpret_cast<uint8_t const *>(d_and_with);
	p.threshold = threshold;
	return LaunchFilter(p, num_condition, lower, upper, &d_bounds, static_cast<cudaStream_t>(cuda_stream));
}

uint32_t RawBits(float v) {
	uint32_t r;
	memcpy(&r, &v, 4);
	return r;
}

uint32_t RawBits(int v) {
	return static_cast<uint32_t>(v);
}

} // namespace

#define SAKURA_B200_RANGES(NAME, TYPE, TYPEID, OP) \
sakura_Status NAME(size_t num_data, TYPE const data[], size_t num_condition, TYPE const lower_bounds[], \
		TYPE const upper_bounds[], bool result[]) noexcept { \
	return FilterHost(OP, TYPEID, sizeof(TYPE), num_data, data, num_condition, lower_bounds, upper_bounds, \
			0u, result); \
}
SAKURA_B200_RANGES(sakura_SetTrueIfInRangesInclusiveFloat, float, kFilterFloat, kFilterRangesInclusive)
SAKURA_B200_RANGES(sakura_SetTrueIfInRangesInclusiveInt, int, kFilterInt32, kFilterRangesInclusive)
SAKURA_B200_RANGES(sakura_SetTrueIfInRangesExclusiveFloat, float, kFilterFloat, kFilterRangesExclusive)
SAKURA_B200_RANGES(sakura_SetTrueIfInRangesExclusiveInt, int, kFilterInt32, kFilterRangesExclusive)
#undef SAKURA_B200_RANGES

#define SAKURA_B200_COMPARE(NAME, TYPE, TYPEID, OP) \
sakura_Status NAME(size_t num_data, TYPE const data[], TYPE threshold, bool result[]) noexcept { \
	return FilterHost(OP, TYPEID, sizeof(TYPE), num_data, data, 0, nullptr, nullptr, RawBits(threshold), \
			result); \
}
SAKURA_B200_COMPARE(sakura_SetTrueIfGreaterThanFloat, float, kFilterFloat, kFilterGreaterThan)
SAKURA_B200_COMPARE(sakura_SetTrueIfGreaterThanInt, int, kFilterInt32, kFilterGreaterThan)
SAKURA_B200_COMPARE(sakura_SetTrueIfGreaterThanOrEqualsFloat, float, kFilterFloat, kFilterGreaterThanOrEquals)
SAKURA_B200_COMPARE(sakura_SetTrueIfGreaterThanOrEqualsInt, int, kFilterInt32, kFilterGreaterThanOrEquals)
SAKURA_B200_COMPARE(sakura_SetTrueIfLessThanFloat, float, kFilterFloat, kFilterLessThan)
SAKURA_B200_COMPARE(sakura_SetTrueIfLessThanInt, int, kFilterInt32, kFilterLessThan)
SAKURA_B200_COMPARE(sakura_SetTrueIfLessThanOrEqualsFloat, float, kFilterFloat, kFilterLessThanOrEquals)
SAKURA_B200_COMPARE(sakura_SetTrueIfLessThanOrEqualsInt, int, kFilterInt32, kFilterLessThanOrEquals)
#undef SAKURA_B200_COMPARE

sakura_Status sakura_SetFalseIfNanOrInfFloat(size_t num_data, float const data[], bool result[]) noexcept {
	return FilterHost(kFilterFiniteFloat, kFilterFloat, 4, num_data, data, 0, nullptr, nullptr, 0u, result);
}

sakura_Status sakura_Uint8ToBool(size_t num_data, uint8_t const data[], bool result[]) noexcept {
	return FilterHost(kFilterNonZero, kFilterUint8, 1, num_data, data, 0, nullptr, nullptr, 0u, result);
}

sakura_Status sakura_Uint32ToBool(size_t num_data, uint32_t const data[], bool result[]) noexcept {
	return FilterHost(kFilterNonZero, kFilterUint32, 4, num_data, data, 0, nullptr, nullptr, 0u, result);
}

sakura_Status sakura_InvertBool(size_t num_data, bool const data[], bool result[]) noexcept {
	return FilterHost(kFilterInvertBool, kFilterUint8, 1, num_data, data, 0, nullptr, nullptr, 0u, result);
}

sakura_Status sakura_SetTrueIfInRangesFloatDevice(bool inclusive, size_t num_data, float const d_data[],
		size_t num_condition, float const lower_bounds[], float const upper_bounds[],
		bool const d_and_with[], bool d_result[], void *cuda_stream) noexcept {
	return FilterDevice(inclusive ? kFilterRangesInclusive : kFilterRangesExclusive, kFilterFloat, num_data,
			d_data, num_condition, lower_bounds, upper_bounds, 0u, d_and_with, d_result, cuda_stream);
}

sakura_Status sakura_SetTrueIfInRangesIntDevice(bool inclusive, size_t num_data, int const d_data[],
		size_t num_condition, int const lower_bounds[], int const upper_bounds[],
		bool const d_and_with[], bool d_result[], void *cuda_stream) noexcept {
	return FilterDevice(inclusive ? kFilterRangesInclusive : kFilterRangesExclusive, kFilterInt32, num_data,
			d_data, num_condition, lower_bounds, upper_bounds, 0u, d_and_with, d_result, cuda_stream);
}

sakura_Status sakura_SetTrueIfComparesFloatDevice(int comparison, size_t num_data, float const d_data[],
		float threshold, bool const d_and_with[], bool d_result[], void *cuda_stream) noexcept {
	CHECK_ARGS(comparison >= 0 && comparison <= 3);
	return FilterDevice(kFilterGreaterThan + comparison, kFilterFloat, num_data, d_data, 0, nullptr, nullptr,
			RawBits(threshold), d_and_with, d_result, cuda_stream);
}

sakura_Status sakura_SetTrueIfComparesIntDevice(int comparison, size_t num_data, int const d_data[],
		int threshold, bool const d_and_with[], bool d_result[], void *cuda_stream) noexcept {
	CHECK_ARGS(comparison >= 0 && comparison <= 3);
	return FilterDevice(kFilterGreaterThan + comparison, kFilterInt32, num_data, d_data, 0, nullptr, nullptr,
			RawBits(threshold), d_and_with, d_result, cuda_stream);
}

sakura_Status sakura_SetFalseIfNanOrInfFloatDevice(size_t num_data, float const d_data[],
		bool const d_and_with[], bool d_result[], void *cuda_stream) noexcept {
	return FilterDevice(kFilterFiniteFloat, kFilterFloat, num_data, d_data, 0, nullptr, nullptr, 0u,
			d_and_with, d_result, cuda_stream);
}

sakura_Status sakura_Uint8ToBoolDevice(size_t num_data, uint8_t const d_data[], bool const d_and_with[],
		bool d_result[], void *cuda_stream) noexcept {
	return FilterDevice(kFilterNonZero, kFilterUint8, num_data, d_data, 0, nullptr, nullptr, 0u, d_and_with,
			d_result, cuda_stream);
}

sakura_Status sakura_Uint32ToBoolDevice(size_t num_data, uint32_t const d_data[],
		bool const d_and_with[], bool d_result[], void *cuda_stream) noexcept {
	return FilterDevice(kFilterNonZero, kFilterUint32, num_data, d_data, 0, nullptr, nullptr, 0u,
			d_and_with, d_result, cuda_stream);
}

sakura_Status sakura_InvertBoolDevice(size_t num_data, bool const d_data[], bool const d_and_with[],
		bool d_result[], void *cuda_stream) noexcept {
	return FilterDevice(kFilterInvertBool, kFilterUint8, num_data, d_data, 0, nullptr, nullptr, 0u,
			d_and_with, d_result, cuda_stream);
}

// ---------------------------------------------------------------- bit operations (bit_operation.cc)
namespace {

sakura_Status BitOperationHost(int op, int wide, uint32_t bit_mask, size_t num_data, void const *data,
		bool const *edit_mask, void *result) {
	// bit_operation.cc:41-52: NULL or unaligned data, result or edit_mask
	CHECK_ARGS(data != nullptr && result != nullptr && edit_mask != nullptr);
	CHECK_ARGS(sakura_IsAligned(data) && sakura_IsAligned(result) && sakura_IsAligned(edit_mask));
	if (num_data == 0) {
		return sakura_Status_kOK;
	}
	DeviceInfo info;
	sakura_Status const st = CurrentDevice(&info);
	if (st != sakura_Status_kOK) {
		return st;
	}
	size_t const bytes = num_data * (wide ? 4 : 1);
	static thread_local DeviceBuffer d_in, d_mask, d_out;
	CUDA_TRY(d_in.Reserve(bytes));
	CUDA_TRY(d_mask.Reserve(num_data));
	CUDA_TRY(d_out.Reserve(bytes));
	cudaStream_t const stream = 0;
	CUDA_TRY(cudaMemcpyAsync(d_in.ptr, data, bytes, cudaMemcpyHostToDevice, stream));
	CUDA_TRY(cudaMemcpyAsync(d_mask.ptr, edit_mask, num_data, cudaMemcpyHostToDevice, stream));
	BitOperationParams p;
	memset(&p, 0, sizeof(p));
	p.op = op;
	p.wide = wide;
	p.bit_mask = bit_mask;
	p.num_data = num_data;
	p.data = d_in.ptr;
	p.edit_mask = d_mask.As<uint8_t>();
	p.result = d_out.ptr;
	cudaError_t const e = LaunchBitOperation(p, info.num_sms, stream);
	g_launch_count += 1;
	g_last_kernel = "bit_operation";
	if (e != cudaSuccess) {
		SetError("LaunchBitOperation", e);
		return sakura_Status_kUnknownError;
	}
	CUDA_TRY(cudaMemcpyAsync(result, d_out.ptr, bytes, cudaMemcpyDeviceToHost, stream)); // result may alias data
	CUDA_TRY(cudaStreamSynchronize(stream));
	return sakura_Status_kOK;
}

sakura_Status BitOperationDevice(int op, int wide, uint32_t bit_mask, size_t num_data, void const *d_data,
		bool const *d_edit_mask, void *d_result, void *cuda_stream) {
	CHECK_ARGS(op >= kBitAnd && op <= kBitXor);
	CHECK_ARGS(d_data != nullptr && d_result != nullptr && d_edit_mask != nullptr);
	if (num_data == 0) {
		return sakura_Status_kOK;
	}
	DeviceInfo info;
	sakura_Status const st = CurrentDevice(&info);
	if (st != sakura_Status_kOK) {
		return st;
	}
	BitOperationParams p;
	memset(&p, 0, sizeof(p));
	p.op = op;
	p.wide = wide;
	p.bit_mask = bit_mask;
	p.num_data = num_data;
	p.data = d_data;
	p.edit_mask = reinterpret_cast<uint8_t const *>(d_edit_mask);
	p.result = d_result;
	cudaError_t const e = LaunchBitOperation(p, info.num_sms, static_cast<cudaStream_t>(cuda_stream));
	g_launch_count += 1;
	g_last_kernel = "bit_operation";
	if (e != cudaSuccess) {
		SetError("LaunchBitOperation", e);
		return sakura_Status_kUnknownError;
	}
	return sakura_Status_kOK;
}

} // namespace

#define SAKURA_B200_BITOP(NAME, OP) \
sakura_Status sakura_OperateBitwise##NAME##Uint8(uint8_t bit_mask, size_t num_data, uint8_t const data[], \
		bool const edit_mask[], uint8_t result[]) noexcept { \
	return BitOperationHost(OP, 0, bit_mask, num_data, data, edit_mask, result); \
} \
sakura_Status sakura_OperateBitwise##NAME##Uint32(uint32_t bit_mask, size_t num_data, uint32_t const data[], \
		bool const edit_mask[], uint32_t result[]) noexcept { \
	return BitOperationHost(OP, 1, bit_mask, num_data, data, edit_mask, result); \
}
SAKURA_B200_BITOP(And, kBitAnd)
SAKURA_B200_BITOP(ConverseNonImplication, kBitConverseNonImplication)
SAKURA_B200_BITOP(Implication, kBitImplication)
SAKURA_B200_BITOP(Or, kBitOr)
SAKURA_B200_BITOP(Xor, kBitXor)
#undef SAKURA_B200_BITOP

sakura_Status sakura_OperateBitwiseNotUint8(size_t num_data, uint8_t const data[], bool const edit_mask[],
		uint8_t result[]) noexcept {
	return BitOperationHost(kBitNot, 0, 0u, num_data, data, edit_mask, result);
}

sakura_Status sakura_OperateBitwiseNotUint32(size_t num_data, uint32_t const data[], bool const edit_mask[],
		uint32_t result[]) noexcept {
	return BitOperationHost(kBitNot, 1, 0u, num_data, data, edit_mask, result);
}

sakura_Status sakura_OperateBitwiseUint8Device(int operation, uint8_t bit_mask, size_t num_data,
		uint8_t const d_data[], bool const d_edit_mask[], uint8_t d_result[], void *cuda_stream) noexcept {
	return BitOperationDevice(operation, 0, bit_mask, num_data, d_data, d_edit_mask, d_result, cuda_stream);
}

sakura_Status sakura_OperateBitwiseUint32Device(int operation, uint32_t bit_mask, size_t num_data,
		uint32_t const d_data[], bool const d_edit_mask[], uint32_t d_result[], void *cuda_stream) noexcept {
	return BitOperationDevice(operation, 1, bit_mask, num_data, d_data, d_edit_mask, d_result, cuda_stream);
}

// ---------------------------------------------------------------- interpolation
namespace {

// argument rules of interpolation.cc:1288-1340 (in that order)
sakura_Status CheckInterpolate(int method, uint8_t polynomial_order, size_t num_base,
		void const *base_position, size_t num_array, void const *base_data, void const *base_mask,
		size_t num_interpolated, void const *interpolated_position, void const *interpolated_data,
		void const *interpolated_mask, bool host, bool *nothing_to_do, int *linear) {
	*nothing_to_do = false;
	CHECK_ARGS(num_base != 0);
	if (num_interpolated == 0 || num_array == 0) {
		*nothing_to_do = true;
		return sakura_Status_kOK;
	}
	if (host) {
		CHECK_ARGS(sakura_IsAligned(base_position) && sakura_IsAligned(base_data)
				&& sakura_IsAligned(interpolated_position) && sakura_IsAligned(interpolated_data)
				&& sakura_IsAligned(base_mask) && sakura_IsAligned(interpolated_mask));
	}
	CHECK_ARGS(base_position != nullptr && base_data != nullptr && interpolated_position != nullptr
			&& interpolated_data != nullptr && base_mask != nullptr && interpolated_mask != nullptr);
	// nearest, linear, and polynomial of order 0 (= nearest, :1419-1422). Polynomial (Neville) and
	// natural-spline interpolation are not part of this library: rejected, never approximated.
	if (method == 0 || (method == 2 && polynomial_order == 0)) {
		*linear = 0;
	} else if (method == 1) {
		*linear = 1;
	} else {
		snprintf(g_last_error, sizeof(g_last_error),
				"interpolation method %d is not implemented on the GPU path (nearest and linear are)", method);
		return sakura_Status_kInvalidArgument;
	}
	return sakura_Status_kOK;
}

sakura_Status InterpolateHost(int y_axis, int method, uint8_t polynomial_order, size_t num_base,
		double const *base_position, size_t num_array, float const *base_data, bool const *base_mask,
		size_t num_interpolated, double const *interpolated_position, float *interpolated_data,
		bool *interpolated_mask) {
	bool nothing = false;
	int linear = 0;
	sakura_Status st = CheckInterpolate(method, polynomial_order, num_base, base_position, num_array,
			base_data, base_mask, num_interpolated, interpolated_position, interpolated_data,
			interpolated_mask, true, &nothing, &linear);
	if (st != sakura_Status_kOK || nothing) {
		return st;
	}
	DeviceInfo info;
	st = CurrentDevice(&info);
	if (st != sakura_Status_kOK) {
		return st;
	}
	static thread_local DeviceBuffer d_bp, d_bd, d_bm, d_ip, d_od, d_om;
	size_t const nbase = num_base * num_array;
	size_t const nout = num_interpolated * num_array;
	CUDA_TRY(d_bp.Reserve(num_base * sizeof(double)));
	CUDA_TRY(d_bd.Reserve(nbase * sizeof(float)));
	CUDA_TRY(d_bm.Reserve(nbase));
	CUDA_TRY(d_ip.Reserve(num_interpolated * sizeof(double)));
	CUDA_TRY(d_od.Reserve(nout * sizeof(float)));
	CUDA_TRY(d_om.Reserve(nout));
	cudaStream_t const stream = 0;
	CUDA_TRY(cudaMemcpyAsync(d_bp.ptr, base_position, num_base * sizeof(double), cudaMemcpyHostToDevice, stream));
	CUDA_TRY(cudaMemcpyAsync(d_bd.ptr, base_data, nbase * sizeof(float), cudaMemcpyHostToDevice, stream));
	CUDA_TRY(cudaMemcpyAsync(d_bm.ptr, base_mask, nbase, cudaMemcpyHostToDevice, stream));
	CUDA_TRY(cudaMemcpyAsync(d_ip.ptr, interpolated_position, num_interpolated * sizeof(double),
			cudaMemcpyHostToDevice, stream));
	// arrays without a valid base point keep the caller's output values (interpolation.cc:1182-1185)
	CUDA_TRY(cudaMemcpyAsync(d_od.ptr, interpolated_data, nout * sizeof(float), cudaMemcpyHostToDevice, stream));
	InterpolateParams p;
	memset(&p, 0, sizeof(p));
	p.y_axis = y_axis;
	p.linear = linear;
	p.num_base = num_base;
	p.num_array = num_array;
	p.num_interpolated = num_interpolated;
	p.base_position = d_bp.As<double>();
	p.base_data = d_bd.As<float>();
	p.base_mask = d_bm.As<uint8_t>();
	p.interpolated_position = d_ip.As<double>();
	p.interpolated_data = d_od.As<float>();
	p.interpolated_mask = d_om.As<uint8_t>();
	cudaError_t const e = LaunchInterpolate(p, info.num_sms, stream);
	g_launch_count += 1;
	g_last_kernel = "interpolate";
	if (e != cudaSuccess) {
		SetError("LaunchInterpolate", e);
		return sakura_Status_kUnknownError;
	}
	CUDA_TRY(cudaMemcpyAsync(interpolated_data, d_od.ptr, nout * sizeof(float), cudaMemcpyDeviceToHost, stream));
	CUDA_TRY(cudaMemcpyAsync(interpolated_mask, d_om.ptr, nout, cudaMemcpyDeviceToHost, stream));
	CUDA_TRY(cudaStreamSynchronize(stream));
	return sakura_Status_kOK;
}

sakura_Status InterpolateDevice(int y_axis, int method, uint8_t polynomial_order, size_t num_base,
		double const *d_base_position, size_t num_array, float const *d_base_data, bool const *d_base_mask,
		size_t num_interpolated, double const *d_interpolated_position, float *d_interpolated_data,
		bool *d_interpolated_mask, void *cuda_stream) {
	bool nothing = false;
	int linear = 0;
	sakura_Status st = CheckInterpolate(method, polynomial_order, num_base, d_base_position, num_array,
			d_base_data, d_base_mask, num_interpolated, d_interpolated_position, d_interpolated_data,
			d_interpolated_mask, false, &nothing, &linear);
	if (st != sakura_Status_kOK || nothing) {
		return st;
	}
	DeviceInfo info;
	st = CurrentDevice(&info);
	if (st != sakura_Status_kOK) {
		return st;
	}
	InterpolateParams p;
	memset(&p, 0, sizeof(p));
	p.y_axis = y_axis;
	p.linear = linear;
	p.num_base = num_base;
	p.num_array = num_array;
	p.num_interpolated = num_interpolated;
	p.base_position = d_base_position;
	p.base_data = d_base_data;
	p.base_mask = reinterpret_cast<uint8_t const *>(d_base_mask);
	p.interpolated_position = d_interpolated_position;
	p.interpolated_data = d_interpolated_data;
	p.interpolated_mask = reinterpret_cast<uint8_t *>(d_interpolated_mask);
	cudaError_t const e = LaunchInterpolate(p, info.num_sms, static_cast<cudaStream_t>(cuda_stream));
	g_launch_count += 1;
	g_last_kernel = "interpolate";
	if (e != cudaSuccess) {
		SetError("LaunchInterpolate", e);
		return sakura_Status_kUnknownError;
	}
	return sakura_Status_kOK;
}

} // namespace

sakura_Status sakura_InterpolateXAxisFloat(sakura_InterpolationMethod interpolation_method,
		uint8_t polynomial_order, size_t num_base, double const base_position[], size_t num_array,
		float const base_data[], bool const base_mask[], size_t num_interpolated,
		double const interpolated_position[], float interpolated_data[],
		bool interpolated_mask[]) noexcept {
	return InterpolateHost(0, static_cast<int>(interpolation_method), polynomial_order, num_base,
			base_position, num_array, base_data, base_mask, num_interpolated, interpolated_position,
			interpolated_data, interpolated_mask);
}

sakura_Status sakura_InterpolateYAxisFloat(sakura_InterpolationMethod interpolation_method,
		uint8_t polynomial_order, size_t num_base, double const base_position[], size_t num_array,
		float const base_data[], bool const base_mask[], size_t num_interpolated,
		double const interpolated_position[], float interpolated_data[],
		bool interpolated_mask[]) noexcept {
	return InterpolateHost(1, static_cast<int>(interpolation_method), polynomial_order, num_base,
			base_position, num_array, base_data, base_mask, num_interpolated, interpolated_position,
			interpolated_data, interpolated_mask);
}

sakura_Status sakura_InterpolateXAxisFloatDevice(sakura_InterpolationMethod interpolation_method,
		uint8_t polynomial_order, size_t num_base, double const d_base_position[], size_t num_array,
		float const d_base_data[], bool const d_base_mask[], size_t num_interpolated,
		double const d_interpolated_position[], float d_interpolated_data[], bool d_interpolated_mask[],
		void *cuda_stream) noexcept {
	return InterpolateDevice(0, static_cast<int>(interpolation_method), polynomial_order, num_base,
			d_base_position, num_array, d_base_data, d_base_mask, num_interpolated, d_interpolated_position,
			d_interpolated_data, d_interpolated_mask, cuda_stream);
}

sakura_Status sakura_InterpolateYAxisFloatDevice(sakura_InterpolationMethod interpolation_method,
		uint8_t polynomial_order, size_t num_base, double const d_base_position[], size_t num_array,
		float const d_base_data[], bool const d_base_mask[], size_t num_interpolated,
		double const d_interpolated_position[], float d_interpolated_data[], bool d_interpolated_mask[],
		void *cuda_stream) noexcept {
	return InterpolateDevice(1, static_cast<int>(interpolation_method), polynomial_order, num_base,
			d_base_position, num_array, d_base_data, d_base_mask, num_interpolated, d_interpolated_position,
			d_interpolated_data, d_interpolated_mask, cuda_stream);
}

// ---------------------------------------------------------------- calibration
namespace sakura_b200 {
namespace capi {
// shared argument rules of the calibration entry points: a row stride of 0 shares one array
sakura_Status CheckCalibration(size_t num_spectra, size_t num_data, float const *scaling,
		size_t scaling_stride, float const *data, size_t data_stride, float const *reference,
		size_t reference_stride) {
	CHECK_ARGS(data != nullptr && reference != nullptr);
	CHECK_ARGS(data_stride >= num_data || num_spectra <= 1);
	CHECK_ARGS(reference_stride == 0 || reference_stride >= num_data || num_spectra <= 1);
	CHECK_ARGS(scaling == nullptr || scaling_stride == 0 || scaling_stride >= num_data
			|| num_spectra <= 1);
	return sakura_Status_kOK;
}
} // namespace capi
} // namespace sakura_b200

namespace {
sakura_Status CalibrateHost(float const *scaling, float const_scaling, size_t num_data,
		float const *data, float const *reference, float *result) {
	if (num_data == 0) { // normalization.cc:229-232: nothing to do, before any other check
		return sakura_Status_kOK;
	}
	CHECK_ARGS(data != nullptr && reference != nullptr && result != nullptr); // :234-238
	CHECK_ARGS(sakura_IsAligned(data) && sakura_IsAligned(reference) && sakura_IsAligned(result));
	DeviceInfo info;
	sakura_Status st = CurrentDevice(&info);
	if (st != sakura_Status_kOK) {
		return st;
	}
	size_t const bytes = num_data * sizeof(float);
	static thread_local DeviceBuffer d_t, d_r, d_s, d_o;
	CUDA_TRY(d_t.Reserve(bytes));
	CUDA_TRY(d_r.Reserve(bytes));
	CUDA_TRY(d_o.Reserve(bytes));
	cudaStream_t const stream = 0;
	CUDA_TRY(cudaMemcpyAsync(d_t.ptr, data, bytes, cudaMemcpyHostToDevice, stream));
	CUDA_TRY(cudaMemcpyAsync(d_r.ptr, reference, bytes, cudaMemcpyHostToDevice, stream));
	CalibrateParams p;
	memset(&p, 0, sizeof(p));
	if (scaling != nullptr) {
		CUDA_TRY(d_s.Reserve(bytes));
		CUDA_TRY(cudaMemcpyAsync(d_s.ptr, scaling, bytes, cudaMemcpyHostToDevice, stream));
		p.spec.scaling = d_s.As<float>();
	}
	p.num_spectra = 1;
	p.num_data = num_data;
	p.data = d_t.As<float>();
	p.data_stride = num_data;
	p.spec.reference = d_r.As<float>();
	p.spec.reference_stride = num_data;
	p.spec.scaling_stride = num_data;
	p.spec.const_scaling = const_scaling;
	p.result = d_o.As<float>();
	p.result_stride = num_data;
	cudaError_t const e = LaunchCalibrate(p, info.num_sms, stream);
	g_launch_count += 1;
	g_last_kernel = "calibrate";
	if (e != cudaSuccess) {
		SetError("LaunchCalibrate", e);
		return sakura_Status_kUnknownError;
	}
	// result may alias data or reference (in-place is allowed, sakura.h:1112-1114): both were
	// copied to the device before anything is written back
	CUDA_TRY(cudaMemcpyAsync(result, d_o.ptr, bytes, cudaMemcpyDeviceToHost, stream));
	CUDA_TRY(cudaStreamSynchronize(stream));
	return sakura_Status_kOK;
}
} // namespace

sakura_Status sakura_CalibrateDataWithArrayScalingFloat(size_t num_data,
		float const scaling_factor[], float const data[], float const reference[],
		float result[]) noexcept {
	if (num_data == 0) {
		return sakura_Status_kOK;
	}
	CHECK_ARGS(scaling_factor != nullptr && sakura_IsAligned(scaling_factor)); // :186-198
	return CalibrateHost(scaling_factor, 0.0f, num_data, data, reference, result);
}

sakura_Status sakura_CalibrateDataWithConstScalingFloat(float scaling_factor, size_t num_data,
		float const data[], float const reference[], float result[]) noexcept {
	return CalibrateHost(nullptr, scaling_factor, num_data, data, reference, result);
}

sakura_Status sakura_CalibrateDataFloatBatchDevice(size_t num_spectra, size_t num_data,
		float const d_scaling_factor[], size_t scaling_stride, float const_scaling_factor,
		float const d_data[], size_t data_stride, float const d_reference[],
		size_t reference_stride, float d_result[], size_t result_stride,
		void *cuda_stream) noexcept {
	if (num_spectra == 0 || num_data == 0) {
		return sakura_Status_kOK;
	}
	CHECK_ARGS(d_result != nullptr && (result_stride >= num_data || num_spectra <= 1));
	sakura_Status st = CheckCalibration(num_spectra, num_data, d_scaling_factor, scaling_stride,
			d_data, data_stride, d_reference, reference_stride);
	if (st != sakura_Status_kOK) {
		return st;
	}
	DeviceInfo info;
	st = CurrentDevice(&info);
	if (st != sakura_Status_kOK) {
		return st;
	}
	CalibrateParams p;
	memset(&p, 0, sizeof(p));
	p.num_spectra = num_spectra;
	p.num_data = num_data;
	p.data = d_data;
	p.data_stride = data_stride;
	p.spec.reference = d_reference;
	p.spec.reference_stride = reference_stride;
	p.spec.scaling = d_scaling_factor;
	p.spec.scaling_stride = scaling_stride;
	p.spec.const_scaling = const_scaling_factor;
	p.result = d_result;
	p.result_stride = result_stride;
	cudaError_t const e = LaunchCalibrate(p, info.num_sms, static_cast<cudaStream_t>(cuda_stream));
	g_launch_count += 1;
	g_last_kernel = "calibrate";
	if (e != cudaSuccess) {
		SetError("LaunchCalibrate", e);
		return sakura_Status_kUnknownError;
	}
	return sakura_Status_kOK;
}
