// Boolean-filter builders of libsakura (libsakura/src/bool_filter_collection.cc): the functions that
// produce the `mask` the baseline fit consumes -- value ranges, thresholds, NaN/Inf flags,
// uint8/uint32 -> bool, inversion -- as one streaming kernel.
#ifndef SAKURA_B200_BOOL_FILTERS_CUH_
#define SAKURA_B200_BOOL_FILTERS_CUH_

#include <stddef.h>
#include <stdint.h>
#include <cuda_runtime.h>

namespace sakura_b200 {

enum BoolFilterOp : int {
	kFilterRangesInclusive = 0, // OR_j lower_j <= x <= upper_j
	kFilterRangesExclusive = 1, // OR_j lower_j <  x <  upper_j
	kFilterGreaterThan = 2,
	kFilterGreaterThanOrEquals = 3,
	kFilterLessThan = 4,
	kFilterLessThanOrEquals = 5,
	kFilterFiniteFloat = 6, // false for NaN and +-Inf
	kFilterNonZero = 7, // uint8 / uint32 -> bool
	kFilterInvertBool = 8 // byte ^ 1
};

enum BoolFilterType : int {
	kFilterFloat = 0, kFilterInt32 = 1, kFilterUint8 = 2, kFilterUint32 = 3
};

constexpr int kFilterInlineBounds = 16;

struct BoolFilterParams {
	int op;
	int type;
	size_t num_data;
	void const *data;
	uint8_t *result;
	// optional: result = filter(data) AND and_with (the merge every caller does next); may alias result
	uint8_t const *and_with;
	// ranges: bounds as raw 32-bit words (float or int32); in the parameter block up to
	// kFilterInlineBounds conditions, else in device memory (lower then upper, num_condition each)
	size_t num_condition;
	uint32_t lower[kFilterInlineBounds];
	uint32_t upper[kFilterInlineBounds];
	uint32_t const *bounds;
	// The reference evaluates "lower <= x && x <= upper" in its SIMD body and the product form
	// "(x - lower) * (upper - x) >= 0" in the scalar remainder (elements from compare_end on) and for
	// more than 16 conditions (bool_filter_collection.cc:41-250). The two differ for NaN-free
	// pathological inputs only (underflowing or overflowing products); both are reproduced.
	size_t compare_end;
	uint32_t threshold; // raw float / int32
};

cudaError_t LaunchBoolFilter(BoolFilterParams const &p, int num_sms, cudaStream_t stream);

// ---- bit operations with an edit mask (libsakura/src/bit_operation.cc): the way back from a bool
// mask to the flag bytes of a data set (bench/e2eReduce/src/main.cc:59-78)
enum BitOp : int {
	kBitAnd = 0, // data & bit_mask
	kBitConverseNonImplication = 1, // ~data & bit_mask
	kBitImplication = 2, // ~data | bit_mask
	kBitNot = 3, // ~data
	kBitOr = 4, // data | bit_mask
	kBitXor = 5 // data ^ bit_mask
};

struct BitOperationParams {
	int op;
	int wide; // 1: uint32 elements, 0: uint8
	uint32_t bit_mask;
	size_t num_data;
	void const *data;
	uint8_t const *edit_mask; // elements whose byte is 0 keep their value (any other value: as written
	                          // in bit_operation.cc, "mask - 1" in the element type)
	void *result; // may alias data
};

cudaError_t LaunchBitOperation(BitOperationParams const &p, int num_sms, cudaStream_t stream);

} // namespace sakura_b200

#endif
