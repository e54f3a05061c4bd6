// Internals shared by the translation units of the C ABI (capi_*.cu): global state, device buffers,
// the context object and the fit dispatch. Not installed; include/sakura_b200.h is the interface.
#pragma once

#include "../../include/sakura_b200.h"

#include "basis_tables.h"
#include "device_common.cuh"
#include "lsq_blocks.cuh"
#include "lsqfit_general.cuh"
#include "lsqfit_poly.cuh"
#include "lsqfit_sinusoid.cuh"
#include "lsqfit_spline.cuh"
#include "calibration.cuh"
#include "bool_filters.cuh"
#include "interpolation.cuh"
#include "statistics.cuh"

#include <cuda_runtime.h>

#include <atomic>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <new>
#include <string>
#include <utility>
#include <vector>

namespace sakura_b200 {
namespace capi {

constexpr size_t kAlignment = 32; // the reference's AVX build value (localdef.h:75-86)
constexpr uint64_t kContextMagic = 0x53414b5552414232ull; // "SAKURAB2"

void *DefaultAllocator(size_t size);
void DefaultFree(void *p);
extern sakura_UserAllocator g_allocator;
extern sakura_UserDeallocator g_deallocator;
extern std::atomic<uint64_t> g_launch_count;
extern thread_local char g_last_error[512];
extern thread_local char const *g_last_kernel;

void SetError(char const *where, cudaError_t err);

inline bool IsAligned(void const *p) {
	return (reinterpret_cast<uintptr_t>(p) % kAlignment) == 0;
}

#define CHECK_ARGS(x) do { if (!(x)) { return sakura_Status_kInvalidArgument; } } while (false)

#define CUDA_TRY(expr) do { cudaError_t e__ = (expr); if (e__ != cudaSuccess) { \
	SetError(#expr, e__); \
	return (e__ == cudaErrorMemoryAllocation) ? sakura_Status_kNoMemory : sakura_Status_kUnknownError; } \
} while (false)

// ---------------------------------------------------------------- device buffers
struct DeviceBuffer {
	void *ptr = nullptr;
	size_t cap = 0;
	cudaError_t Reserve(size_t bytes) {
		if (bytes <= cap) {
			return cudaSuccess;
		}
		if (ptr != nullptr) {
			cudaFree(ptr);
			ptr = nullptr;
			cap = 0;
		}
		size_t const want = bytes + bytes / 8 + 256;
		cudaError_t e = cudaMalloc(&ptr, want);
		if (e != cudaSuccess) {
			(void) cudaGetLastError();
			e = cudaMalloc(&ptr, bytes);
			if (e != cudaSuccess) {
				ptr = nullptr;
				return e;
			}
			cap = bytes;
			return cudaSuccess;
		}
		cap = want;
		return cudaSuccess;
	}
	void Release() {
		if (ptr != nullptr) {
			cudaFree(ptr);
		}
		ptr = nullptr;
		cap = 0;
	}
	template<typename T> T *As() const {
		return static_cast<T *>(ptr);
	}
};

struct DeviceInfo {
	int device = -1;
	int num_sms = 0;
};

sakura_Status CurrentDevice(DeviceInfo *info);

constexpr int kPipelineSlots = 4;

/** One stage of the host-buffer pipeline: a stream and the device buffers of one chunk. */
struct PipelineSlot {
	// One direction per stream: a stream that has carried device->host copies gets its later
	// host->device copies queued on the same copy engine (measured: uploads then wait for the bulk
	// downloads of other slots), so uploads + kernels live on `stream`, every download on `down`.
	cudaStream_t stream = nullptr; // upload + kernel
	cudaStream_t down = nullptr; // per-row status / flags download, then the bulk result download
	cudaEvent_t kernel_done = nullptr; // the chunk's kernel(s) have finished
	cudaEvent_t flags_ready = nullptr; // ... and its per-row flags are on the host
	cudaEvent_t down_done = nullptr; // result buffers of this slot may be overwritten
	DeviceBuffer data, mask, coeff, best_fit, residual, final_mask, small, handover, boundary;
	int32_t *h_flags = nullptr; // page-locked (a pageable target would make the copy block the host)
	size_t h_flags_cap = 0;
	// bit-packed mask transfer (host_bits.h): device words, page-locked staging of both directions, and
	// the chunk whose downloaded final-mask bits still wait to be expanded into the caller's array
	DeviceBuffer mask_bits, fmask_bits;
	uint32_t *h_mask_bits = nullptr, *h_fmask_bits = nullptr;
	size_t h_bits_cap = 0; // words each
	size_t unpack_rows = 0, unpack_row0 = 0;
	bool unpack_all = true;
	std::vector<int32_t> unpack_flags; // per-row flags of that chunk when not every row published
	size_t row0 = 0;
	size_t rows = 0; // non-zero while a chunk is in flight in this slot
	size_t chunk = 0; // its index in the call
	void Release() {
		data.Release();
		mask.Release();
		coeff.Release();
		best_fit.Release();
		residual.Release();
		final_mask.Release();
		small.Release();
		handover.Release();
		boundary.Release();
		if (stream != nullptr) {
			cudaStreamDestroy(stream);
			stream = nullptr;
		}
		if (down != nullptr) {
			cudaStreamDestroy(down);
			down = nullptr;
		}
		if (flags_ready != nullptr) {
			cudaEventDestroy(flags_ready);
			flags_ready = nullptr;
		}
		if (kernel_done != nullptr) {
			cudaEventDestroy(kernel_done);
			kernel_done = nullptr;
		}
		if (h_flags != nullptr) {
			cudaFreeHost(h_flags);
			h_flags = nullptr;
			h_flags_cap = 0;
		}
		mask_bits.Release();
		fmask_bits.Release();
		if (h_mask_bits != nullptr) {
			cudaFreeHost(h_mask_bits);
			h_mask_bits = nullptr;
		}
		if (h_fmask_bits != nullptr) {
			cudaFreeHost(h_fmask_bits);
			h_fmask_bits = nullptr;
		}
		h_bits_cap = 0;
		unpack_rows = 0;
		if (down_done != nullptr) {
			cudaEventDestroy(down_done);
			down_done = nullptr;
		}
	}
};

} // namespace capi
} // namespace sakura_b200

// ---------------------------------------------------------------- context
// Mirrors what the reference context carries (baseline.h:58-84): fit type, parameter,
// channel count, numbers of bases, the FP64 basis table -- plus the device-side copy
// of the table and per-context device scratch.  Like the reference's, it is mutated
// by fit calls although passed as const, hence one context per thread.
struct sakura_LSQFitContextFloat {
	uint64_t magic;
	int type; // FitType
	uint16_t param; // order / npiece / nwave
	size_t num_data;
	size_t num_bases; // columns of the context table
	size_t num_lsq_bases_max;
	double *h_basis; // host table, user allocator
	// device state
	int device;
	sakura_b200::capi::DeviceBuffer d_basis;
	sakura_b200::capi::DeviceBuffer ws_residual, ws_best_fit, ws_spline;
	// sinusoid fast path: extended trigonometric table (1, sin/cos(m theta) for m <= 2*nwave),
	// its column sums over all channels, the per-call compact basis table and scratch
	std::vector<double> *h_ext;
	int ext_cols;
	sakura_b200::capi::DeviceBuffer d_ext, d_table, d_useidx, d_fallback, ws_mask, ws_sin_residual,
			ws_sin_best_fit;
	sakura_b200::capi::DeviceBuffer ws_cal, ws_calmask; // calibrated copy of a batch for the paths that cannot fuse it
	// staging for the host-pointer entry points
	sakura_b200::capi::DeviceBuffer s_data, s_mask, s_coeff, s_best_fit, s_residual, s_final_mask, s_rms,
			s_status, s_lsq, s_flags, s_boundary;
	std::vector<int32_t> *h_flags;
	int32_t const *last_handover; // device counter of the rows the last fit handed to a slower kernel
	sakura_b200::capi::PipelineSlot *slots; // kPipelineSlots entries, created on first pipelined call
	// small host batches (the reference's single-row signatures above all): one pinned host arena, one
	// device arena, one upload and one download per call on a stream of the context's own
	void *h_pack;
	size_t h_pack_cap;
	sakura_b200::capi::DeviceBuffer d_pack;
	cudaStream_t pack_stream;
	// ...MultiDevice entry points: one replica of this context per device (created on first use)
	std::vector<std::pair<int, sakura_LSQFitContextFloat *> > *replicas;
};

namespace sakura_b200 {
namespace capi {

typedef sakura_LSQFitContextFloat Context;

inline bool ValidContext(Context const *c) {
	return c != nullptr && c->magic == kContextMagic;
}

size_t ContextBases(int type, uint16_t param);
size_t LsqBases(int type, size_t param);
void ReleaseDeviceState(Context *c);
sakura_Status CreateContext(int type, uint16_t param, size_t num_data, Context **out);
/** Makes sure the basis table lives on the current device. */
sakura_Status BindDevice(Context *c, DeviceInfo const &info, cudaStream_t stream);

inline size_t RoundUp(size_t v, size_t m) {
	return (v + m - 1) / m * m;
}

// ---------------------------------------------------------------- fit dispatch
struct FitCall {
	int type; // FitType of the call
	size_t num_spectra;
	size_t num_data;
	size_t data_stride, mask_stride;
	float clip;
	uint16_t nfit;
	size_t K; // fitted bases
	size_t npiece; // spline
	size_t coeff_stride;
	int16_t use_idx[kMaxBases];
	// device pointers
	float const *data;
	uint8_t const *mask;
	double *coeff;
	float *best_fit;
	float *residual;
	uint8_t *final_mask;
	float *rms;
	int32_t *status;
	int32_t *lsq;
	int32_t *flags;
	unsigned long long *boundary;
	CalibrationSpec cal; // optional calibration fused in front of the fit (reference == NULL: none)
	// hand-over list of the screening kernel (num_spectra + 1 int32); NULL: the context's own buffer.
	// Chunks of the host pipeline run concurrently on one context and bring their slot's buffer.
	int32_t *handover;
};

sakura_Status LaunchFit(Context *c, FitCall const &f_in, cudaStream_t stream);

// ---- argument checks shared by the single-row, batch and device entry points -----

inline bool RowsAligned(void const *base, size_t stride_bytes, size_t num_spectra) {
	if (!IsAligned(base)) {
		return false;
	}
	return num_spectra <= 1 || (stride_bytes % kAlignment) == 0;
}

inline bool DeviceRowsAligned(void const *base, size_t stride_bytes, size_t num_spectra) {
	if ((reinterpret_cast<uintptr_t>(base) % 16) != 0) {
		return false;
	}
	return num_spectra <= 1 || (stride_bytes % 16) == 0;
}

sakura_Status CheckCalibration(size_t num_spectra, size_t num_data, float const *scaling,
		size_t scaling_stride, float const *data, size_t data_stride, float const *reference,
		size_t reference_stride);

} // namespace capi
} // namespace sakura_b200
