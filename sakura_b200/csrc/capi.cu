// C ABI of libsakura_b200.so: argument validation, context objects, device staging and
// kernel dispatch.  Declarations and the reference interface each entry point
// replaces are in include/sakura_b200.h.
//
// There is NO CPU fallback: every compute call runs a CUDA kernel of this library
// and fails loudly (status + message on stderr) if no CUDA device is usable.
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
#include <vector>

using namespace sakura_b200;

namespace {

constexpr size_t kAlignment = 32; // the reference's AVX build value (localdef.h:75-86)
constexpr uint64_t kContextMagic = 0x53414b5552414232ull; // "SAKURAB2"

// ---------------------------------------------------------------- global state
void *DefaultAllocator(size_t size) {
	void *p = nullptr;
	if (posix_memalign(&p, kAlignment, size < 1 ? 1 : size) != 0) {
		return nullptr;
	}
	return p;
}
void DefaultFree(void *p) {
	free(p);
}

sakura_UserAllocator g_allocator = DefaultAllocator;
sakura_UserDeallocator g_deallocator = DefaultFree;
std::atomic<uint64_t> g_launch_count(0);
thread_local char g_last_error[512] = "";
thread_local char const *g_last_kernel = "";

void SetError(char const *where, cudaError_t err) {
	snprintf(g_last_error, sizeof(g_last_error), "%s: %s (%s)", where, cudaGetErrorName(err),
			cudaGetErrorString(err));
	fprintf(stderr, "sakura_b200: CUDA failure in %s\n", g_last_error);
}

bool IsAligned(void const *p) {
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

sakura_Status CurrentDevice(DeviceInfo *info) {
	int dev = -1;
	CUDA_TRY(cudaGetDevice(&dev));
	static std::mutex mu;
	static std::vector<int> sms;
	std::lock_guard<std::mutex> lock(mu);
	if (static_cast<size_t>(dev) >= sms.size()) {
		sms.resize(dev + 1, 0);
	}
	if (sms[dev] == 0) {
		int v = 0;
		CUDA_TRY(cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev));
		sms[dev] = v;
	}
	info->device = dev;
	info->num_sms = sms[dev];
	return sakura_Status_kOK;
}

constexpr int kPipelineSlots = 4;

/** One stage of the host-buffer pipeline: a stream and the device buffers of one chunk. */
struct PipelineSlot {
	cudaStream_t stream = nullptr; // upload + kernel + per-row status download
	cudaStream_t down = nullptr; // bulk result download, so that it never blocks the next upload
	cudaEvent_t flags_ready = nullptr; // kernel done and per-row flags on the host
	cudaEvent_t down_done = nullptr; // result buffers of this slot may be overwritten
	DeviceBuffer data, mask, coeff, best_fit, residual, final_mask, small, handover;
	std::vector<int32_t> h_flags;
	size_t row0 = 0;
	size_t rows = 0; // non-zero while a chunk is in flight in this slot
	void Release() {
		data.Release();
		mask.Release();
		coeff.Release();
		best_fit.Release();
		residual.Release();
		final_mask.Release();
		small.Release();
		handover.Release();
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
		if (down_done != nullptr) {
			cudaEventDestroy(down_done);
			down_done = nullptr;
		}
	}
};

} // namespace

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
	DeviceBuffer d_basis;
	DeviceBuffer ws_residual, ws_best_fit, ws_spline;
	// sinusoid fast path: extended trigonometric table (1, sin/cos(m theta) for m <= 2*nwave),
	// its column sums over all channels, the per-call compact basis table and scratch
	std::vector<double> *h_ext;
	int ext_cols;
	DeviceBuffer d_ext, d_table, d_useidx, d_fallback, ws_mask, ws_sin_residual,
			ws_sin_best_fit;
	DeviceBuffer ws_cal, ws_calmask; // calibrated copy of a batch for the paths that cannot fuse it
	// staging for the host-pointer entry points
	DeviceBuffer s_data, s_mask, s_coeff, s_best_fit, s_residual, s_final_mask, s_rms,
			s_status, s_lsq, s_flags, s_boundary;
	std::vector<int32_t> *h_flags;
	int32_t const *last_handover; // device counter of the rows the last fit handed to a slower kernel
	PipelineSlot *slots; // kPipelineSlots entries, created on first pipelined call
};

namespace {

typedef sakura_LSQFitContextFloat Context;

bool ValidContext(Context const *c) {
	return c != nullptr && c->magic == kContextMagic;
}

size_t ContextBases(int type, uint16_t param) {
	switch (type) {
	case kFitPolynomial:
	case kFitChebyshev:
		return static_cast<size_t>(param) + 1;
	case kFitCubicSpline:
		return 4;
	default:
		return 2 * static_cast<size_t>(param) + 1;
	}
}

size_t LsqBases(int type, size_t param) {
	if (type == kFitCubicSpline) {
		return 3 + param;
	}
	return ContextBases(type, static_cast<uint16_t>(param));
}

void ReleaseDeviceState(Context *c) {
	int prev = -1;
	if (c->device >= 0 && cudaGetDevice(&prev) == cudaSuccess) {
		if (prev != c->device) {
			cudaSetDevice(c->device);
		}
		c->d_basis.Release();
		c->ws_residual.Release();
		c->ws_best_fit.Release();
		c->ws_spline.Release();
		c->d_ext.Release();
		c->d_table.Release();
		c->d_useidx.Release();
		c->d_fallback.Release();
		c->ws_mask.Release();
		c->ws_sin_residual.Release();
		c->ws_sin_best_fit.Release();
		c->ws_cal.Release();
		c->ws_calmask.Release();
		c->s_data.Release();
		c->s_mask.Release();
		c->s_coeff.Release();
		c->s_best_fit.Release();
		c->s_residual.Release();
		c->s_final_mask.Release();
		c->s_rms.Release();
		c->s_status.Release();
		c->s_lsq.Release();
		c->s_flags.Release();
		c->s_boundary.Release();
		if (c->slots != nullptr) {
			for (int s = 0; s < kPipelineSlots; ++s) {
				c->slots[s].Release();
			}
		}
		if (prev != c->device) {
			cudaSetDevice(prev);
		}
	}
	c->device = -1;
}

sakura_Status CreateContext(int type, uint16_t param, size_t num_data, Context **out) {
	size_t const num_bases = ContextBases(type, param);
	size_t const lsq_max = LsqBases(type, param);
	if (lsq_max > static_cast<size_t>(kMaxBases)) {
		fprintf(stderr, "sakura_b200: %zu simultaneous LSQ bases requested; this build supports "
				"at most %d\n", lsq_max, kMaxBases);
		return sakura_Status_kInvalidArgument;
	}
	if (num_data > static_cast<size_t>(INT32_MAX)) {
		return sakura_Status_kInvalidArgument;
	}
	void *mem = g_allocator(sizeof(Context));
	if (mem == nullptr) {
		return sakura_Status_kNoMemory;
	}
	Context *c = new (mem) Context();
	c->magic = kContextMagic;
	c->type = type;
	c->param = param;
	c->num_data = num_data;
	c->num_bases = num_bases;
	c->num_lsq_bases_max = lsq_max;
	c->device = -1;
	c->h_flags = nullptr;
	c->last_handover = nullptr;
	c->slots = nullptr;
	c->h_ext = nullptr;
	c->ext_cols = 0;
	c->h_basis = static_cast<double *>(g_allocator(sizeof(double) * num_bases * num_data));
	if (c->h_basis == nullptr) {
		c->~Context();
		g_deallocator(mem);
		return sakura_Status_kNoMemory;
	}
	switch (type) {
	case kFitPolynomial:
	case kFitCubicSpline:
		FillPolynomialTable(num_data, num_bases, c->h_basis);
		break;
	case kFitChebyshev:
		FillChebyshevTable(num_data, num_bases, c->h_basis);
		if (num_data > 1 && param >= 1 && 2 * static_cast<size_t>(param) + 1 <= 96) {
			// T_j T_k = (T_{j+k} + T_{|j-k|}) / 2: the Gram matrix of a Chebyshev fit needs only the
			// masked sums of T_m, m <= 2*order, exactly like the cosine terms of the sinusoid kernel
			// (T_m(cos t) = cos(m t)). Extended table: column m = T_m(x_i), same chunked layout.
			int const cols = 2 * param + 1;
			size_t const chunks = static_cast<size_t>(SinusoidExtLd(cols)) / 32;
			size_t const chunk_doubles = static_cast<size_t>(num_data) * kSinExtChunkLd;
			c->ext_cols = cols;
			c->h_ext = new std::vector<double>(chunks * chunk_doubles, 0.0);
			double const max_x = static_cast<double>(num_data - 1);
			std::vector<long double> t(cols);
			for (size_t i = 0; i < num_data; ++i) {
				long double const x = static_cast<long double>(
						2.0 * static_cast<double>(i) / max_x - 1.0); // the table's own x
				t[0] = 1.0L;
				if (cols > 1) {
					t[1] = x;
				}
				for (int m = 2; m < cols; ++m) {
					t[m] = 2.0L * x * t[m - 1] - t[m - 2];
				}
				for (int m = 0; m < cols; ++m) {
					(*c->h_ext)[(m / 32) * chunk_doubles + i * kSinExtChunkLd + (m % 32)] =
							static_cast<double>(t[m]);
				}
			}
		}
		break;
	default:
		FillSinusoidTable(num_data, param, c->h_basis);
		if (num_data > 1 && param >= 1 && 4 * static_cast<size_t>(param) + 1 <= 96) {
			// same expression as the reference's table (baseline.cc:356-362), up to 2*nwave
			// chunk-major layout [chunks][num_data][kSinExtChunkLd], 32 columns per chunk, zero padded
			int const cols = 4 * param + 1;
			size_t const chunks = static_cast<size_t>(SinusoidExtLd(cols)) / 32;
			size_t const chunk_doubles = static_cast<size_t>(num_data) * kSinExtChunkLd;
			c->ext_cols = cols;
			c->h_ext = new std::vector<double>(chunks * chunk_doubles, 0.0);
			double const norm = 2.0 * M_PI / static_cast<double>(num_data - 1);
			auto at = [&](size_t i, int col) -> double & {
				return (*c->h_ext)[(col / 32) * chunk_doubles + i * kSinExtChunkLd + (col % 32)];
			};
			for (size_t i = 0; i < num_data; ++i) {
				at(i, 0) = 1.0;
				for (int m = 1; m <= 2 * param; ++m) {
					double const x = static_cast<double>(i) * static_cast<double>(m) * norm;
					at(i, 2 * m - 1) = sin(x);
					at(i, 2 * m) = cos(x);
				}
			}
		}
		break;
	}
	*out = c;
	return sakura_Status_kOK;
}

/** Makes sure the basis table lives on the current device. */
sakura_Status BindDevice(Context *c, DeviceInfo const &info, cudaStream_t stream) {
	if (c->device == info.device && c->d_basis.ptr != nullptr) {
		return sakura_Status_kOK;
	}
	if (c->device >= 0 && c->device != info.device) {
		ReleaseDeviceState(c);
	}
	c->device = info.device;
	size_t const bytes = sizeof(double) * c->num_bases * c->num_data;
	CUDA_TRY(c->d_basis.Reserve(bytes));
	CUDA_TRY(cudaMemcpyAsync(c->d_basis.ptr, c->h_basis, bytes, cudaMemcpyHostToDevice, stream));
	if (c->h_ext != nullptr) {
		CUDA_TRY(c->d_ext.Reserve(c->h_ext->size() * sizeof(double)));
		CUDA_TRY(cudaMemcpyAsync(c->d_ext.ptr, c->h_ext->data(), c->h_ext->size() * sizeof(double),
				cudaMemcpyHostToDevice, stream));
	}
	return sakura_Status_kOK;
}

size_t RoundUp(size_t v, size_t m) {
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

sakura_Status LaunchFit(Context *c, FitCall const &f_in, cudaStream_t stream) {
	DeviceInfo info;
	sakura_Status st = CurrentDevice(&info);
	if (st != sakura_Status_kOK) {
		return st;
	}
	st = BindDevice(c, info, stream);
	if (st != sakura_Status_kOK) {
		return st;
	}
	if (f_in.num_spectra == 0) {
		return sakura_Status_kOK;
	}
	FitCall f = f_in;
	if (f.cal.reference != nullptr) {
		// The polynomial kernel calibrates while it loads a row. Every other path gets the
		// calibrated spectra (and the NaN/Inf-merged mask) from a pre-pass into context scratch.
		bool const fused = (f.type == kFitPolynomial || (f.type == kFitChebyshev && f.K <= 6))
				&& PolyFusedCalibrationSupported(f.num_data)
				&& PolyFastSupported(f.num_data, f.K, f.data_stride, f.mask_stride, f.data, f.mask,
						f.best_fit, f.residual, f.final_mask)
				&& (reinterpret_cast<uintptr_t>(f.cal.reference) % 16) == 0
				&& f.cal.reference_stride % 4 == 0
				&& (reinterpret_cast<uintptr_t>(f.cal.scaling) % 16) == 0
				&& f.cal.scaling_stride % 4 == 0;
		if (!fused) {
			CalibrateParams cp;
			memset(&cp, 0, sizeof(cp));
			cp.num_spectra = f.num_spectra;
			cp.num_data = f.num_data;
			cp.data = f.data;
			cp.data_stride = f.data_stride;
			cp.spec = f.cal;
			CUDA_TRY(c->ws_cal.Reserve(f.num_spectra * f.data_stride * sizeof(float)));
			cp.result = c->ws_cal.As<float>();
			cp.result_stride = f.data_stride;
			if (f.cal.flag_nonfinite) {
				CUDA_TRY(c->ws_calmask.Reserve(f.num_spectra * f.mask_stride));
				cp.mask = f.mask;
				cp.mask_stride = f.mask_stride;
				cp.mask_out = c->ws_calmask.As<uint8_t>();
			}
			cudaError_t const e = LaunchCalibrate(cp, info.num_sms, stream);
			g_launch_count += 1;
			if (e != cudaSuccess) {
				SetError("LaunchCalibrate", e);
				return sakura_Status_kUnknownError;
			}
			f.data = cp.result;
			if (cp.mask_out != nullptr) {
				f.mask = cp.mask_out;
			}
			memset(&f.cal, 0, sizeof(f.cal));
		}
	}
	// ---- polynomial hot path: register-resident kernel
	// Chebyshev contexts share the polynomial kernel up to order 5: the model space is the same and
	// only the reported coefficients change basis. Beyond that the monomial normal equations the
	// kernel solves (cond > 1e9) would lose accuracy the reference's well-conditioned Chebyshev
	// solve keeps (measured: coefficients 6e-10 relative at order 5, 1e-7 at 6, 5e-6 at 7).
	if ((f.type == kFitPolynomial || (f.type == kFitChebyshev && f.K <= 6))
			&& PolyFastSupported(f.num_data, f.K, f.data_stride, f.mask_stride, f.data, f.mask,
					f.best_fit, f.residual, f.final_mask)) {
		PolyFitParams p;
		memset(&p, 0, sizeof(p));
		p.num_spectra = f.num_spectra;
		p.n = static_cast<int>(f.num_data);
		p.K = static_cast<int>(f.K);
		p.data = f.data;
		p.data_stride = f.data_stride;
		p.mask = f.mask;
		p.mask_stride = f.mask_stride;
		p.clip_threshold_sigma = f.clip;
		p.num_fitting_max = f.nfit;
		p.coeff = f.coeff;
		p.best_fit = f.best_fit;
		p.residual = f.residual;
		p.final_mask = f.final_mask;
		p.rms = f.rms;
		p.status = f.status;
		p.lsqfit_status = f.lsq;
		p.flags = f.flags;
		p.cal = f.cal;
		p.chebyshev = (f.type == kFitChebyshev) ? 1 : 0;
		if (f.handover != nullptr) {
			p.fallback_rows = f.handover;
		} else if (f.num_spectra <= static_cast<size_t>(INT32_MAX)) {
			CUDA_TRY(c->d_fallback.Reserve((f.num_spectra + 1) * sizeof(int32_t)));
			p.fallback_rows = c->d_fallback.As<int32_t>();
		}
		int launches = 0;
		char const *kernel_name = "poly_fast";
		cudaError_t e = LaunchPolyFit(p, info.num_sms, stream, &launches, &kernel_name);
		g_launch_count += launches;
		g_last_kernel = kernel_name;
		c->last_handover = launches > 1 ? p.fallback_rows : nullptr;
		if (e != cudaSuccess) {
			SetError("LaunchPolyFit", e);
			return sakura_Status_kUnknownError;
		}
		return sakura_Status_kOK;
	}
	// ---- sinusoid with many bases: 32-row tiles, tensor-core contractions; rows whose normal
	// equations are not safely positive definite are handed to the general kernel below
	int32_t const *fb_index = nullptr;
	int32_t const *fb_count = nullptr;
	bool const cheb_dmma = f.type == kFitChebyshev;
	if ((f.type == kFitSinusoid || cheb_dmma) && c->h_ext != nullptr && f.num_data % 32 == 0
			&& f.num_data >= 64
			&& f.K >= 1 && f.K <= static_cast<size_t>(kSinMaxKp) && f.data_stride % 4 == 0
			&& f.mask_stride % 8 == 0 && f.num_spectra <= static_cast<size_t>(INT32_MAX)
			&& (reinterpret_cast<uintptr_t>(f.data) % 16) == 0
			&& (reinterpret_cast<uintptr_t>(f.mask) % 8) == 0
			&& (reinterpret_cast<uintptr_t>(f.final_mask) % 8) == 0
			&& (reinterpret_cast<uintptr_t>(f.residual) % 16) == 0
			&& (reinterpret_cast<uintptr_t>(f.best_fit) % 16) == 0) {
		SinusoidFitParams sp;
		memset(&sp, 0, sizeof(sp));
		size_t const n = f.num_data;
		sp.num_spectra = f.num_spectra;
		sp.n = static_cast<int>(n);
		sp.K = static_cast<int>(f.K);
		sp.Kp = static_cast<int>(RoundUp(f.K, 8));
		sp.next = c->ext_cols;
		bool columns_ok = true;
		sp.cos_only = cheb_dmma ? 1 : 0;
		for (size_t k = 0; k < f.K; ++k) {
			int const col = f.use_idx[k];
			if (cheb_dmma) { // basis T_col: a "cosine" of wave number col
				sp.kind[k] = static_cast<int8_t>(col == 0 ? 0 : 2);
				sp.wave[k] = static_cast<int16_t>(col);
			} else {
				sp.kind[k] = static_cast<int8_t>(col == 0 ? 0 : ((col & 1) ? 1 : 2));
				sp.wave[k] = static_cast<int16_t>((col + 1) / 2);
			}
			columns_ok = columns_ok && (col >= 0 && static_cast<size_t>(col) < c->num_bases);
		}
		if (columns_ok) {
			size_t tiles = (f.num_spectra + kSinRowsPerTile - 1) / kSinRowsPerTile;
			size_t sgrid = static_cast<size_t>(info.num_sms);
			if (sgrid > tiles) {
				sgrid = tiles;
			}
			CUDA_TRY(c->d_table.Reserve(SinusoidTableDoubles(sp.n, sp.Kp) * sizeof(double)));
			CUDA_TRY(c->d_useidx.Reserve(kMaxBases * sizeof(int16_t)));
			CUDA_TRY(c->d_fallback.Reserve((f.num_spectra + 1) * sizeof(int32_t)));
			CUDA_TRY(c->ws_mask.Reserve(sgrid * kSinRowsPerTile * n));
			CUDA_TRY(c->ws_sin_residual.Reserve(sgrid * kSinRowsPerTile * n * sizeof(float)));
			if (f.best_fit != nullptr) {
				CUDA_TRY(c->ws_sin_best_fit.Reserve(sgrid * kSinRowsPerTile * n * sizeof(float)));
			}
			CUDA_TRY(cudaMemcpyAsync(c->d_useidx.ptr, f.use_idx, kMaxBases * sizeof(int16_t),
					cudaMemcpyHostToDevice, stream));
			cudaError_t e = LaunchGatherTable(c->d_basis.As<double>(), static_cast<int>(c->num_bases),
					sp.n, sp.K, sp.Kp, c->d_useidx.As<int16_t>(), c->d_table.As<double>(), stream);
			g_launch_count += 1;
			if (e != cudaSuccess) {
				SetError("LaunchGatherTable", e);
				return sakura_Status_kUnknownError;
			}
			int32_t *fb = c->d_fallback.As<int32_t>();
			CUDA_TRY(cudaMemsetAsync(fb, 0, sizeof(int32_t), stream));
			sp.ld_p = sp.Kp + 1;
			sp.ld_a = SinusoidLdA(sp.Kp);
			sp.table_p = c->d_table.As<double>();
			sp.table_a = sp.table_p + n * sp.ld_p;
			sp.ext_p = c->d_ext.As<double>();
			sp.ext_ld = SinusoidExtLd(c->ext_cols);
			sp.data = f.data;
			sp.data_stride = f.data_stride;
			sp.mask = f.mask;
			sp.mask_stride = f.mask_stride;
			sp.clip_threshold_sigma = f.clip;
			sp.num_fitting_max = f.nfit;
			sp.coeff = f.coeff;
			sp.best_fit = f.best_fit;
			sp.residual = f.residual;
			sp.final_mask = f.final_mask;
			sp.rms = f.rms;
			sp.status = f.status;
			sp.lsqfit_status = f.lsq;
			sp.flags = f.flags;
			sp.ws_residual = c->ws_sin_residual.As<float>();
			sp.ws_best_fit = f.best_fit ? c->ws_sin_best_fit.As<float>() : nullptr;
			sp.ws_mask = c->ws_mask.As<uint8_t>();
			sp.fallback_count = fb;
			sp.fallback_rows = fb + 1;
			e = LaunchSinusoidFit(sp, static_cast<int>(sgrid), stream);
			g_launch_count += 1;
			if (e != cudaSuccess) {
				SetError("LaunchSinusoidFit", e);
				return sakura_Status_kUnknownError;
			}
			fb_count = fb;
			fb_index = fb + 1;
		}
	}
	// ---- cubic spline: row-resident kernel on per-piece local moments; same hand-over of nearly
	// singular rows to the general kernel
	if (f.type == kFitCubicSpline && f.num_data <= static_cast<size_t>(kSplineMaxChannels)
			&& f.npiece <= static_cast<size_t>(kSplineMaxPieces)
			&& SplineFastSupported(static_cast<int>(f.num_data), static_cast<int>(f.npiece))
			&& f.data_stride % 4 == 0 && f.mask_stride % 4 == 0
			&& f.num_spectra <= static_cast<size_t>(INT32_MAX)
			&& (reinterpret_cast<uintptr_t>(f.data) % 16) == 0
			&& (reinterpret_cast<uintptr_t>(f.mask) % 4) == 0
			&& (reinterpret_cast<uintptr_t>(f.final_mask) % 4) == 0
			&& (reinterpret_cast<uintptr_t>(f.residual) % 16) == 0) {
		SplineFitParams sp;
		memset(&sp, 0, sizeof(sp));
		sp.num_spectra = f.num_spectra;
		sp.n = static_cast<int>(f.num_data);
		sp.npiece = static_cast<int>(f.npiece);
		sp.data = f.data;
		sp.data_stride = f.data_stride;
		sp.mask = f.mask;
		sp.mask_stride = f.mask_stride;
		sp.clip_threshold_sigma = f.clip;
		sp.num_fitting_max = f.nfit;
		sp.coeff = f.coeff;
		sp.coeff_stride = f.coeff_stride;
		sp.best_fit = f.best_fit;
		sp.residual = f.residual;
		sp.final_mask = f.final_mask;
		sp.rms = f.rms;
		sp.status = f.status;
		sp.lsqfit_status = f.lsq;
		sp.flags = f.flags;
		sp.boundary = f.boundary;
		CUDA_TRY(c->d_fallback.Reserve((f.num_spectra + 1) * sizeof(int32_t)));
		int32_t *fb = c->d_fallback.As<int32_t>();
		CUDA_TRY(cudaMemsetAsync(fb, 0, sizeof(int32_t), stream));
		sp.fallback_count = fb;
		sp.fallback_rows = fb + 1;
		size_t sgrid = static_cast<size_t>(SplineResidentCtas(sp.n, sp.npiece, info.num_sms));
		if (sgrid > f.num_spectra) {
			sgrid = f.num_spectra;
		}
		cudaError_t e = LaunchSplineFit(sp, static_cast<int>(sgrid), stream);
		g_launch_count += 1;
		g_last_kernel = "spline_moments";
		if (e != cudaSuccess) {
			SetError("LaunchSplineFit", e);
			return sakura_Status_kUnknownError;
		}
		fb_count = fb;
		fb_index = fb + 1;
	}
	// ---- general path
	size_t grid = static_cast<size_t>(info.num_sms) * 4;
	if (fb_index != nullptr) {
		grid = static_cast<size_t>(info.num_sms);
	}
	if (grid > f.num_spectra) {
		grid = f.num_spectra;
	}
	GeneralFitParams p;
	memset(&p, 0, sizeof(p));
	p.row_index = fb_index;
	p.row_count = fb_count;
	p.fit_type = f.type;
	p.num_spectra = f.num_spectra;
	p.n = static_cast<int>(f.num_data);
	p.nb_ctx = static_cast<int>(c->num_bases);
	p.K = static_cast<int>(f.K);
	p.nb_row = static_cast<int>(c->num_lsq_bases_max);
	p.npiece = static_cast<int>(f.npiece);
	p.basis = c->d_basis.As<double>();
	p.data = f.data;
	p.data_stride = f.data_stride;
	p.mask = f.mask;
	p.mask_stride = f.mask_stride;
	p.clip_threshold_sigma = f.clip;
	p.num_fitting_max = f.nfit;
	p.coeff = f.coeff;
	p.coeff_stride = f.coeff_stride;
	p.best_fit = f.best_fit;
	p.residual = f.residual;
	p.final_mask = f.final_mask;
	p.rms = f.rms;
	p.status = f.status;
	p.lsqfit_status = f.lsq;
	p.flags = f.flags;
	p.boundary = f.boundary;
	memcpy(p.use_idx, f.use_idx, sizeof(p.use_idx));
	CUDA_TRY(c->ws_residual.Reserve(grid * f.num_data * sizeof(float)));
	p.ws_residual = c->ws_residual.As<float>();
	if (f.best_fit != nullptr) {
		CUDA_TRY(c->ws_best_fit.Reserve(grid * f.num_data * sizeof(float)));
		p.ws_best_fit = c->ws_best_fit.As<float>();
	}
	if (f.type == kFitCubicSpline) {
		CUDA_TRY(c->ws_spline.Reserve(grid * f.num_data * c->num_lsq_bases_max * sizeof(double)));
		p.ws_spline_basis = c->ws_spline.As<double>();
	}
	cudaError_t e = LaunchGeneralFit(p, static_cast<int>(grid), stream);
	g_launch_count += 1;
	if (fb_index == nullptr) {
		g_last_kernel = "general";
	} else if (f.type == kFitSinusoid) {
		g_last_kernel = "sinusoid_dmma";
	} else if (f.type == kFitChebyshev) {
		g_last_kernel = "chebyshev_dmma";
	}
	if (e != cudaSuccess) {
		SetError("LaunchGeneralFit", e);
		return sakura_Status_kUnknownError;
	}
	return sakura_Status_kOK;
}


// Host-pointer batch: stage through device buffers owned by the context.
struct HostFitArgs {
	float const *data;
	bool const *mask;
	double *coeff;
	float *best_fit;
	float *residual;
	bool *final_mask;
	float *rms;
	int32_t *status;
	int32_t *lsq;
	size_t *boundary;
};

sakura_Status RunHostBatchSerial(Context *c, FitCall f, HostFitArgs const &h);

/** Row-strided copy; a single 1-D copy when both sides are tightly packed (faster on the
 *  copy engines than the pitched form). */
cudaError_t CopyRows(void *dst, size_t dpitch, void const *src, size_t spitch, size_t width,
		size_t rows, cudaMemcpyKind kind, cudaStream_t stream) {
	if (dpitch == width && spitch == width) {
		return cudaMemcpyAsync(dst, src, width * rows, kind, stream);
	}
	return cudaMemcpy2DAsync(dst, dpitch, src, spitch, width, rows, kind, stream);
}

// Pipelined variant for the polynomial hot path: the batch is cut into chunks that cycle
// through kPipelineSlots (stream, device buffers) slots, so that the host->device copy of
// chunk i+1, the kernel of chunk i and the device->host copy of chunk i-1 overlap on the
// copy engines and the SMs. With pinned host buffers every copy is truly asynchronous.
sakura_Status RunHostBatch(Context *c, FitCall f, HostFitArgs const &h) {
	if (f.num_spectra == 0) {
		return sakura_Status_kOK;
	}
	size_t const n = f.num_data;
	size_t const R = f.num_spectra;
	size_t const dstride = RoundUp(n, 32);
	bool const fast = (f.type == kFitPolynomial || (f.type == kFitChebyshev && f.K <= 6))
			&& PolyFastSupported(n, f.K, dstride, dstride,
			nullptr, nullptr, nullptr, nullptr, nullptr);
	// chunk size of the pipeline in MiB of data (tuning knob; default from measurements on B200)
	static size_t const chunk_mib = []() -> size_t {
		char const *e = getenv("SAKURA_B200_CHUNK_MIB");
		long const v = (e != nullptr) ? atol(e) : 0;
		return (v >= 1 && v <= 4096) ? static_cast<size_t>(v) : 96;
	}();
	size_t chunk_rows = (chunk_mib << 20) / (dstride * sizeof(float));
	if (chunk_rows < 1) {
		chunk_rows = 1;
	}
	if (!fast || R <= chunk_rows) {
		return RunHostBatchSerial(c, f, h);
	}
	DeviceInfo info;
	sakura_Status st = CurrentDevice(&info);
	if (st != sakura_Status_kOK) {
		return st;
	}
	if (c->device >= 0 && c->device != info.device) {
		ReleaseDeviceState(c);
	}
	c->device = info.device;
	if (c->slots == nullptr) {
		c->slots = new PipelineSlot[kPipelineSlots];
	}
	size_t const hds = f.data_stride, hms = f.mask_stride;
	size_t const num_chunks = (R + chunk_rows - 1) / chunk_rows;
	for (int s = 0; s < kPipelineSlots; ++s) {
		PipelineSlot &sl = c->slots[s];
		if (sl.stream == nullptr) {
			CUDA_TRY(cudaStreamCreateWithFlags(&sl.stream, cudaStreamNonBlocking));
			CUDA_TRY(cudaStreamCreateWithFlags(&sl.down, cudaStreamNonBlocking));
			CUDA_TRY(cudaEventCreateWithFlags(&sl.flags_ready, cudaEventDisableTiming));
			CUDA_TRY(cudaEventCreateWithFlags(&sl.down_done, cudaEventDisableTiming));
		}
		CUDA_TRY(sl.data.Reserve(chunk_rows * dstride * sizeof(float)));
		CUDA_TRY(sl.mask.Reserve(chunk_rows * dstride));
		CUDA_TRY(sl.final_mask.Reserve(chunk_rows * dstride));
		CUDA_TRY(sl.small.Reserve(chunk_rows * (sizeof(float) + 3 * sizeof(int32_t))));
		if (h.coeff != nullptr) {
			CUDA_TRY(sl.coeff.Reserve(chunk_rows * f.coeff_stride * sizeof(double)));
		}
		if (h.best_fit != nullptr) {
			CUDA_TRY(sl.best_fit.Reserve(chunk_rows * dstride * sizeof(float)));
		}
		if (h.residual != nullptr) {
			CUDA_TRY(sl.residual.Reserve(chunk_rows * dstride * sizeof(float)));
		}
		CUDA_TRY(sl.handover.Reserve((chunk_rows + 1) * sizeof(int32_t)));
		sl.h_flags.resize(chunk_rows);
		sl.rows = 0;
	}
	auto finish = [&](PipelineSlot &sl) -> sakura_Status {
		// the kernel of this slot's chunk is done and its per-row flags are on the host
		CUDA_TRY(cudaEventSynchronize(sl.flags_ready));
		size_t const r0 = sl.row0, rows = sl.rows;
		float *d_rms = sl.small.As<float>();
		bool all_published = true;
		for (size_t r = 0; r < rows; ++r) {
			if ((sl.h_flags[r] & 3) != 3) {
				all_published = false;
				break;
			}
		}
		if (all_published) {
			if (h.coeff != nullptr) {
				CUDA_TRY(cudaMemcpyAsync(h.coeff + r0 * f.coeff_stride, sl.coeff.ptr,
						rows * f.coeff_stride * sizeof(double), cudaMemcpyDeviceToHost, sl.down));
			}
			if (h.best_fit != nullptr) {
				CUDA_TRY(CopyRows(h.best_fit + r0 * hds, hds * sizeof(float), sl.best_fit.ptr,
						dstride * sizeof(float), n * sizeof(float), rows, cudaMemcpyDeviceToHost,
						sl.down));
			}
			if (h.residual != nullptr) {
				CUDA_TRY(CopyRows(h.residual + r0 * hds, hds * sizeof(float), sl.residual.ptr,
						dstride * sizeof(float), n * sizeof(float), rows, cudaMemcpyDeviceToHost,
						sl.down));
			}
			CUDA_TRY(CopyRows(h.final_mask + r0 * hms, hms, sl.final_mask.ptr, dstride, n, rows,
					cudaMemcpyDeviceToHost, sl.down));
			CUDA_TRY(cudaMemcpyAsync(h.rms + r0, d_rms, rows * sizeof(float), cudaMemcpyDeviceToHost,
					sl.down));
		} else {
			for (size_t r = 0; r < rows; ++r) {
				int const fl = sl.h_flags[r];
				if (fl & 1) {
					if (h.coeff != nullptr) {
						CUDA_TRY(cudaMemcpyAsync(h.coeff + (r0 + r) * f.coeff_stride,
								sl.coeff.As<double>() + r * f.coeff_stride,
								f.coeff_stride * sizeof(double), cudaMemcpyDeviceToHost, sl.down));
					}
					if (h.best_fit != nullptr) {
						CUDA_TRY(cudaMemcpyAsync(h.best_fit + (r0 + r) * hds,
								sl.best_fit.As<float>() + r * dstride, n * sizeof(float),
								cudaMemcpyDeviceToHost, sl.down));
					}
					if (h.residual != nullptr) {
						CUDA_TRY(cudaMemcpyAsync(h.residual + (r0 + r) * hds,
								sl.residual.As<float>() + r * dstride, n * sizeof(float),
								cudaMemcpyDeviceToHost, sl.down));
					}
					CUDA_TRY(cudaMemcpyAsync(h.rms + r0 + r, d_rms + r, sizeof(float),
							cudaMemcpyDeviceToHost, sl.down));
				}
				if (fl & 2) {
					CUDA_TRY(cudaMemcpyAsync(h.final_mask + (r0 + r) * hms,
							sl.final_mask.As<uint8_t>() + r * dstride, n, cudaMemcpyDeviceToHost,
							sl.down));
				}
			}
		}
		CUDA_TRY(cudaEventRecord(sl.down_done, sl.down));
		sl.rows = 0;
		return sakura_Status_kOK;
	};
	for (size_t ci = 0; ci < num_chunks; ++ci) {
		PipelineSlot &sl = c->slots[ci % kPipelineSlots];
		if (sl.rows != 0) {
			st = finish(sl);
			if (st != sakura_Status_kOK) {
				return st;
			}
		}
		size_t const r0 = ci * chunk_rows;
		size_t const rows = (r0 + chunk_rows <= R) ? chunk_rows : (R - r0);
		sl.row0 = r0;
		sl.rows = rows;
		float *d_rms = sl.small.As<float>();
		int32_t *d_status = reinterpret_cast<int32_t *>(d_rms + chunk_rows);
		int32_t *d_lsq = d_status + chunk_rows;
		int32_t *d_flags = d_lsq + chunk_rows;
		CUDA_TRY(CopyRows(sl.data.ptr, dstride * sizeof(float), h.data + r0 * hds,
				hds * sizeof(float), n * sizeof(float), rows, cudaMemcpyHostToDevice, sl.stream));
		CUDA_TRY(CopyRows(sl.mask.ptr, dstride, h.mask + r0 * hms, hms, n, rows,
				cudaMemcpyHostToDevice, sl.stream));
		// the kernel overwrites this slot's result buffers: wait for their previous download
		CUDA_TRY(cudaStreamWaitEvent(sl.stream, sl.down_done, 0));
		CUDA_TRY(cudaMemsetAsync(d_flags, 0, rows * sizeof(int32_t), sl.stream));
		FitCall fc = f;
		fc.num_spectra = rows;
		fc.data = sl.data.As<float>();
		fc.mask = sl.mask.As<uint8_t>();
		fc.data_stride = dstride;
		fc.mask_stride = dstride;
		fc.coeff = h.coeff ? sl.coeff.As<double>() : nullptr;
		fc.best_fit = h.best_fit ? sl.best_fit.As<float>() : nullptr;
		fc.residual = h.residual ? sl.residual.As<float>() : nullptr;
		fc.final_mask = sl.final_mask.As<uint8_t>();
		fc.rms = d_rms;
		fc.status = d_status;
		fc.lsq = d_lsq;
		fc.flags = d_flags;
		fc.boundary = nullptr;
		fc.handover = sl.handover.As<int32_t>();
		st = LaunchFit(c, fc, sl.stream);
		if (st != sakura_Status_kOK) {
			return st;
		}
		CUDA_TRY(cudaMemcpyAsync(h.status + r0, d_status, rows * sizeof(int32_t),
				cudaMemcpyDeviceToHost, sl.stream));
		CUDA_TRY(cudaMemcpyAsync(h.lsq + r0, d_lsq, rows * sizeof(int32_t), cudaMemcpyDeviceToHost,
				sl.stream));
		CUDA_TRY(cudaMemcpyAsync(sl.h_flags.data(), d_flags, rows * sizeof(int32_t),
				cudaMemcpyDeviceToHost, sl.stream));
		CUDA_TRY(cudaEventRecord(sl.flags_ready, sl.stream));
		CUDA_TRY(cudaStreamWaitEvent(sl.down, sl.flags_ready, 0));
	}
	for (size_t k = 0; k < static_cast<size_t>(kPipelineSlots); ++k) {
		PipelineSlot &sl = c->slots[(num_chunks + k) % kPipelineSlots];
		if (sl.rows != 0) {
			st = finish(sl);
			if (st != sakura_Status_kOK) {
				return st;
			}
		}
	}
	for (int s = 0; s < kPipelineSlots; ++s) {
		CUDA_TRY(cudaStreamSynchronize(c->slots[s].stream));
		CUDA_TRY(cudaStreamSynchronize(c->slots[s].down));
	}
	return sakura_Status_kOK;
}

sakura_Status RunHostBatchSerial(Context *c, FitCall f, HostFitArgs const &h) {
	if (f.num_spectra == 0) {
		return sakura_Status_kOK;
	}
	cudaStream_t const stream = 0;
	size_t const n = f.num_data;
	size_t const R = f.num_spectra;
	size_t const dstride = RoundUp(n, 32);
	size_t const hds = f.data_stride, hms = f.mask_stride;
	size_t const nb = f.npiece + 1;
	bool const spline = (f.type == kFitCubicSpline);

	CUDA_TRY(c->s_data.Reserve(R * dstride * sizeof(float)));
	CUDA_TRY(c->s_mask.Reserve(R * dstride));
	CUDA_TRY(c->s_final_mask.Reserve(R * dstride));
	CUDA_TRY(c->s_rms.Reserve(R * sizeof(float)));
	CUDA_TRY(c->s_status.Reserve(R * sizeof(int32_t)));
	CUDA_TRY(c->s_lsq.Reserve(R * sizeof(int32_t)));
	CUDA_TRY(c->s_flags.Reserve(R * sizeof(int32_t)));
	if (h.coeff != nullptr) {
		CUDA_TRY(c->s_coeff.Reserve(R * f.coeff_stride * sizeof(double)));
	}
	if (h.best_fit != nullptr) {
		CUDA_TRY(c->s_best_fit.Reserve(R * dstride * sizeof(float)));
	}
	if (h.residual != nullptr) {
		CUDA_TRY(c->s_residual.Reserve(R * dstride * sizeof(float)));
	}
	if (spline) {
		CUDA_TRY(c->s_boundary.Reserve(R * nb * sizeof(unsigned long long)));
	}
	CUDA_TRY(cudaMemcpy2DAsync(c->s_data.ptr, dstride * sizeof(float), h.data, hds * sizeof(float),
			n * sizeof(float), R, cudaMemcpyHostToDevice, stream));
	CUDA_TRY(cudaMemcpy2DAsync(c->s_mask.ptr, dstride, h.mask, hms, n, R, cudaMemcpyHostToDevice,
			stream));
	CUDA_TRY(cudaMemsetAsync(c->s_flags.ptr, 0, R * sizeof(int32_t), stream));

	f.data = c->s_data.As<float>();
	f.mask = c->s_mask.As<uint8_t>();
	f.data_stride = dstride;
	f.mask_stride = dstride;
	f.coeff = h.coeff ? c->s_coeff.As<double>() : nullptr;
	f.best_fit = h.best_fit ? c->s_best_fit.As<float>() : nullptr;
	f.residual = h.residual ? c->s_residual.As<float>() : nullptr;
	f.final_mask = c->s_final_mask.As<uint8_t>();
	f.rms = c->s_rms.As<float>();
	f.status = c->s_status.As<int32_t>();
	f.lsq = c->s_lsq.As<int32_t>();
	f.flags = c->s_flags.As<int32_t>();
	f.boundary = spline ? c->s_boundary.As<unsigned long long>() : nullptr;

	sakura_Status st = LaunchFit(c, f, stream);
	if (st != sakura_Status_kOK) {
		return st;
	}
	if (c->h_flags == nullptr) {
		c->h_flags = new std::vector<int32_t>();
	}
	std::vector<int32_t> &flags = *c->h_flags;
	flags.resize(R);
	CUDA_TRY(cudaMemcpyAsync(h.status, f.status, R * sizeof(int32_t), cudaMemcpyDeviceToHost, stream));
	CUDA_TRY(cudaMemcpyAsync(h.lsq, f.lsq, R * sizeof(int32_t), cudaMemcpyDeviceToHost, stream));
	CUDA_TRY(cudaMemcpyAsync(flags.data(), f.flags, R * sizeof(int32_t), cudaMemcpyDeviceToHost,
			stream));
	CUDA_TRY(cudaStreamSynchronize(stream));

	// Row flags (set by the kernels): bit0 = results published, bit1 = final_mask written,
	// bit2 = boundary written.
	bool all_published = true;
	for (size_t r = 0; r < R; ++r) {
		if ((flags[r] & 7) != (spline ? 7 : 3)) {
			all_published = false;
			break;
		}
	}
	if (all_published) {
		if (h.coeff != nullptr) {
			CUDA_TRY(cudaMemcpyAsync(h.coeff, f.coeff, R * f.coeff_stride * sizeof(double),
					cudaMemcpyDeviceToHost, stream));
		}
		if (h.best_fit != nullptr) {
			CUDA_TRY(cudaMemcpy2DAsync(h.best_fit, hds * sizeof(float), f.best_fit,
					dstride * sizeof(float), n * sizeof(float), R, cudaMemcpyDeviceToHost, stream));
		}
		if (h.residual != nullptr) {
			CUDA_TRY(cudaMemcpy2DAsync(h.residual, hds * sizeof(float), f.residual,
					dstride * sizeof(float), n * sizeof(float), R, cudaMemcpyDeviceToHost, stream));
		}
		CUDA_TRY(cudaMemcpy2DAsync(h.final_mask, hms, f.final_mask, dstride, n, R,
				cudaMemcpyDeviceToHost, stream));
		CUDA_TRY(cudaMemcpyAsync(h.rms, f.rms, R * sizeof(float), cudaMemcpyDeviceToHost, stream));
		if (spline) {
			CUDA_TRY(cudaMemcpyAsync(h.boundary, f.boundary, R * nb * sizeof(unsigned long long),
					cudaMemcpyDeviceToHost, stream));
		}
	} else {
		// Some rows failed or wrote nothing: copy row by row exactly what the reference
		// would have written.
		for (size_t r = 0; r < R; ++r) {
			int const fl = flags[r];
			if (fl & 1) {
				if (h.coeff != nullptr) {
					CUDA_TRY(cudaMemcpyAsync(h.coeff + r * f.coeff_stride, f.coeff + r * f.coeff_stride,
							f.coeff_stride * sizeof(double), cudaMemcpyDeviceToHost, stream));
				}
				if (h.best_fit != nullptr) {
					CUDA_TRY(cudaMemcpyAsync(h.best_fit + r * hds, f.best_fit + r * dstride,
							n * sizeof(float), cudaMemcpyDeviceToHost, stream));
				}
				if (h.residual != nullptr) {
					CUDA_TRY(cudaMemcpyAsync(h.residual + r * hds, f.residual + r * dstride,
							n * sizeof(float), cudaMemcpyDeviceToHost, stream));
				}
				CUDA_TRY(cudaMemcpyAsync(h.rms + r, f.rms + r, sizeof(float), cudaMemcpyDeviceToHost,
						stream));
			}
			if (fl & 2) {
				CUDA_TRY(cudaMemcpyAsync(h.final_mask + r * hms, f.final_mask + r * dstride, n,
						cudaMemcpyDeviceToHost, stream));
			}
			if (spline && (fl & 4)) {
				CUDA_TRY(cudaMemcpyAsync(h.boundary + r * nb, f.boundary + r * nb,
						nb * sizeof(unsigned long long), cudaMemcpyDeviceToHost, stream));
			}
		}
	}
	CUDA_TRY(cudaStreamSynchronize(stream));
	return sakura_Status_kOK;
}

// ---- argument checks shared by the single-row, batch and device entry points -----

bool RowsAligned(void const *base, size_t stride_bytes, size_t num_spectra) {
	if (!IsAligned(base)) {
		return false;
	}
	return num_spectra <= 1 || (stride_bytes % kAlignment) == 0;
}

bool DeviceRowsAligned(void const *base, size_t stride_bytes, size_t num_spectra) {
	if ((reinterpret_cast<uintptr_t>(base) % 16) != 0) {
		return false;
	}
	return num_spectra <= 1 || (stride_bytes % 16) == 0;
}

bool UniqueAscending(size_t n, size_t const *a) {
	for (size_t i = 1; i < n; ++i) {
		if (a[i - 1] >= a[i]) {
			return false;
		}
	}
	return true;
}

void FillIdentity(int16_t *use_idx) {
	for (int i = 0; i < kMaxBases; ++i) {
		use_idx[i] = static_cast<int16_t>(i);
	}
}

// baseline.cc:1239-1254; entries past the ones the wave list defines continue with
// the next unused columns (the reference leaves them unspecified).
void FillSinusoidIndex(size_t num_nwave, size_t const *nwave, size_t num_bases, int16_t *use_idx) {
	size_t iuse = 0, i = 0;
	bool used[kMaxBases + 2];
	memset(used, 0, sizeof(used));
	if (nwave[0] == 0) {
		use_idx[iuse++] = 0;
		used[0] = true;
		++i;
	}
	for (; i < num_nwave; ++i) {
		use_idx[iuse++] = static_cast<int16_t>(2 * nwave[i] - 1);
		use_idx[iuse++] = static_cast<int16_t>(2 * nwave[i]);
		used[2 * nwave[i] - 1] = true;
		used[2 * nwave[i]] = true;
	}
	for (size_t col = 0; col < num_bases && iuse < static_cast<size_t>(kMaxBases); ++col) {
		if (!used[col]) {
			use_idx[iuse++] = static_cast<int16_t>(col);
		}
	}
	for (; iuse < static_cast<size_t>(kMaxBases); ++iuse) {
		use_idx[iuse] = 0;
	}
}

template<typename PtrCheck>
sakura_Status CheckPolynomial(Context const *c, uint16_t order, size_t num_spectra, size_t num_data,
		void const *data, size_t data_stride, void const *mask, size_t mask_stride, float clip,
		size_t num_coeff, void const *coeff, void const *best_fit, void const *residual,
		void const *final_mask, void const *rms, PtrCheck rows_ok) {
	CHECK_ARGS(ValidContext(c));
	CHECK_ARGS(c->type == kFitPolynomial || c->type == kFitChebyshev); // baseline.cc:1699-1701
	CHECK_ARGS(order <= c->param); // :1702
	CHECK_ARGS(num_data == c->num_data); // :1703
	CHECK_ARGS(data != nullptr && mask != nullptr);
	CHECK_ARGS(num_spectra <= 1 || (data_stride >= num_data && mask_stride >= num_data));
	CHECK_ARGS(rows_ok(data, data_stride * sizeof(float)));
	CHECK_ARGS(rows_ok(mask, mask_stride));
	CHECK_ARGS(0.0f < clip); // :1708
	if (coeff != nullptr) {
		CHECK_ARGS(num_coeff == static_cast<size_t>(order) + 1); // :1710-1712
		CHECK_ARGS(rows_ok(coeff, kAlignment));
	}
	if (best_fit != nullptr) {
		CHECK_ARGS(rows_ok(best_fit, data_stride * sizeof(float)));
	}
	if (residual != nullptr) {
		CHECK_ARGS(rows_ok(residual, data_stride * sizeof(float)));
	}
	CHECK_ARGS(final_mask != nullptr);
	CHECK_ARGS(rows_ok(final_mask, mask_stride));
	CHECK_ARGS(rms != nullptr);
	// num_coeff selects the number of fitted bases (baseline.cc:1301-1302); it must address
	// columns that exist in the context table.
	CHECK_ARGS(num_coeff <= c->num_bases);
	CHECK_ARGS(num_coeff <= num_data);
	return sakura_Status_kOK;
}

} // namespace

// =============================================================================
// extern "C" entry points (C linkage comes from the declarations in sakura_b200.h)
// =============================================================================

sakura_Status sakura_Initialize(sakura_UserAllocator allocator, sakura_UserDeallocator deallocator)
		noexcept {
	g_allocator = allocator == nullptr ? DefaultAllocator : allocator;
	g_deallocator = deallocator == nullptr ? DefaultFree : deallocator;
	return sakura_Status_kOK;
}

void sakura_CleanUp(void) noexcept {
}

size_t sakura_GetAlignment(void) noexcept {
	return kAlignment;
}

bool sakura_IsAligned(void const *ptr) noexcept {
	return IsAligned(ptr);
}

void *sakura_AlignAny(size_t size_of_arena, void *vp, size_t size_required) noexcept {
	if (vp == nullptr) {
		return nullptr;
	}
	uintptr_t const addr = reinterpret_cast<uintptr_t>(vp);
	uintptr_t const pad = kAlignment - 1u;
	uintptr_t const new_addr = (addr + pad) & ~pad;
	if (size_of_arena < size_required + static_cast<size_t>(new_addr - addr)) {
		return nullptr;
	}
	return reinterpret_cast<void *>(new_addr);
}

float *sakura_AlignFloat(size_t elements_in_arena, float *fp, size_t elements_required) noexcept {
	return static_cast<float *>(sakura_AlignAny(elements_in_arena * sizeof(float), fp,
			elements_required * sizeof(float)));
}

double *sakura_AlignDouble(size_t elements_in_arena, double *dp, size_t elements_required) noexcept {
	return static_cast<double *>(sakura_AlignAny(elements_in_arena * sizeof(double), dp,
			elements_required * sizeof(double)));
}

uint64_t sakura_b200_GetKernelLaunchCount(void) noexcept {
	return g_launch_count.load();
}

char const *sakura_b200_GetLastKernelName(void) noexcept {
	return g_last_kernel;
}

int64_t sakura_b200_GetLastHandOverCount(struct sakura_LSQFitContextFloat const *context) noexcept {
	if (!ValidContext(context) || context->last_handover == nullptr) {
		return -1;
	}
	int32_t v = 0;
	if (cudaDeviceSynchronize() != cudaSuccess
			|| cudaMemcpy(&v, context->last_handover, sizeof(v), cudaMemcpyDeviceToHost) != cudaSuccess) {
		return -1;
	}
	return v;
}

char const *sakura_b200_GetLastErrorMessage(void) noexcept {
	return g_last_error;
}

char const *sakura_b200_GetVersionString(void) noexcept {
	return "sakura_b200 0.1 (sm_100a; libsakura 5.3 LSQ-fit path)";
}

// ---------------------------------------------------------------- contexts

sakura_Status sakura_CreateLSQFitContextPolynomialFloat(sakura_LSQFitType const lsqfit_type,
		uint16_t order, size_t num_data, struct sakura_LSQFitContextFloat **context) noexcept {
	CHECK_ARGS(context != nullptr);
	*context = nullptr;
	CHECK_ARGS(lsqfit_type == sakura_LSQFitType_kPolynomial
			|| lsqfit_type == sakura_LSQFitType_kChebyshev);
	int const type = (lsqfit_type == sakura_LSQFitType_kChebyshev) ? kFitChebyshev : kFitPolynomial;
	size_t const num_lsq_bases = LsqBases(type, order);
	CHECK_ARGS(num_lsq_bases <= num_data); // baseline.cc:1529
	CHECK_ARGS(num_lsq_bases <= SIZE_MAX / num_data);
	return CreateContext(type, order, num_data, context);
}

sakura_Status sakura_CreateLSQFitContextCubicSplineFloat(uint16_t npiece, size_t num_data,
		struct sakura_LSQFitContextFloat **context) noexcept {
	CHECK_ARGS(context != nullptr);
	*context = nullptr;
	CHECK_ARGS(0 < npiece); // baseline.cc:1560
	size_t const num_lsq_bases = LsqBases(kFitCubicSpline, npiece);
	CHECK_ARGS(num_lsq_bases <= num_data);
	CHECK_ARGS(num_lsq_bases <= SIZE_MAX / num_data);
	return CreateContext(kFitCubicSpline, npiece, num_data, context);
}

sakura_Status sakura_CreateLSQFitContextSinusoidFloat(uint16_t nwave, size_t num_data,
		struct sakura_LSQFitContextFloat **context) noexcept {
	CHECK_ARGS(context != nullptr);
	*context = nullptr;
	size_t const num_lsq_bases = LsqBases(kFitSinusoid, nwave);
	CHECK_ARGS(num_lsq_bases + 1 <= num_data); // baseline.cc:1601-1602
	CHECK_ARGS(num_lsq_bases + 1 <= SIZE_MAX / num_data);
	return CreateContext(kFitSinusoid, nwave, num_data, context);
}

sakura_Status sakura_DestroyLSQFitContextFloat(struct sakura_LSQFitContextFloat *context) noexcept {
	CHECK_ARGS(context != nullptr);
	CHECK_ARGS(context->magic == kContextMagic);
	ReleaseDeviceState(context);
	delete context->h_flags;
	delete context->h_ext;
	delete[] context->slots;
	context->slots = nullptr;
	context->h_flags = nullptr;
	if (context->h_basis != nullptr) {
		g_deallocator(context->h_basis);
	}
	context->magic = 0;
	context->~sakura_LSQFitContextFloat();
	g_deallocator(context);
	return sakura_Status_kOK;
}

sakura_Status sakura_GetNumberOfCoefficientsFloat(struct sakura_LSQFitContextFloat const *context,
		uint16_t order, size_t *num_coeff) noexcept {
	CHECK_ARGS(ValidContext(context));
	int const type = context->type;
	CHECK_ARGS(type == kFitPolynomial || type == kFitChebyshev || type == kFitCubicSpline);
	if (type == kFitCubicSpline) {
		CHECK_ARGS(0 < order);
	}
	CHECK_ARGS(order <= context->param);
	CHECK_ARGS(num_coeff != nullptr);
	*num_coeff = (type == kFitCubicSpline) ? 4 * static_cast<size_t>(order) :
			static_cast<size_t>(order) + 1; // baseline.cc:203-230
	return sakura_Status_kOK;
}

// ---------------------------------------------------------------- polynomial

sakura_Status sakura_LSQFitPolynomialFloatBatch(struct sakura_LSQFitContextFloat const *context,
		uint16_t order, size_t num_spectra, size_t num_data, float const data[], size_t data_stride,
		bool const mask[], size_t mask_stride, float clip_threshold_sigma, uint16_t num_fitting_max,
		size_t num_coeff, double coeff[], float best_fit[], float residual[], bool final_mask[],
		float rms[], int32_t status[], int32_t lsqfit_status[]) noexcept {
	CHECK_ARGS(status != nullptr && lsqfit_status != nullptr);
	auto rows_ok = [&](void const *p, size_t stride_bytes) {
		return RowsAligned(p, stride_bytes, num_spectra);
	};
	sakura_Status st = CheckPolynomial(context, order, num_spectra, num_data, data, data_stride, mask,
			mask_stride, clip_threshold_sigma, num_coeff, coeff, best_fit, residual, final_mask, rms,
			rows_ok);
	if (st != sakura_Status_kOK) {
		return st;
	}
	Context *c = const_cast<Context *>(context);
	FitCall f;
	memset(&f, 0, sizeof(f));
	f.type = c->type;
	f.num_spectra = num_spectra;
	f.num_data = num_data;
	f.data_stride = data_stride;
	f.mask_stride = mask_stride;
	f.clip = clip_threshold_sigma;
	f.nfit = num_fitting_max;
	f.K = num_coeff;
	f.coeff_stride = num_coeff;
	FillIdentity(f.use_idx);
	HostFitArgs h = { data, mask, coeff, best_fit, residual, final_mask, rms, status, lsqfit_status,
			nullptr };
	return RunHostBatch(c, f, h);
}

namespace {
sakura_Status PolynomialBatchDevice(struct sakura_LSQFitContextFloat const *context, uint16_t order,
		size_t num_spectra, size_t num_data, float const d_data[], size_t data_stride,
		bool const d_mask[], size_t mask_stride, float clip_threshold_sigma,
		uint16_t num_fitting_max, size_t num_coeff, double d_coeff[], float d_best_fit[],
		float d_residual[], bool d_final_mask[], float d_rms[], int32_t d_status[],
		int32_t d_lsqfit_status[], void *cuda_stream, CalibrationSpec const *cal) {
	CHECK_ARGS(d_status != nullptr && d_lsqfit_status != nullptr);
	auto rows_ok = [&](void const *p, size_t stride_bytes) {
		return DeviceRowsAligned(p, stride_bytes, num_spectra);
	};
	sakura_Status st = CheckPolynomial(context, order, num_spectra, num_data, d_data, data_stride,
			d_mask, mask_stride, clip_threshold_sigma, num_coeff, d_coeff, d_best_fit, d_residual,
			d_final_mask, d_rms, rows_ok);
	if (st != sakura_Status_kOK) {
		return st;
	}
	CHECK_ARGS(static_cast<void const *>(d_best_fit) != static_cast<void const *>(d_data)
			&& static_cast<void const *>(d_residual) != static_cast<void const *>(d_data));
	Context *c = const_cast<Context *>(context);
	FitCall f;
	memset(&f, 0, sizeof(f));
	f.type = c->type;
	f.num_spectra = num_spectra;
	f.num_data = num_data;
	f.data_stride = data_stride;
	f.mask_stride = mask_stride;
	f.clip = clip_threshold_sigma;
	f.nfit = num_fitting_max;
	f.K = num_coeff;
	f.coeff_stride = num_coeff;
	FillIdentity(f.use_idx);
	f.data = d_data;
	f.mask = reinterpret_cast<uint8_t const *>(d_mask);
	f.coeff = d_coeff;
	f.best_fit = d_best_fit;
	f.residual = d_residual;
	f.final_mask = reinterpret_cast<uint8_t *>(d_final_mask);
	f.rms = d_rms;
	f.status = d_status;
	f.lsq = d_lsqfit_status;
	f.flags = nullptr;
	if (cal != nullptr) {
		f.cal = *cal;
	}
	return LaunchFit(c, f, static_cast<cudaStream_t>(cuda_stream));
}
} // namespace

sakura_Status sakura_LSQFitPolynomialFloatBatchDevice(
		struct sakura_LSQFitContextFloat const *context, uint16_t order, size_t num_spectra,
		size_t num_data, float const d_data[], size_t data_stride, bool const d_mask[],
		size_t mask_stride, float clip_threshold_sigma, uint16_t num_fitting_max, size_t num_coeff,
		double d_coeff[], float d_best_fit[], float d_residual[], bool d_final_mask[], float d_rms[],
		int32_t d_status[], int32_t d_lsqfit_status[], void *cuda_stream) noexcept {
	return PolynomialBatchDevice(context, order, num_spectra, num_data, d_data, data_stride, d_mask,
			mask_stride, clip_threshold_sigma, num_fitting_max, num_coeff, d_coeff, d_best_fit,
			d_residual, d_final_mask, d_rms, d_status, d_lsqfit_status, cuda_stream, nullptr);
}

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
namespace {
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

sakura_Status sakura_CalibrateAndLSQFitPolynomialFloatBatchDevice(
		struct sakura_LSQFitContextFloat const *context, uint16_t order, size_t num_spectra,
		size_t num_data, float const d_scaling_factor[], size_t scaling_stride,
		float const_scaling_factor, float const d_data[], size_t data_stride,
		float const d_reference[], size_t reference_stride, bool flag_nan_or_inf,
		bool const d_mask[], size_t mask_stride, float clip_threshold_sigma,
		uint16_t num_fitting_max, size_t num_coeff, double d_coeff[], float d_best_fit[],
		float d_residual[], bool d_final_mask[], float d_rms[], int32_t d_status[],
		int32_t d_lsqfit_status[], void *cuda_stream) noexcept {
	sakura_Status const st = CheckCalibration(num_spectra, num_data, d_scaling_factor,
			scaling_stride, d_data, data_stride, d_reference, reference_stride);
	if (st != sakura_Status_kOK) {
		return st;
	}
	CalibrationSpec cal;
	memset(&cal, 0, sizeof(cal));
	cal.reference = d_reference;
	cal.reference_stride = reference_stride;
	cal.scaling = d_scaling_factor;
	cal.scaling_stride = scaling_stride;
	cal.const_scaling = const_scaling_factor;
	cal.flag_nonfinite = flag_nan_or_inf ? 1 : 0;
	return PolynomialBatchDevice(context, order, num_spectra, num_data, d_data, data_stride, d_mask,
			mask_stride, clip_threshold_sigma, num_fitting_max, num_coeff, d_coeff, d_best_fit,
			d_residual, d_final_mask, d_rms, d_status, d_lsqfit_status, cuda_stream, &cal);
}

sakura_Status sakura_LSQFitPolynomialFloat(struct sakura_LSQFitContextFloat const *context,
		uint16_t order, size_t num_data, float const data[], bool const mask[],
		float clip_threshold_sigma, uint16_t num_fitting_max, size_t num_coeff, double coeff[],
		float best_fit[], float residual[], bool final_mask[], float *rms,
		sakura_LSQFitStatus *lsqfit_status) noexcept {
	CHECK_ARGS(lsqfit_status != nullptr);
	*lsqfit_status = sakura_LSQFitStatus_kNG; // baseline.cc:1696
	int32_t st_row = sakura_Status_kNG, lsq_row = sakura_LSQFitStatus_kNG;
	sakura_Status st = sakura_LSQFitPolynomialFloatBatch(context, order, 1, num_data, data, num_data,
			mask, num_data, clip_threshold_sigma, num_fitting_max, num_coeff, coeff, best_fit,
			residual, final_mask, rms, &st_row, &lsq_row);
	if (st != sakura_Status_kOK) {
		return st;
	}
	*lsqfit_status = static_cast<sakura_LSQFitStatus>(lsq_row);
	return static_cast<sakura_Status>(st_row);
}

// ---------------------------------------------------------------- cubic spline

namespace {
template<typename PtrCheck>
sakura_Status CheckSpline(Context const *c, size_t num_pieces, size_t num_spectra, size_t num_data,
		void const *data, size_t data_stride, void const *mask, size_t mask_stride, float clip,
		void const *coeff, void const *best_fit, void const *residual, void const *final_mask,
		void const *rms, void const *boundary, PtrCheck rows_ok) {
	CHECK_ARGS(ValidContext(c));
	CHECK_ARGS(c->type == kFitCubicSpline); // baseline.cc:1774
	CHECK_ARGS(0 < num_pieces);
	CHECK_ARGS(num_pieces <= c->param);
	CHECK_ARGS(num_pieces + 3 <= num_data);
	CHECK_ARGS(num_data == c->num_data);
	CHECK_ARGS(data != nullptr && mask != nullptr);
	CHECK_ARGS(num_spectra <= 1 || (data_stride >= num_data && mask_stride >= num_data));
	CHECK_ARGS(rows_ok(data, data_stride * sizeof(float)));
	CHECK_ARGS(rows_ok(mask, mask_stride));
	CHECK_ARGS(0.0f < clip);
	if (coeff != nullptr) {
		CHECK_ARGS(rows_ok(coeff, num_pieces * 4 * sizeof(double)));
	}
	if (best_fit != nullptr) {
		CHECK_ARGS(rows_ok(best_fit, data_stride * sizeof(float)));
	}
	if (residual != nullptr) {
		CHECK_ARGS(rows_ok(residual, data_stride * sizeof(float)));
	}
	CHECK_ARGS(final_mask != nullptr);
	CHECK_ARGS(rows_ok(final_mask, mask_stride));
	CHECK_ARGS(rms != nullptr);
	CHECK_ARGS(boundary != nullptr);
	CHECK_ARGS(rows_ok(boundary, kAlignment));
	return sakura_Status_kOK;
}

void SetupSplineCall(Context *c, FitCall *f, size_t num_pieces, size_t num_spectra, size_t num_data,
		size_t data_stride, size_t mask_stride, float clip, uint16_t nfit) {
	memset(f, 0, sizeof(*f));
	f->type = kFitCubicSpline;
	f->num_spectra = num_spectra;
	f->num_data = num_data;
	f->data_stride = data_stride;
	f->mask_stride = mask_stride;
	f->clip = clip;
	f->nfit = nfit;
	f->K = num_pieces + 3;
	f->npiece = num_pieces;
	f->coeff_stride = num_pieces * 4;
	FillIdentity(f->use_idx);
	(void) c;
}
} // namespace

sakura_Status sakura_LSQFitCubicSplineFloatBatch(struct sakura_LSQFitContextFloat const *context,
		size_t num_pieces, size_t num_spectra, size_t num_data, float const data[], size_t data_stride,
		bool const mask[], size_t mask_stride, float clip_threshold_sigma, uint16_t num_fitting_max,
		double coeff[], float best_fit[], float residual[], bool final_mask[], float rms[],
		size_t boundary[], int32_t status[], int32_t lsqfit_status[]) noexcept {
	CHECK_ARGS(status != nullptr && lsqfit_status != nullptr);
	sakura_Status st = CheckSpline(context, num_pieces, num_spectra, num_data, data, data_stride,
			mask, mask_stride, clip_threshold_sigma, coeff, best_fit, residual, final_mask, rms,
			boundary, [&](void const *p, size_t sb) {
				return RowsAligned(p, sb, num_spectra);
			});
	if (st != sakura_Status_kOK) {
		return st;
	}
	Context *c = const_cast<Context *>(context);
	FitCall f;
	SetupSplineCall(c, &f, num_pieces, num_spectra, num_data, data_stride, mask_stride,
			clip_threshold_sigma, num_fitting_max);
	HostFitArgs h = { data, mask, coeff, best_fit, residual, final_mask, rms, status, lsqfit_status,
			boundary };
	return RunHostBatch(c, f, h);
}

sakura_Status sakura_LSQFitCubicSplineFloatBatchDevice(
		struct sakura_LSQFitContextFloat const *context, size_t num_pieces, size_t num_spectra,
		size_t num_data, float const d_data[], size_t data_stride, bool const d_mask[],
		size_t mask_stride, float clip_threshold_sigma, uint16_t num_fitting_max, double d_coeff[],
		float d_best_fit[], float d_residual[], bool d_final_mask[], float d_rms[], size_t d_boundary[],
		int32_t d_status[], int32_t d_lsqfit_status[], void *cuda_stream) noexcept {
	CHECK_ARGS(d_status != nullptr && d_lsqfit_status != nullptr);
	sakura_Status st = CheckSpline(context, num_pieces, num_spectra, num_data, d_data, data_stride,
			d_mask, mask_stride, clip_threshold_sigma, d_coeff, d_best_fit, d_residual, d_final_mask,
			d_rms, d_boundary, [&](void const *p, size_t sb) {
				return DeviceRowsAligned(p, sb, num_spectra);
			});
	if (st != sakura_Status_kOK) {
		return st;
	}
	CHECK_ARGS(static_cast<void const *>(d_best_fit) != static_cast<void const *>(d_data)
			&& static_cast<void const *>(d_residual) != static_cast<void const *>(d_data));
	Context *c = const_cast<Context *>(context);
	FitCall f;
	SetupSplineCall(c, &f, num_pieces, num_spectra, num_data, data_stride, mask_stride,
			clip_threshold_sigma, num_fitting_max);
	f.data = d_data;
	f.mask = reinterpret_cast<uint8_t const *>(d_mask);
	f.coeff = d_coeff;
	f.best_fit = d_best_fit;
	f.residual = d_residual;
	f.final_mask = reinterpret_cast<uint8_t *>(d_final_mask);
	f.rms = d_rms;
	f.status = d_status;
	f.lsq = d_lsqfit_status;
	f.boundary = reinterpret_cast<unsigned long long *>(d_boundary);
	return LaunchFit(c, f, static_cast<cudaStream_t>(cuda_stream));
}

sakura_Status sakura_LSQFitCubicSplineFloat(struct sakura_LSQFitContextFloat const *context,
		size_t num_pieces, size_t num_data, float const data[], bool const mask[],
		float clip_threshold_sigma, uint16_t num_fitting_max, double coeff[][4], float best_fit[],
		float residual[], bool final_mask[], float *rms, size_t boundary[],
		sakura_LSQFitStatus *lsqfit_status) noexcept {
	CHECK_ARGS(lsqfit_status != nullptr);
	*lsqfit_status = sakura_LSQFitStatus_kNG; // baseline.cc:1771
	int32_t st_row = sakura_Status_kNG, lsq_row = sakura_LSQFitStatus_kNG;
	sakura_Status st = sakura_LSQFitCubicSplineFloatBatch(context, num_pieces, 1, num_data, data,
			num_data, mask, num_data, clip_threshold_sigma, num_fitting_max,
			reinterpret_cast<double *>(coeff), best_fit, residual, final_mask, rms, boundary, &st_row,
			&lsq_row);
	if (st != sakura_Status_kOK) {
		return st;
	}
	*lsqfit_status = static_cast<sakura_LSQFitStatus>(lsq_row);
	return static_cast<sakura_Status>(st_row);
}

// ---------------------------------------------------------------- sinusoid

namespace {
template<typename PtrCheck>
sakura_Status CheckSinusoid(Context const *c, size_t num_nwave, size_t const *nwave,
		size_t num_spectra, size_t num_data, void const *data, size_t data_stride, void const *mask,
		size_t mask_stride, float clip, size_t num_coeff, void const *coeff, void const *best_fit,
		void const *residual, void const *final_mask, void const *rms, PtrCheck rows_ok) {
	CHECK_ARGS(ValidContext(c));
	CHECK_ARGS(c->type == kFitSinusoid); // baseline.cc:1848
	CHECK_ARGS(0 < num_nwave);
	CHECK_ARGS(nwave != nullptr);
	CHECK_ARGS(UniqueAscending(num_nwave, nwave));
	CHECK_ARGS(nwave[num_nwave - 1] <= c->param);
	CHECK_ARGS(num_data == c->num_data);
	CHECK_ARGS(data != nullptr && mask != nullptr);
	CHECK_ARGS(num_spectra <= 1 || (data_stride >= num_data && mask_stride >= num_data));
	CHECK_ARGS(rows_ok(data, data_stride * sizeof(float)));
	CHECK_ARGS(rows_ok(mask, mask_stride));
	CHECK_ARGS(0.0f < clip);
	size_t const needed = (nwave[0] == 0) ? (2 * num_nwave - 1) : (2 * num_nwave);
	CHECK_ARGS(needed <= num_coeff); // baseline.cc:1859-1861
	CHECK_ARGS(num_coeff <= c->num_bases); // :1862
	if (coeff != nullptr) {
		CHECK_ARGS(rows_ok(coeff, kAlignment));
	}
	if (best_fit != nullptr) {
		CHECK_ARGS(rows_ok(best_fit, data_stride * sizeof(float)));
	}
	if (residual != nullptr) {
		CHECK_ARGS(rows_ok(residual, data_stride * sizeof(float)));
	}
	CHECK_ARGS(final_mask != nullptr);
	CHECK_ARGS(rows_ok(final_mask, mask_stride));
	CHECK_ARGS(rms != nullptr);
	return sakura_Status_kOK;
}

void SetupSinusoidCall(Context *c, FitCall *f, size_t num_nwave, size_t const *nwave,
		size_t num_spectra, size_t num_data, size_t data_stride, size_t mask_stride, float clip,
		uint16_t nfit, size_t num_coeff) {
	memset(f, 0, sizeof(*f));
	f->type = kFitSinusoid;
	f->num_spectra = num_spectra;
	f->num_data = num_data;
	f->data_stride = data_stride;
	f->mask_stride = mask_stride;
	f->clip = clip;
	f->nfit = nfit;
	f->K = num_coeff;
	f->coeff_stride = num_coeff;
	FillSinusoidIndex(num_nwave, nwave, c->num_bases, f->use_idx);
}
} // namespace

sakura_Status sakura_LSQFitSinusoidFloatBatch(struct sakura_LSQFitContextFloat const *context,
		size_t num_nwave, size_t const nwave[], size_t num_spectra, size_t num_data,
		float const data[], size_t data_stride, bool const mask[], size_t mask_stride,
		float clip_threshold_sigma, uint16_t num_fitting_max, size_t num_coeff, double coeff[],
		float best_fit[], float residual[], bool final_mask[], float rms[], int32_t status[],
		int32_t lsqfit_status[]) noexcept {
	CHECK_ARGS(status != nullptr && lsqfit_status != nullptr);
	sakura_Status st = CheckSinusoid(context, num_nwave, nwave, num_spectra, num_data, data,
			data_stride, mask, mask_stride, clip_threshold_sigma, num_coeff, coeff, best_fit,
			residual, final_mask, rms, [&](void const *p, size_t sb) {
				return RowsAligned(p, sb, num_spectra);
			});
	if (st != sakura_Status_kOK) {
		return st;
	}
	Context *c = const_cast<Context *>(context);
	FitCall f;
	SetupSinusoidCall(c, &f, num_nwave, nwave, num_spectra, num_data, data_stride, mask_stride,
			clip_threshold_sigma, num_fitting_max, num_coeff);
	HostFitArgs h = { data, mask, coeff, best_fit, residual, final_mask, rms, status, lsqfit_status,
			nullptr };
	return RunHostBatch(c, f, h);
}

sakura_Status sakura_LSQFitSinusoidFloatBatchDevice(struct sakura_LSQFitContextFloat const *context,
		size_t num_nwave, size_t const nwave[], size_t num_spectra, size_t num_data,
		float const d_data[], size_t data_stride, bool const d_mask[], size_t mask_stride,
		float clip_threshold_sigma, uint16_t num_fitting_max, size_t num_coeff, double d_coeff[],
		float d_best_fit[], float d_residual[], bool d_final_mask[], float d_rms[],
		int32_t d_status[], int32_t d_lsqfit_status[], void *cuda_stream) noexcept {
	CHECK_ARGS(d_status != nullptr && d_lsqfit_status != nullptr);
	sakura_Status st = CheckSinusoid(context, num_nwave, nwave, num_spectra, num_data, d_data,
			data_stride, d_mask, mask_stride, clip_threshold_sigma, num_coeff, d_coeff, d_best_fit,
			d_residual, d_final_mask, d_rms, [&](void const *p, size_t sb) {
				return DeviceRowsAligned(p, sb, num_spectra);
			});
	if (st != sakura_Status_kOK) {
		return st;
	}
	CHECK_ARGS(static_cast<void const *>(d_best_fit) != static_cast<void const *>(d_data)
			&& static_cast<void const *>(d_residual) != static_cast<void const *>(d_data));
	Context *c = const_cast<Context *>(context);
	FitCall f;
	SetupSinusoidCall(c, &f, num_nwave, nwave, num_spectra, num_data, data_stride, mask_stride,
			clip_threshold_sigma, num_fitting_max, num_coeff);
	f.data = d_data;
	f.mask = reinterpret_cast<uint8_t const *>(d_mask);
	f.coeff = d_coeff;
	f.best_fit = d_best_fit;
	f.residual = d_residual;
	f.final_mask = reinterpret_cast<uint8_t *>(d_final_mask);
	f.rms = d_rms;
	f.status = d_status;
	f.lsq = d_lsqfit_status;
	return LaunchFit(c, f, static_cast<cudaStream_t>(cuda_stream));
}

sakura_Status sakura_LSQFitSinusoidFloat(struct sakura_LSQFitContextFloat const *context,
		size_t num_nwave, size_t const nwave[], size_t num_data, float const data[],
		bool const mask[], float clip_threshold_sigma, uint16_t num_fitting_max, size_t num_coeff,
		double coeff[], float best_fit[], float residual[], bool final_mask[], float *rms,
		sakura_LSQFitStatus *lsqfit_status) noexcept {
	CHECK_ARGS(lsqfit_status != nullptr);
	*lsqfit_status = sakura_LSQFitStatus_kNG; // baseline.cc:1845
	int32_t st_row = sakura_Status_kNG, lsq_row = sakura_LSQFitStatus_kNG;
	sakura_Status st = sakura_LSQFitSinusoidFloatBatch(context, num_nwave, nwave, 1, num_data, data,
			num_data, mask, num_data, clip_threshold_sigma, num_fitting_max, num_coeff, coeff,
			best_fit, residual, final_mask, rms, &st_row, &lsq_row);
	if (st != sakura_Status_kOK) {
		return st;
	}
	*lsqfit_status = static_cast<sakura_LSQFitStatus>(lsq_row);
	return static_cast<sakura_Status>(st_row);
}

// ---------------------------------------------------------------- Subtract*

namespace {
sakura_Status RunSubtract(Context *c, int type, size_t num_data, float const *data, size_t K,
		size_t npiece, std::vector<double> const &coeff_apply, size_t const *boundary,
		int16_t const *use_idx, float *out) {
	DeviceInfo info;
	sakura_Status st = CurrentDevice(&info);
	if (st != sakura_Status_kOK) {
		return st;
	}
	cudaStream_t const stream = 0;
	st = BindDevice(c, info, stream);
	if (st != sakura_Status_kOK) {
		return st;
	}
	size_t const n = num_data;
	CUDA_TRY(c->s_data.Reserve(n * sizeof(float)));
	CUDA_TRY(c->s_residual.Reserve(n * sizeof(float)));
	CUDA_TRY(c->s_coeff.Reserve(coeff_apply.size() * sizeof(double)));
	CUDA_TRY(cudaMemcpyAsync(c->s_data.ptr, data, n * sizeof(float), cudaMemcpyHostToDevice, stream));
	CUDA_TRY(cudaMemcpyAsync(c->s_coeff.ptr, coeff_apply.data(), coeff_apply.size() * sizeof(double),
			cudaMemcpyHostToDevice, stream));
	SubtractParams p;
	memset(&p, 0, sizeof(p));
	p.fit_type = type;
	p.num_spectra = 1;
	p.n = static_cast<int>(n);
	p.nb_ctx = static_cast<int>(c->num_bases);
	p.K = static_cast<int>(K);
	p.npiece = static_cast<int>(npiece);
	p.basis = c->d_basis.As<double>();
	p.data = c->s_data.As<float>();
	p.data_stride = n;
	p.coeff = c->s_coeff.As<double>();
	p.coeff_stride = coeff_apply.size();
	p.out = c->s_residual.As<float>();
	if (type == kFitCubicSpline) {
		CUDA_TRY(c->s_boundary.Reserve((npiece + 1) * sizeof(unsigned long long)));
		CUDA_TRY(cudaMemcpyAsync(c->s_boundary.ptr, boundary, (npiece + 1) * sizeof(size_t),
				cudaMemcpyHostToDevice, stream));
		p.boundary = c->s_boundary.As<unsigned long long>();
	}
	memcpy(p.use_idx, use_idx, sizeof(p.use_idx));
	cudaError_t e = LaunchSubtract(p, 1, stream);
	g_launch_count += 1;
	g_last_kernel = "subtract";
	if (e != cudaSuccess) {
		SetError("LaunchSubtract", e);
		return sakura_Status_kUnknownError;
	}
	CUDA_TRY(cudaMemcpyAsync(out, p.out, n * sizeof(float), cudaMemcpyDeviceToHost, stream));
	CUDA_TRY(cudaStreamSynchronize(stream));
	return sakura_Status_kOK;
}
} // namespace

sakura_Status sakura_SubtractPolynomialFloat(struct sakura_LSQFitContextFloat const *context,
		size_t num_data, float const data[], size_t num_coeff, double const coeff[], float out[])
				noexcept {
	CHECK_ARGS(ValidContext(context));
	int const type = context->type;
	CHECK_ARGS(type == kFitPolynomial || type == kFitChebyshev);
	CHECK_ARGS(num_data == context->num_data);
	CHECK_ARGS(data != nullptr && IsAligned(data));
	CHECK_ARGS(0 < num_coeff);
	CHECK_ARGS(num_coeff <= context->num_bases);
	CHECK_ARGS(coeff != nullptr && IsAligned(coeff));
	CHECK_ARGS(out != nullptr && IsAligned(out));
	std::vector<double> apply(coeff, coeff + num_coeff);
	if (type == kFitPolynomial) { // baseline.cc:1444-1451
		double const max_x = static_cast<double>(num_data - 1);
		double factor = 1.0;
		for (size_t i = 0; i < num_coeff; ++i) {
			apply[i] *= factor;
			factor *= max_x;
		}
	}
	int16_t use_idx[kMaxBases];
	FillIdentity(use_idx);
	return RunSubtract(const_cast<Context *>(context), type, num_data, data, num_coeff, 0, apply,
			nullptr, use_idx, out);
}

sakura_Status sakura_SubtractCubicSplineFloat(struct sakura_LSQFitContextFloat const *context,
		size_t num_data, float const data[], size_t num_pieces, double const coeff[][4],
		size_t const boundary[], float out[]) noexcept {
	CHECK_ARGS(ValidContext(context));
	CHECK_ARGS(context->type == kFitCubicSpline);
	CHECK_ARGS(0 < num_pieces);
	CHECK_ARGS(num_pieces <= context->param);
	CHECK_ARGS(num_data == context->num_data);
	CHECK_ARGS(data != nullptr && IsAligned(data));
	CHECK_ARGS(coeff != nullptr && IsAligned(coeff));
	CHECK_ARGS(boundary != nullptr && IsAligned(boundary));
	CHECK_ARGS(boundary[0] == 0);
	CHECK_ARGS(boundary[num_pieces] == num_data);
	CHECK_ARGS(out != nullptr && IsAligned(out));
	std::vector<double> apply(num_pieces * 4);
	double const max_x = static_cast<double>(num_data - 1);
	for (size_t i = 0; i < num_pieces; ++i) { // baseline.cc:1490-1498
		double factor = 1.0;
		for (size_t j = 0; j < 4; ++j) {
			apply[4 * i + j] = coeff[i][j] * factor;
			factor *= max_x;
		}
	}
	int16_t use_idx[kMaxBases];
	FillIdentity(use_idx);
	return RunSubtract(const_cast<Context *>(context), kFitCubicSpline, num_data, data, 4,
			num_pieces, apply, boundary, use_idx, out);
}

sakura_Status sakura_SubtractSinusoidFloat(struct sakura_LSQFitContextFloat const *context,
		size_t num_data, float const data[], size_t num_nwave, size_t const nwave[],
		size_t num_coeff, double const coeff[], float out[]) noexcept {
	CHECK_ARGS(ValidContext(context));
	CHECK_ARGS(context->type == kFitSinusoid);
	CHECK_ARGS(num_data == context->num_data);
	CHECK_ARGS(data != nullptr && IsAligned(data));
	CHECK_ARGS(0 < num_nwave);
	CHECK_ARGS(nwave != nullptr);
	CHECK_ARGS(UniqueAscending(num_nwave, nwave));
	CHECK_ARGS(nwave[num_nwave - 1] <= context->param);
	size_t const needed = (nwave[0] == 0) ? (2 * num_nwave - 1) : (2 * num_nwave);
	CHECK_ARGS(needed <= num_coeff);
	CHECK_ARGS(num_coeff <= context->num_bases);
	CHECK_ARGS(coeff != nullptr && IsAligned(coeff));
	CHECK_ARGS(out != nullptr && IsAligned(out));
	std::vector<double> apply(coeff, coeff + num_coeff);
	int16_t use_idx[kMaxBases];
	FillSinusoidIndex(num_nwave, nwave, context->num_bases, use_idx);
	return RunSubtract(const_cast<Context *>(context), kFitSinusoid, num_data, data, num_coeff, 0,
			apply, nullptr, use_idx, out);
}

// ---------------------------------------------------------------- statistics

namespace {
struct StatScratch {
	DeviceBuffer data, mask, result, partial;
};
thread_local StatScratch *g_stat_scratch = nullptr;

sakura_Status RunStatsDevice(size_t num_spectra, size_t num_data, float const *d_data,
		size_t data_stride, uint8_t const *d_mask, size_t mask_stride, StatResult *d_result,
		DeviceBuffer *partial, cudaStream_t stream) {
	DeviceInfo info;
	sakura_Status st = CurrentDevice(&info);
	if (st != sakura_Status_kOK) {
		return st;
	}
	size_t chunks, chunk_len;
	size_t const pbytes = StatsPartialBytes(num_spectra, num_data, &chunks, &chunk_len);
	if (pbytes > 0) {
		CUDA_TRY(partial->Reserve(pbytes));
	}
	int launches = 0;
	cudaError_t e = LaunchStats(num_spectra, num_data, d_data, data_stride, d_mask, mask_stride,
			d_result, partial->ptr, info.num_sms, stream, &launches);
	g_launch_count += launches;
	g_last_kernel = "statistics";
	if (e != cudaSuccess) {
		SetError("LaunchStats", e);
		return sakura_Status_kUnknownError;
	}
	return sakura_Status_kOK;
}
} // namespace

sakura_Status sakura_ComputeStatisticsFloatBatchDevice(size_t num_spectra, size_t num_data,
		float const d_data[], size_t data_stride, bool const d_is_valid[], size_t mask_stride,
		sakura_StatisticsResultFloat d_result[], void *cuda_stream) noexcept {
	CHECK_ARGS(num_data <= static_cast<size_t>(INT32_MAX)); // statistics.cc:855
	CHECK_ARGS(d_data != nullptr && d_is_valid != nullptr && d_result != nullptr);
	if (g_stat_scratch == nullptr) {
		g_stat_scratch = new StatScratch();
	}
	return RunStatsDevice(num_spectra, num_data, d_data, data_stride,
			reinterpret_cast<uint8_t const *>(d_is_valid), mask_stride,
			reinterpret_cast<StatResult *>(d_result), &g_stat_scratch->partial,
			static_cast<cudaStream_t>(cuda_stream));
}

sakura_Status sakura_ComputeStatisticsFloatBatch(size_t num_spectra, size_t num_data,
		float const data[], size_t data_stride, bool const is_valid[], size_t mask_stride,
		sakura_StatisticsResultFloat result[]) noexcept {
	CHECK_ARGS(num_data <= static_cast<size_t>(INT32_MAX));
	CHECK_ARGS(data != nullptr);
	CHECK_ARGS(is_valid != nullptr);
	CHECK_ARGS(RowsAligned(data, data_stride * sizeof(float), num_spectra));
	CHECK_ARGS(RowsAligned(is_valid, mask_stride, num_spectra));
	CHECK_ARGS(result != nullptr);
	CHECK_ARGS(num_spectra <= 1 || (data_stride >= num_data && mask_stride >= num_data));
	if (num_spectra == 0) {
		return sakura_Status_kOK;
	}
	static_assert(sizeof(sakura_StatisticsResultFloat) == sizeof(StatResult), "layout");
	if (g_stat_scratch == nullptr) {
		g_stat_scratch = new StatScratch();
	}
	StatScratch &s = *g_stat_scratch;
	cudaStream_t const stream = 0;
	size_t const n = num_data;
	size_t const ds = RoundUp(n > 0 ? n : 1, 32);
	CUDA_TRY(s.data.Reserve(num_spectra * ds * sizeof(float)));
	CUDA_TRY(s.mask.Reserve(num_spectra * ds));
	CUDA_TRY(s.result.Reserve(num_spectra * sizeof(StatResult)));
	if (n > 0) {
		CUDA_TRY(cudaMemcpy2DAsync(s.data.ptr, ds * sizeof(float), data, data_stride * sizeof(float),
				n * sizeof(float), num_spectra, cudaMemcpyHostToDevice, stream));
		CUDA_TRY(cudaMemcpy2DAsync(s.mask.ptr, ds, is_valid, mask_stride, n, num_spectra,
				cudaMemcpyHostToDevice, stream));
	}
	sakura_Status st = RunStatsDevice(num_spectra, n, s.data.As<float>(), ds, s.mask.As<uint8_t>(),
			ds, s.result.As<StatResult>(), &s.partial, stream);
	if (st != sakura_Status_kOK) {
		return st;
	}
	CUDA_TRY(cudaMemcpyAsync(result, s.result.ptr, num_spectra * sizeof(StatResult),
			cudaMemcpyDeviceToHost, stream));
	CUDA_TRY(cudaStreamSynchronize(stream));
	return sakura_Status_kOK;
}

sakura_Status sakura_ComputeStatisticsFloat(size_t num_data, float const data[],
		bool const is_valid[], sakura_StatisticsResultFloat *result) noexcept {
	return sakura_ComputeStatisticsFloatBatch(1, num_data, data, num_data, is_valid, num_data,
			result);
}

sakura_Status sakura_ComputeAccurateStatisticsFloat(size_t num_data, float const data[],
		bool const is_valid[], sakura_StatisticsResultFloat *result) noexcept {
	return sakura_ComputeStatisticsFloatBatch(1, num_data, data, num_data, is_valid, num_data,
			result);
}


// ---------------------------------------------------------------- normal-equation building blocks

namespace {
struct BlockScratch {
	DeviceBuffer data, mask, basis, use_idx, G, b, x, tr, exclude, count;
};
thread_local BlockScratch *g_block_scratch = nullptr;

BlockScratch &BlockScratchGet() {
	if (g_block_scratch == nullptr) {
		g_block_scratch = new BlockScratch();
	}
	return *g_block_scratch;
}

sakura_Status UploadModel(BlockScratch &s, size_t num_data, float const *data, bool const *mask,
		size_t num_model_bases, double const *basis, size_t num_lsq_bases, size_t const *use_idx,
		cudaStream_t stream) {
	CUDA_TRY(s.data.Reserve(num_data * sizeof(float)));
	CUDA_TRY(s.mask.Reserve(num_data));
	CUDA_TRY(s.basis.Reserve(num_data * num_model_bases * sizeof(double)));
	CUDA_TRY(s.use_idx.Reserve(num_lsq_bases * sizeof(int)));
	CUDA_TRY(s.G.Reserve(num_lsq_bases * num_lsq_bases * sizeof(double)));
	CUDA_TRY(s.b.Reserve(num_lsq_bases * sizeof(double)));
	CUDA_TRY(s.count.Reserve(sizeof(int)));
	std::vector<int> idx(num_lsq_bases);
	for (size_t i = 0; i < num_lsq_bases; ++i) {
		idx[i] = static_cast<int>(use_idx[i]);
	}
	CUDA_TRY(cudaMemcpyAsync(s.data.ptr, data, num_data * sizeof(float), cudaMemcpyHostToDevice,
			stream));
	CUDA_TRY(cudaMemcpyAsync(s.mask.ptr, mask, num_data, cudaMemcpyHostToDevice, stream));
	CUDA_TRY(cudaMemcpyAsync(s.basis.ptr, basis, num_data * num_model_bases * sizeof(double),
			cudaMemcpyHostToDevice, stream));
	CUDA_TRY(cudaMemcpy(s.use_idx.ptr, idx.data(), num_lsq_bases * sizeof(int),
			cudaMemcpyHostToDevice));
	return sakura_Status_kOK;
}
} // namespace

sakura_Status sakura_GetLSQCoefficientsDouble(size_t const num_data, float const data[],
		bool const mask[], size_t const num_model_bases, double const basis_data[],
		size_t const num_lsq_bases, size_t const use_bases_idx[], double lsq_matrix[],
		double lsq_vector[]) noexcept {
	CHECK_ARGS(num_data != 0 && num_data <= static_cast<size_t>(INT32_MAX));
	CHECK_ARGS(data != nullptr && IsAligned(data));
	CHECK_ARGS(mask != nullptr && IsAligned(mask));
	CHECK_ARGS(num_model_bases != 0);
	CHECK_ARGS(num_model_bases <= num_data);
	CHECK_ARGS(basis_data != nullptr && IsAligned(basis_data));
	CHECK_ARGS(num_lsq_bases != 0);
	CHECK_ARGS(num_lsq_bases <= num_model_bases);
	CHECK_ARGS(use_bases_idx != nullptr && IsAligned(use_bases_idx));
	CHECK_ARGS(lsq_matrix != nullptr && IsAligned(lsq_matrix));
	CHECK_ARGS(lsq_vector != nullptr && IsAligned(lsq_vector));
	for (size_t i = 0; i < num_lsq_bases; ++i) {
		CHECK_ARGS(use_bases_idx[i] < num_model_bases);
	}
	BlockScratch &s = BlockScratchGet();
	cudaStream_t const stream = 0;
	sakura_Status st = UploadModel(s, num_data, data, mask, num_model_bases, basis_data,
			num_lsq_bases, use_bases_idx, stream);
	if (st != sakura_Status_kOK) {
		return st;
	}
	cudaError_t e = LaunchNormalEquations(static_cast<int>(num_data), static_cast<int>(num_lsq_bases),
			s.basis.As<double>(), static_cast<int>(num_model_bases), s.use_idx.As<int>(),
			s.mask.As<uint8_t>(), s.data.As<float>(), s.G.As<double>(), s.b.As<double>(),
			s.count.As<int>(), stream);
	g_launch_count += 1;
	g_last_kernel = "normal_equations";
	if (e != cudaSuccess) {
		SetError("LaunchNormalEquations", e);
		return sakura_Status_kUnknownError;
	}
	int count = 0;
	CUDA_TRY(cudaMemcpyAsync(lsq_matrix, s.G.ptr, num_lsq_bases * num_lsq_bases * sizeof(double),
			cudaMemcpyDeviceToHost, stream));
	CUDA_TRY(cudaMemcpyAsync(lsq_vector, s.b.ptr, num_lsq_bases * sizeof(double),
			cudaMemcpyDeviceToHost, stream));
	CUDA_TRY(cudaMemcpyAsync(&count, s.count.ptr, sizeof(int), cudaMemcpyDeviceToHost, stream));
	CUDA_TRY(cudaStreamSynchronize(stream));
	// numeric_operation.cc:221-224 (templates, <= 100 bases) / :266-269 (generic)
	size_t const need = (num_lsq_bases <= 100) ? num_lsq_bases : num_model_bases;
	if (static_cast<size_t>(count) < need) {
		return sakura_Status_kNG;
	}
	return sakura_Status_kOK;
}

sakura_Status sakura_UpdateLSQCoefficientsDouble(size_t const num_data, float const data[],
		bool const mask[], size_t const num_exclude_indices, size_t const exclude_indices[],
		size_t const num_model_bases, double const basis_data[], size_t const num_lsq_bases,
		size_t const use_bases_idx[], double lsq_matrix[], double lsq_vector[]) noexcept {
	CHECK_ARGS(num_data != 0 && num_data <= static_cast<size_t>(INT32_MAX));
	CHECK_ARGS(data != nullptr && IsAligned(data));
	CHECK_ARGS(mask != nullptr && IsAligned(mask));
	CHECK_ARGS(exclude_indices != nullptr && IsAligned(exclude_indices));
	CHECK_ARGS(num_exclude_indices <= num_data);
	for (size_t i = 0; i < num_exclude_indices; ++i) {
		CHECK_ARGS(exclude_indices[i] < num_data);
	}
	CHECK_ARGS(num_model_bases != 0);
	CHECK_ARGS(num_model_bases <= num_data);
	CHECK_ARGS(basis_data != nullptr && IsAligned(basis_data));
	CHECK_ARGS(num_lsq_bases != 0);
	CHECK_ARGS(num_lsq_bases <= num_model_bases);
	CHECK_ARGS(use_bases_idx != nullptr && IsAligned(use_bases_idx));
	CHECK_ARGS(lsq_matrix != nullptr && IsAligned(lsq_matrix));
	CHECK_ARGS(lsq_vector != nullptr && IsAligned(lsq_vector));
	for (size_t i = 0; i < num_lsq_bases; ++i) {
		CHECK_ARGS(use_bases_idx[i] < num_model_bases);
	}
	BlockScratch &s = BlockScratchGet();
	cudaStream_t const stream = 0;
	sakura_Status st = UploadModel(s, num_data, data, mask, num_model_bases, basis_data,
			num_lsq_bases, use_bases_idx, stream);
	if (st != sakura_Status_kOK) {
		return st;
	}
	CUDA_TRY(s.exclude.Reserve((num_exclude_indices + 1) * sizeof(unsigned long long)));
	CUDA_TRY(cudaMemcpyAsync(s.exclude.ptr, exclude_indices, num_exclude_indices * sizeof(size_t),
			cudaMemcpyHostToDevice, stream));
	CUDA_TRY(cudaMemcpyAsync(s.G.ptr, lsq_matrix, num_lsq_bases * num_lsq_bases * sizeof(double),
			cudaMemcpyHostToDevice, stream));
	CUDA_TRY(cudaMemcpyAsync(s.b.ptr, lsq_vector, num_lsq_bases * sizeof(double),
			cudaMemcpyHostToDevice, stream));
	cudaError_t e = LaunchUpdateNormalEquations(static_cast<int>(num_lsq_bases), s.basis.As<double>(),
			static_cast<int>(num_model_bases), s.use_idx.As<int>(), s.mask.As<uint8_t>(),
			s.data.As<float>(), static_cast<int>(num_exclude_indices),
			s.exclude.As<unsigned long long>(), s.G.As<double>(), s.b.As<double>(), stream);
	g_launch_count += 1;
	g_last_kernel = "update_normal_equations";
	if (e != cudaSuccess) {
		SetError("LaunchUpdateNormalEquations", e);
		return sakura_Status_kUnknownError;
	}
	CUDA_TRY(cudaMemcpyAsync(lsq_matrix, s.G.ptr, num_lsq_bases * num_lsq_bases * sizeof(double),
			cudaMemcpyDeviceToHost, stream));
	CUDA_TRY(cudaMemcpyAsync(lsq_vector, s.b.ptr, num_lsq_bases * sizeof(double),
			cudaMemcpyDeviceToHost, stream));
	CUDA_TRY(cudaStreamSynchronize(stream));
	return sakura_Status_kOK;
}

sakura_Status sakura_SolveSimultaneousEquationsByLUDouble(size_t num_equations,
		double const in_matrix[], double const in_vector[], double out[]) noexcept {
	CHECK_ARGS(in_matrix != nullptr && IsAligned(in_matrix));
	CHECK_ARGS(in_vector != nullptr && IsAligned(in_vector));
	CHECK_ARGS(in_vector != out);
	CHECK_ARGS(out != nullptr && IsAligned(out));
	if (num_equations == 0) {
		return sakura_Status_kOK;
	}
	CHECK_ARGS(num_equations <= 32768);
	BlockScratch &s = BlockScratchGet();
	cudaStream_t const stream = 0;
	size_t const K = num_equations;
	CUDA_TRY(s.G.Reserve(K * K * sizeof(double)));
	CUDA_TRY(s.b.Reserve(K * sizeof(double)));
	CUDA_TRY(s.x.Reserve(K * sizeof(double)));
	CUDA_TRY(s.tr.Reserve(2 * K * sizeof(int)));
	CUDA_TRY(cudaMemcpyAsync(s.G.ptr, in_matrix, K * K * sizeof(double), cudaMemcpyHostToDevice,
			stream));
	CUDA_TRY(cudaMemcpyAsync(s.b.ptr, in_vector, K * sizeof(double), cudaMemcpyHostToDevice, stream));
	cudaError_t e = LaunchSolve(static_cast<int>(K), s.G.As<double>(), s.b.As<double>(),
			s.x.As<double>(), s.tr.As<int>(), stream);
	g_launch_count += 1;
	g_last_kernel = "solve_full_piv_lu";
	if (e != cudaSuccess) {
		SetError("LaunchSolve", e);
		return sakura_Status_kUnknownError;
	}
	CUDA_TRY(cudaMemcpyAsync(out, s.x.ptr, K * sizeof(double), cudaMemcpyDeviceToHost, stream));
	CUDA_TRY(cudaStreamSynchronize(stream));
	return sakura_Status_kOK;
}
