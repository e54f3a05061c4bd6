// C ABI of libsakura_b200.so, part 1: global state, context objects and the fit dispatch.
// Declarations and the reference interface each entry point replaces are in include/sakura_b200.h.
//
// There is NO CPU fallback: every compute call runs a CUDA kernel of this library
// and fails loudly (status + message on stderr) if no CUDA device is usable.
#include "capi_internal.cuh"

#include <algorithm>

using namespace sakura_b200;
using namespace sakura_b200::capi;

namespace sakura_b200 {
namespace capi {

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
		c->d_pack.Release();
		if (c->h_pack != nullptr) {
			cudaFreeHost(c->h_pack);
			c->h_pack = nullptr;
			c->h_pack_cap = 0;
		}
		if (c->pack_stream != nullptr) {
			cudaStreamDestroy(c->pack_stream);
			c->pack_stream = nullptr;
		}
		if (prev != c->device) {
			cudaSetDevice(prev);
		}
	}
	c->device = -1;
}

// Column sums of the extended table over blocks of kSinSumStep channels and over all channels,
// appended to the table: [steps + 1][ext_ld], the last row the total. The fit kernel gets the sums over
// a spectrum's VALID channels as "all minus flagged" (a few per cent of the channels) instead of
// contracting the whole mask with the table. Sums of the table's own doubles, in extended precision.
void AppendExtSums(std::vector<double> *ext, int cols, size_t num_data) {
	size_t const ext_ld = static_cast<size_t>(SinusoidExtLd(cols));
	size_t const chunk_doubles = num_data * kSinExtChunkLd;
	size_t const steps = (num_data + kSinSumStep - 1) / kSinSumStep;
	size_t const base = ext->size();
	ext->resize(base + (steps + 1) * ext_ld, 0.0);
	std::vector<long double> total(ext_ld, 0.0L);
	for (size_t s = 0; s < steps; ++s) {
		size_t const i1 = std::min(num_data, (s + 1) * kSinSumStep);
		for (size_t col = 0; col < ext_ld; ++col) {
			long double acc = 0.0L;
			for (size_t i = s * kSinSumStep; i < i1; ++i) {
				acc += static_cast<long double>((*ext)[(col / 32) * chunk_doubles + i * kSinExtChunkLd + (col % 32)]);
			}
			(*ext)[base + s * ext_ld + col] = static_cast<double>(acc);
			total[col] += acc;
		}
	}
	for (size_t col = 0; col < ext_ld; ++col) {
		(*ext)[base + steps * ext_ld + col] = static_cast<double>(total[col]);
	}
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
	c->h_pack = nullptr;
	c->h_pack_cap = 0;
	c->pack_stream = nullptr;
	c->replicas = nullptr;
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
			AppendExtSums(c->h_ext, cols, num_data);
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
			AppendExtSums(c->h_ext, cols, num_data);
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
		// Exception (measured, profiles/README.md): where the warp-per-spectrum kernel fits (2048 / 4096
		// channels from two fits on) the pre-pass at HBM speed plus the plain fit beats calibrating inside
		// that latency-bound kernel (10.9 against 12.4 ms at 524 288 x 2048), so the pre-pass is used there.
		bool const fused = (f.type == kFitPolynomial || (f.type == kFitChebyshev && f.K <= 6))
				&& !PolyPrefersCalibrationPrePass(f.num_data, f.nfit)
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
		sp.next = c->ext_cols;
		bool columns_ok = true;
		sp.cos_only = cheb_dmma ? 1 : 0;
		// Slots of the fit: the bases that are EVEN about the middle of the row first (constant, cosines;
		// Chebyshev T_m with even m), padded to a multiple of 8, then the ODD ones (sines; odd m), padded
		// likewise -- the kernel contracts over half a row with the folded data y_i +- y_{n-1-i}
		// (lsqfit_sinusoid.cu). Padding slots have kind -1 and no column of the context table.
		int16_t slot_col[kSinMaxKp];
		int num_slots = 0;
		for (int pass = 0; pass < 2; ++pass) {
			for (size_t k = 0; k < f.K; ++k) {
				int const col = f.use_idx[k];
				columns_ok = columns_ok && (col >= 0 && static_cast<size_t>(col) < c->num_bases);
				bool const odd = (col & 1) != 0; // sinusoid: 1, sin w, cos w, ...; Chebyshev: T_col
				if (odd != (pass == 1) || num_slots >= kSinMaxKp) {
					continue;
				}
				if (cheb_dmma) { // basis T_col: a "cosine" of wave number col
					sp.kind[num_slots] = static_cast<int8_t>(col == 0 ? 0 : 2);
					sp.wave[num_slots] = static_cast<int16_t>(col);
				} else {
					sp.kind[num_slots] = static_cast<int8_t>(col == 0 ? 0 : ((col & 1) ? 1 : 2));
					sp.wave[num_slots] = static_cast<int16_t>((col + 1) / 2);
				}
				sp.out_pos[num_slots] = static_cast<int8_t>(k);
				slot_col[num_slots] = static_cast<int16_t>(col);
				++num_slots;
			}
			while (num_slots % 8 != 0 && num_slots < kSinMaxKp) {
				sp.kind[num_slots] = -1;
				sp.wave[num_slots] = 0;
				sp.out_pos[num_slots] = -1;
				slot_col[num_slots] = -1;
				++num_slots;
			}
			if (pass == 0) {
				sp.kc_p = num_slots;
			}
		}
		sp.Kp = num_slots;
		{ // every basis has a slot (the padded groups fit into kSinMaxKp slots)?
			int real = 0;
			for (int s2 = 0; s2 < num_slots; ++s2) {
				real += sp.out_pos[s2] >= 0 ? 1 : 0;
			}
			columns_ok = columns_ok && real == sp.K && SinusoidLayoutSupported(sp.kc_p, sp.Kp);
		}
		if (columns_ok) {
			size_t tiles = (f.num_spectra + kSinRowsPerTile - 1) / kSinRowsPerTile;
			size_t sgrid = static_cast<size_t>(info.num_sms) * kSinCtasPerSm;
			if (sgrid > tiles) {
				sgrid = tiles;
			}
			CUDA_TRY(c->d_table.Reserve(SinusoidTableDoubles(sp.n, sp.Kp) * sizeof(double)
					+ SinusoidGramEntries(sp.Kp) * sizeof(unsigned)));
			CUDA_TRY(c->d_useidx.Reserve(kMaxBases * sizeof(int16_t)));
			CUDA_TRY(c->d_fallback.Reserve((f.num_spectra + 1) * sizeof(int32_t)));
			CUDA_TRY(c->ws_mask.Reserve(sgrid * kSinRowsPerTile * n));
			CUDA_TRY(c->ws_sin_residual.Reserve(sgrid * kSinRowsPerTile * n * sizeof(float)));
			if (f.best_fit != nullptr) {
				CUDA_TRY(c->ws_sin_best_fit.Reserve(sgrid * kSinRowsPerTile * n * sizeof(float)));
			}
			// (pageable source: the copy is staged by the runtime before the call returns)
			CUDA_TRY(cudaMemcpyAsync(c->d_useidx.ptr, slot_col, kSinMaxKp * sizeof(int16_t),
					cudaMemcpyHostToDevice, stream));
			cudaError_t e = LaunchGatherTable(c->d_basis.As<double>(), static_cast<int>(c->num_bases),
					sp.n, sp.K, sp.Kp, c->d_useidx.As<int16_t>(), c->d_table.As<double>(), sp, stream);
			g_launch_count += 2;
			if (e != cudaSuccess) {
				SetError("LaunchGatherTable", e);
				return sakura_Status_kUnknownError;
			}
			int32_t *fb = c->d_fallback.As<int32_t>();
			CUDA_TRY(cudaMemsetAsync(fb, 0, sizeof(int32_t), stream));
			sp.ld_p = sp.Kp + 1;
			sp.ld_a = SinusoidLdA(sp.Kp);
			sp.gram_table = reinterpret_cast<unsigned const *>(c->d_table.As<double>()
					+ SinusoidTableDoubles(sp.n, sp.Kp));
			sp.table_p = c->d_table.As<double>();
			sp.table_a = sp.table_p + n * sp.ld_p;
			sp.ext_p = c->d_ext.As<double>();
			sp.ext_ld = SinusoidExtLd(c->ext_cols);
			sp.ext_sums = sp.ext_p + static_cast<size_t>(sp.ext_ld / 32) * n * kSinExtChunkLd;
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

} // namespace capi
} // namespace sakura_b200

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
	if (context->replicas != nullptr) {
		for (auto &r : *context->replicas) {
			sakura_DestroyLSQFitContextFloat(r.second);
		}
		delete context->replicas;
		context->replicas = nullptr;
	}
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
