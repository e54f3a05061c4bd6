// C ABI of libsakura_b200.so, part 2: the least-squares fit entry points (host-pointer batches with the
// copy/compute pipeline, device-array batches, the reference's single-row signatures, Subtract*).
#include "capi_internal.cuh"
#include "host_bits.h"
#include "mask_bits.cuh"

#include <thread>
#include <time.h>

using namespace sakura_b200;
using namespace sakura_b200::capi;

namespace {

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

// Row flags (set by the kernels): bit0 = results published, bit1 = final_mask written,
// bit2 = boundary written (cubic splines). A row copies back exactly what the reference would have
// written.
inline int AllPublished(int type) {
	return type == kFitCubicSpline ? 7 : 3;
}

/** Row-strided copy; a single 1-D copy when both sides are tightly packed (faster on the
 *  copy engines than the pitched form). */
cudaError_t CopyRows(void *dst, size_t dpitch, void const *src, size_t spitch, size_t width,
		size_t rows, cudaMemcpyKind kind, cudaStream_t stream) {
	if (dpitch == width && spitch == width) {
		return cudaMemcpyAsync(dst, src, width * rows, kind, stream);
	}
	return cudaMemcpy2DAsync(dst, dpitch, src, spitch, width, rows, kind, stream);
}

// Tuning knobs of the host-buffer paths: defaults from measurements on B200, overridden by the
// environment at load time or by sakura_b200_SetHostPipelineTuning at any time.
size_t EnvSize(char const *name, long lo, long hi, size_t fallback) {
	char const *e = getenv(name);
	if (e == nullptr) {
		return fallback;
	}
	long const v = atol(e);
	return (v >= lo && v <= hi) ? static_cast<size_t>(v) : fallback;
}

// chunk size of the pipeline in MiB of data
std::atomic<size_t> g_chunk_mib(EnvSize("SAKURA_B200_CHUNK_MIB", 1, 4096, 96));
// batches whose staged size is below this many KiB go through the single-arena path
std::atomic<size_t> g_packed_kib(EnvSize("SAKURA_B200_PACKED_KIB", 0, 1 << 20, 2048));

size_t ChunkMiB() {
	return g_chunk_mib.load();
}

size_t PackedLimitBytes() {
	return g_packed_kib.load() << 10;
}

// ---------------------------------------------------------------------------------------------
// Pipelined variant: the batch is cut into chunks that cycle through kPipelineSlots (stream, device
// buffers) slots, so that the host->device copy of chunk i+1, the kernel of chunk i and the
// device->host copy of chunk i-1 overlap on the copy engines and the SMs. With pinned host buffers
// every copy is truly asynchronous. The polynomial fast path brings a hand-over list per slot, so its
// kernels may overlap freely; every other kernel uses scratch of the context and is ordered after
// the previous chunk's kernel by an event (the uploads still run ahead).
// ---------------------------------------------------------------------------------------------
// Optional timeline of a pipelined call (SAKURA_B200_PIPELINE_TRACE=1): device timestamps of every
// chunk's upload, kernel and download and the host times of their submission, printed to stderr.
struct ChunkTrace {
	cudaEvent_t up0, up1, k1, d0, d1;
	double host_submit_ms, host_finish_ms;
};

double HostMs() {
	timespec ts;
	clock_gettime(CLOCK_MONOTONIC, &ts);
	return ts.tv_sec * 1e3 + ts.tv_nsec * 1e-6;
}

// Masks over PCIe as bits (host_bits.h): on unless SAKURA_B200_PACK_MASKS=0 (read per call: A/B timing).
bool PackMasksEnabled() {
	char const *e = getenv("SAKURA_B200_PACK_MASKS");
	return e == nullptr || e[0] != '0';
}

sakura_Status PipelineBody(Context *c, FitCall const &f, HostFitArgs const &h, size_t chunk_rows,
		bool fast) {
	static bool const trace_on = getenv("SAKURA_B200_PIPELINE_TRACE") != nullptr;
	std::vector<ChunkTrace> trace;
	double const host_t0 = HostMs();
	size_t const n = f.num_data;
	size_t const R = f.num_spectra;
	size_t const dstride = RoundUp(n, 32);
	size_t const hds = f.data_stride, hms = f.mask_stride;
	size_t const nb = f.npiece + 1;
	bool const spline = (f.type == kFitCubicSpline);
	int const published = AllPublished(f.type);
	bool const pack = (n % 32 == 0) && PackMasksEnabled();
	size_t const words = n / 32;
	std::vector<size_t> start; // chunk boundaries (a ramp of shorter chunks at both ends was measured: no gain)
	for (size_t r = 0; r < R; r += chunk_rows) {
		start.push_back(r);
	}
	start.push_back(R);
	size_t const num_chunks = start.size() - 1;
	sakura_Status st = sakura_Status_kOK;
	if (trace_on) {
		trace.resize(num_chunks);
		for (auto &t : trace) {
			cudaEventCreate(&t.up0);
			cudaEventCreate(&t.up1);
			cudaEventCreate(&t.k1);
			cudaEventCreate(&t.d0);
			cudaEventCreate(&t.d1);
			t.host_submit_ms = t.host_finish_ms = 0.0;
		}
	}
	for (int s = 0; s < kPipelineSlots; ++s) {
		PipelineSlot &sl = c->slots[s];
		if (sl.stream == nullptr) {
			CUDA_TRY(cudaStreamCreateWithFlags(&sl.stream, cudaStreamNonBlocking));
			CUDA_TRY(cudaStreamCreateWithFlags(&sl.down, cudaStreamNonBlocking));
			CUDA_TRY(cudaEventCreateWithFlags(&sl.kernel_done, cudaEventDisableTiming));
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
		if (spline) {
			CUDA_TRY(sl.boundary.Reserve(chunk_rows * nb * sizeof(unsigned long long)));
		}
		CUDA_TRY(sl.handover.Reserve((chunk_rows + 1) * sizeof(int32_t)));
		if (sl.h_flags_cap < chunk_rows) {
			if (sl.h_flags != nullptr) {
				cudaFreeHost(sl.h_flags);
				sl.h_flags = nullptr;
				sl.h_flags_cap = 0;
			}
			CUDA_TRY(cudaHostAlloc(reinterpret_cast<void **>(&sl.h_flags), chunk_rows * sizeof(int32_t),
					cudaHostAllocDefault));
			sl.h_flags_cap = chunk_rows;
		}
		if (pack) {
			CUDA_TRY(sl.mask_bits.Reserve(chunk_rows * words * sizeof(uint32_t)));
			CUDA_TRY(sl.fmask_bits.Reserve(chunk_rows * words * sizeof(uint32_t)));
			if (sl.h_bits_cap < chunk_rows * words) {
				if (sl.h_mask_bits != nullptr) {
					cudaFreeHost(sl.h_mask_bits);
					sl.h_mask_bits = nullptr;
				}
				if (sl.h_fmask_bits != nullptr) {
					cudaFreeHost(sl.h_fmask_bits);
					sl.h_fmask_bits = nullptr;
				}
				sl.h_bits_cap = 0;
				CUDA_TRY(cudaHostAlloc(reinterpret_cast<void **>(&sl.h_mask_bits),
						chunk_rows * words * sizeof(uint32_t), cudaHostAllocDefault));
				CUDA_TRY(cudaHostAlloc(reinterpret_cast<void **>(&sl.h_fmask_bits),
						chunk_rows * words * sizeof(uint32_t), cudaHostAllocDefault));
				sl.h_bits_cap = chunk_rows * words;
			}
		}
		sl.unpack_rows = 0;
		sl.rows = 0;
	}
	// final-mask bits of an earlier chunk of this slot, downloaded meanwhile: expand them into the caller's
	// array (worker threads) before the staging buffer is used again
	auto complete_unpack = [&](PipelineSlot &sl) -> sakura_Status {
		if (sl.unpack_rows == 0) {
			return sakura_Status_kOK;
		}
		CUDA_TRY(cudaEventSynchronize(sl.down_done));
		UnpackMaskRows(sl.h_fmask_bits, n, sl.unpack_rows,
				reinterpret_cast<uint8_t *>(h.final_mask) + sl.unpack_row0 * hms, hms,
				sl.unpack_all ? nullptr : sl.unpack_flags.data(), 2);
		sl.unpack_rows = 0;
		return sakura_Status_kOK;
	};
	auto finish = [&](PipelineSlot &sl) -> sakura_Status {
		// the kernel of this slot's chunk is done and its per-row flags are on the host
		CUDA_TRY(cudaEventSynchronize(sl.flags_ready));
		size_t const r0 = sl.row0, rows = sl.rows;
		ChunkTrace *tr = trace_on ? &trace[sl.chunk] : nullptr;
		if (tr != nullptr) {
			tr->host_finish_ms = HostMs() - host_t0;
			cudaEventRecord(tr->d0, sl.down);
		}
		float *d_rms = sl.small.As<float>();
		bool all_published = true;
		for (size_t r = 0; r < rows; ++r) {
			if ((sl.h_flags[r] & 7) != published) {
				all_published = false;
				break;
			}
		}
		if (pack) {
			sakura_Status const us = complete_unpack(sl);
			if (us != sakura_Status_kOK) {
				return us;
			}
			CUDA_TRY(cudaMemcpyAsync(sl.h_fmask_bits, sl.fmask_bits.ptr, rows * words * sizeof(uint32_t),
					cudaMemcpyDeviceToHost, sl.down));
			sl.unpack_rows = rows;
			sl.unpack_row0 = r0;
			sl.unpack_all = all_published;
			if (!all_published) {
				sl.unpack_flags.assign(sl.h_flags, sl.h_flags + rows);
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
			if (!pack) {
				CUDA_TRY(CopyRows(h.final_mask + r0 * hms, hms, sl.final_mask.ptr, dstride, n, rows,
						cudaMemcpyDeviceToHost, sl.down));
			}
			CUDA_TRY(cudaMemcpyAsync(h.rms + r0, d_rms, rows * sizeof(float), cudaMemcpyDeviceToHost,
					sl.down));
			if (spline) {
				CUDA_TRY(cudaMemcpyAsync(h.boundary + r0 * nb, sl.boundary.ptr,
						rows * nb * sizeof(unsigned long long), cudaMemcpyDeviceToHost, sl.down));
			}
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
				if ((fl & 2) && !pack) {
					CUDA_TRY(cudaMemcpyAsync(h.final_mask + (r0 + r) * hms,
							sl.final_mask.As<uint8_t>() + r * dstride, n, cudaMemcpyDeviceToHost,
							sl.down));
				}
				if (spline && (fl & 4)) {
					CUDA_TRY(cudaMemcpyAsync(h.boundary + (r0 + r) * nb,
							sl.boundary.As<unsigned long long>() + r * nb,
							nb * sizeof(unsigned long long), cudaMemcpyDeviceToHost, sl.down));
				}
			}
		}
		if (tr != nullptr) {
			cudaEventRecord(tr->d1, sl.down);
		}
		CUDA_TRY(cudaEventRecord(sl.down_done, sl.down));
		sl.rows = 0;
		return sakura_Status_kOK;
	};
	PipelineSlot *previous = nullptr;
	for (size_t ci = 0; ci < num_chunks; ++ci) {
		PipelineSlot &sl = c->slots[ci % kPipelineSlots];
		if (sl.rows != 0) {
			st = finish(sl);
			if (st != sakura_Status_kOK) {
				return st;
			}
		}
		size_t const r0 = start[ci];
		size_t const rows = start[ci + 1] - r0;
		sl.row0 = r0;
		sl.rows = rows;
		sl.chunk = ci;
		float *d_rms = sl.small.As<float>();
		int32_t *d_status = reinterpret_cast<int32_t *>(d_rms + chunk_rows);
		int32_t *d_lsq = d_status + chunk_rows;
		int32_t *d_flags = d_lsq + chunk_rows;
		if (trace_on) {
			trace[ci].host_submit_ms = HostMs() - host_t0;
			cudaEventRecord(trace[ci].up0, sl.stream);
		}
		CUDA_TRY(CopyRows(sl.data.ptr, dstride * sizeof(float), h.data + r0 * hds,
				hds * sizeof(float), n * sizeof(float), rows, cudaMemcpyHostToDevice, sl.stream));
		if (pack) {
			// (the slot's previous upload of the staging words is long done: its kernel has finished)
			PackMaskRows(reinterpret_cast<uint8_t const *>(h.mask) + r0 * hms, hms, n, rows, sl.h_mask_bits);
			CUDA_TRY(cudaMemcpyAsync(sl.mask_bits.ptr, sl.h_mask_bits, rows * words * sizeof(uint32_t),
					cudaMemcpyHostToDevice, sl.stream));
			CUDA_TRY(LaunchMaskBitsToBytes(sl.mask_bits.As<uint32_t>(), words, rows, sl.mask.As<uint8_t>(),
					dstride, sl.stream));
			g_launch_count += 1;
		} else {
			CUDA_TRY(CopyRows(sl.mask.ptr, dstride, h.mask + r0 * hms, hms, n, rows,
					cudaMemcpyHostToDevice, sl.stream));
		}
		if (trace_on) {
			cudaEventRecord(trace[ci].up1, sl.stream);
		}
		// the kernel overwrites this slot's result buffers: wait for their previous download
		CUDA_TRY(cudaStreamWaitEvent(sl.stream, sl.down_done, 0));
		if (!fast && previous != nullptr) {
			// scratch of the context is shared by the chunks: kernels one after the other
			CUDA_TRY(cudaStreamWaitEvent(sl.stream, previous->kernel_done, 0));
		}
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
		fc.boundary = spline ? sl.boundary.As<unsigned long long>() : nullptr;
		fc.handover = fast ? sl.handover.As<int32_t>() : nullptr;
		st = LaunchFit(c, fc, sl.stream);
		if (st != sakura_Status_kOK) {
			return st;
		}
		if (pack) {
			CUDA_TRY(LaunchMaskBytesToBits(sl.final_mask.As<uint8_t>(), dstride, words, rows,
					sl.fmask_bits.As<uint32_t>(), sl.stream));
			g_launch_count += 1;
		}
		if (trace_on) {
			cudaEventRecord(trace[ci].k1, sl.stream);
		}
		CUDA_TRY(cudaEventRecord(sl.kernel_done, sl.stream));
		CUDA_TRY(cudaStreamWaitEvent(sl.down, sl.kernel_done, 0));
		CUDA_TRY(cudaMemcpyAsync(sl.h_flags, d_flags, rows * sizeof(int32_t), cudaMemcpyDeviceToHost,
				sl.down));
		CUDA_TRY(cudaMemcpyAsync(h.status + r0, d_status, rows * sizeof(int32_t),
				cudaMemcpyDeviceToHost, sl.down));
		CUDA_TRY(cudaMemcpyAsync(h.lsq + r0, d_lsq, rows * sizeof(int32_t), cudaMemcpyDeviceToHost,
				sl.down));
		CUDA_TRY(cudaEventRecord(sl.flags_ready, sl.down));
		previous = &sl;
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
	if (pack) {
		for (int s = 0; s < kPipelineSlots; ++s) {
			st = complete_unpack(c->slots[s]);
			if (st != sakura_Status_kOK) {
				return st;
			}
		}
	}
	if (trace_on) {
		fprintf(stderr, "sakura_b200 pipeline trace: %zu chunks of %zu rows, host total %.3f ms\n"
				"chunk  submit  up0     up1     k1      finish  d0      d1   (ms; device times relative to up0 of chunk 0)\n",
				num_chunks, chunk_rows, HostMs() - host_t0);
		for (size_t ci = 0; ci < num_chunks; ++ci) {
			float t[5] = { 0, 0, 0, 0, 0 };
			cudaEventElapsedTime(&t[0], trace[0].up0, trace[ci].up0);
			cudaEventElapsedTime(&t[1], trace[0].up0, trace[ci].up1);
			cudaEventElapsedTime(&t[2], trace[0].up0, trace[ci].k1);
			cudaEventElapsedTime(&t[3], trace[0].up0, trace[ci].d0);
			cudaEventElapsedTime(&t[4], trace[0].up0, trace[ci].d1);
			fprintf(stderr, "%5zu %7.3f %7.3f %7.3f %7.3f %7.3f %7.3f %7.3f\n", ci, trace[ci].host_submit_ms, t[0],
					t[1], t[2], trace[ci].host_finish_ms, t[3], t[4]);
		}
		for (auto &t : trace) {
			cudaEventDestroy(t.up0);
			cudaEventDestroy(t.up1);
			cudaEventDestroy(t.k1);
			cudaEventDestroy(t.d0);
			cudaEventDestroy(t.d1);
		}
	}
	return sakura_Status_kOK;
}

sakura_Status RunHostBatchPipelined(Context *c, FitCall const &f, HostFitArgs const &h,
		size_t chunk_rows, bool fast, DeviceInfo const &info) {
	if (c->device >= 0 && c->device != info.device) {
		ReleaseDeviceState(c);
	}
	c->device = info.device;
	if (c->slots == nullptr) {
		c->slots = new PipelineSlot[kPipelineSlots];
	}
	sakura_Status const st = PipelineBody(c, f, h, chunk_rows, fast);
	if (st != sakura_Status_kOK) {
		// copies into the caller's buffers may still be in flight on the slot streams: drain them
		// before the caller sees the error, and leave no slot marked busy for the next call
		for (int s = 0; s < kPipelineSlots; ++s) {
			PipelineSlot &sl = c->slots[s];
			if (sl.stream != nullptr) {
				(void) cudaStreamSynchronize(sl.stream);
			}
			if (sl.down != nullptr) {
				(void) cudaStreamSynchronize(sl.down);
			}
			sl.rows = 0;
			sl.unpack_rows = 0;
		}
		(void) cudaGetLastError();
	}
	return st;
}

// ---------------------------------------------------------------------------------------------
// Small batches (the reference's single-row signatures are batches of one): everything a call moves
// lives in ONE device arena mirrored by ONE pinned host arena,
//     [ data | mask | flags ]  ->  one host->device copy (flags = 0)
//     [ flags | status | lsqfit_status | rms | coeff | boundary | final_mask | residual | best_fit ]
//                              <-  one device->host copy,
// on a stream of the context's own: two copies, the kernel(s) and one synchronisation per call
// instead of two synchronisations and ten copies out of pageable memory.
// ---------------------------------------------------------------------------------------------
sakura_Status RunHostBatchPacked(Context *c, FitCall f, HostFitArgs const &h, DeviceInfo const &info) {
	size_t const n = f.num_data;
	size_t const R = f.num_spectra;
	size_t const dstride = RoundUp(n, 32);
	size_t const hds = f.data_stride, hms = f.mask_stride;
	size_t const nb = f.npiece + 1;
	bool const spline = (f.type == kFitCubicSpline);
	if (c->device >= 0 && c->device != info.device) {
		ReleaseDeviceState(c);
	}
	c->device = info.device;
	if (c->pack_stream == nullptr) {
		CUDA_TRY(cudaStreamCreateWithFlags(&c->pack_stream, cudaStreamNonBlocking));
	}
	cudaStream_t const stream = c->pack_stream;
	size_t off = 0;
	auto take = [&off](size_t bytes) {
		size_t const at = off;
		off = RoundUp(off + bytes, 256);
		return at;
	};
	size_t const o_data = take(R * dstride * sizeof(float));
	size_t const o_mask = take(R * dstride);
	size_t const o_flags = take(R * sizeof(int32_t));
	size_t const o_status = take(R * sizeof(int32_t));
	size_t const o_lsq = take(R * sizeof(int32_t));
	size_t const o_rms = take(R * sizeof(float));
	size_t const o_coeff = take(h.coeff ? R * f.coeff_stride * sizeof(double) : 0);
	size_t const o_boundary = take(spline ? R * nb * sizeof(unsigned long long) : 0);
	size_t const o_final = take(R * dstride);
	size_t const o_resid = take(h.residual ? R * dstride * sizeof(float) : 0);
	size_t const o_best = take(h.best_fit ? R * dstride * sizeof(float) : 0);
	size_t const total = off;
	if (c->h_pack_cap < total) {
		if (c->h_pack != nullptr) {
			cudaFreeHost(c->h_pack);
			c->h_pack = nullptr;
			c->h_pack_cap = 0;
		}
		size_t const want = total + total / 4;
		CUDA_TRY(cudaHostAlloc(&c->h_pack, want, cudaHostAllocDefault));
		c->h_pack_cap = want;
	}
	CUDA_TRY(c->d_pack.Reserve(total));
	char *hp = static_cast<char *>(c->h_pack);
	char *dp = c->d_pack.As<char>();
	for (size_t r = 0; r < R; ++r) {
		memcpy(hp + o_data + r * dstride * sizeof(float), h.data + r * hds, n * sizeof(float));
		memcpy(hp + o_mask + r * dstride, h.mask + r * hms, n);
	}
	memset(hp + o_flags, 0, R * sizeof(int32_t));
	CUDA_TRY(cudaMemcpyAsync(dp, hp, o_flags + R * sizeof(int32_t), cudaMemcpyHostToDevice, stream));
	f.data = reinterpret_cast<float const *>(dp + o_data);
	f.mask = reinterpret_cast<uint8_t const *>(dp + o_mask);
	f.data_stride = dstride;
	f.mask_stride = dstride;
	f.coeff = h.coeff ? reinterpret_cast<double *>(dp + o_coeff) : nullptr;
	f.best_fit = h.best_fit ? reinterpret_cast<float *>(dp + o_best) : nullptr;
	f.residual = h.residual ? reinterpret_cast<float *>(dp + o_resid) : nullptr;
	f.final_mask = reinterpret_cast<uint8_t *>(dp + o_final);
	f.rms = reinterpret_cast<float *>(dp + o_rms);
	f.status = reinterpret_cast<int32_t *>(dp + o_status);
	f.lsq = reinterpret_cast<int32_t *>(dp + o_lsq);
	f.flags = reinterpret_cast<int32_t *>(dp + o_flags);
	f.boundary = spline ? reinterpret_cast<unsigned long long *>(dp + o_boundary) : nullptr;
	sakura_Status const st = LaunchFit(c, f, stream);
	if (st != sakura_Status_kOK) {
		(void) cudaStreamSynchronize(stream);
		return st;
	}
	CUDA_TRY(cudaMemcpyAsync(hp + o_flags, dp + o_flags, total - o_flags, cudaMemcpyDeviceToHost, stream));
	CUDA_TRY(cudaStreamSynchronize(stream));
	int32_t const *flags = reinterpret_cast<int32_t const *>(hp + o_flags);
	memcpy(h.status, hp + o_status, R * sizeof(int32_t));
	memcpy(h.lsq, hp + o_lsq, R * sizeof(int32_t));
	for (size_t r = 0; r < R; ++r) {
		int const fl = flags[r];
		if (fl & 1) {
			if (h.coeff != nullptr) {
				memcpy(h.coeff + r * f.coeff_stride, hp + o_coeff + r * f.coeff_stride * sizeof(double),
						f.coeff_stride * sizeof(double));
			}
			if (h.best_fit != nullptr) {
				memcpy(h.best_fit + r * hds, hp + o_best + r * dstride * sizeof(float), n * sizeof(float));
			}
			if (h.residual != nullptr) {
				memcpy(h.residual + r * hds, hp + o_resid + r * dstride * sizeof(float), n * sizeof(float));
			}
			memcpy(h.rms + r, hp + o_rms + r * sizeof(float), sizeof(float));
		}
		if (fl & 2) {
			memcpy(h.final_mask + r * hms, hp + o_final + r * dstride, n);
		}
		if (spline && (fl & 4)) {
			memcpy(h.boundary + r * nb, hp + o_boundary + r * nb * sizeof(unsigned long long),
					nb * sizeof(unsigned long long));
		}
	}
	return sakura_Status_kOK;
}

sakura_Status RunHostBatchSerial(Context *c, FitCall f, HostFitArgs const &h);

sakura_Status RunHostBatch(Context *c, FitCall f, HostFitArgs const &h) {
	if (f.num_spectra == 0) {
		return sakura_Status_kOK;
	}
	size_t const n = f.num_data;
	size_t const R = f.num_spectra;
	size_t const dstride = RoundUp(n, 32);
	DeviceInfo info;
	sakura_Status st = CurrentDevice(&info);
	if (st != sakura_Status_kOK) {
		return st;
	}
	// staged bytes per row: data, mask, final_mask, residual, best_fit (+ the small per-row outputs)
	size_t const row_bytes = dstride * (4 + 1 + 1 + (h.residual ? 4 : 0) + (h.best_fit ? 4 : 0))
			+ f.coeff_stride * sizeof(double) + 64;
	if (R * row_bytes <= PackedLimitBytes()) {
		return RunHostBatchPacked(c, f, h, info);
	}
	bool const fast = (f.type == kFitPolynomial || (f.type == kFitChebyshev && f.K <= 6))
			&& PolyFastSupported(n, f.K, dstride, dstride,
			nullptr, nullptr, nullptr, nullptr, nullptr);
	size_t chunk_rows = (ChunkMiB() << 20) / (dstride * sizeof(float));
	if (chunk_rows < 1) {
		chunk_rows = 1;
	}
	if (f.type == kFitSinusoid || (f.type == kFitChebyshev && !fast)) {
		// the tensor-core kernel fits tiles of kSinRowsPerTile rows, one tile per SM at a time:
		// whole waves per chunk
		chunk_rows = RoundUp(chunk_rows, static_cast<size_t>(kSinRowsPerTile) * info.num_sms);
	}
	if (R <= chunk_rows) {
		return RunHostBatchSerial(c, f, h);
	}
	return RunHostBatchPipelined(c, f, h, chunk_rows, fast, info);
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

// ---------------------------------------------------------------------------------------------
// ...MultiDevice entry points: rows are independent, so a batch in host memory is cut into one
// contiguous block of rows per device (block g = rows [R g / G, R (g+1) / G), the split of
// sakura_b200/sharding.py and of the reference's e2e driver, bench/e2eReduce/src/main.cc:520-608,
// which fans blocks of rows out to worker threads). Every device gets a replica of the context
// (same type / parameter / channel count; tables built once, kept for the context's lifetime) and a
// host thread that drives the replica's own copy/compute pipeline; there is no exchange between the
// devices and no collective.
// ---------------------------------------------------------------------------------------------
sakura_Status ResolveDevices(size_t num_devices, int const *device_ids, std::vector<int> *out) {
	int visible = 0;
	CUDA_TRY(cudaGetDeviceCount(&visible));
	if (visible < 1) {
		SetError("cudaGetDeviceCount", cudaErrorNoDevice);
		return sakura_Status_kUnknownError;
	}
	out->clear();
	if (num_devices == 0 || device_ids == nullptr) { // every visible device (the first num_devices of them)
		size_t const take = (num_devices == 0 || num_devices > static_cast<size_t>(visible)) ?
				static_cast<size_t>(visible) : num_devices;
		for (size_t g = 0; g < take; ++g) {
			out->push_back(static_cast<int>(g));
		}
		return sakura_Status_kOK;
	}
	for (size_t g = 0; g < num_devices; ++g) {
		CHECK_ARGS(0 <= device_ids[g] && device_ids[g] < visible);
		for (size_t k = 0; k < g; ++k) {
			CHECK_ARGS(device_ids[k] != device_ids[g]);
		}
		out->push_back(device_ids[g]);
	}
	return sakura_Status_kOK;
}

Context *ReplicaFor(Context *c, int device) {
	if (c->replicas == nullptr) {
		c->replicas = new std::vector<std::pair<int, Context *> >();
	}
	for (auto &r : *c->replicas) {
		if (r.first == device) {
			return r.second;
		}
	}
	Context *replica = nullptr;
	if (CreateContext(c->type, c->param, c->num_data, &replica) != sakura_Status_kOK) {
		return nullptr;
	}
	c->replicas->push_back(std::make_pair(device, replica));
	return replica;
}

sakura_Status RunHostBatchMultiDevice(Context *c, FitCall const &f, HostFitArgs const &h,
		size_t num_devices, int const *device_ids) {
	std::vector<int> devices;
	sakura_Status st = ResolveDevices(num_devices, device_ids, &devices);
	if (st != sakura_Status_kOK) {
		return st;
	}
	size_t const R = f.num_spectra;
	if (R == 0) {
		return sakura_Status_kOK;
	}
	size_t G = devices.size();
	if (G > R) {
		G = R;
	}
	size_t const nb = f.npiece + 1;
	struct Shard {
		int device;
		Context *replica;
		FitCall f;
		HostFitArgs h;
		sakura_Status status;
		std::string error;
		char const *kernel;
	};
	std::vector<Shard> shards(G);
	for (size_t g = 0; g < G; ++g) {
		Shard &s = shards[g];
		size_t const r0 = R / G * g + (R % G) * g / G;
		size_t const r1 = R / G * (g + 1) + (R % G) * (g + 1) / G;
		s.device = devices[g];
		s.replica = ReplicaFor(c, s.device);
		if (s.replica == nullptr) {
			return sakura_Status_kNoMemory;
		}
		s.f = f;
		s.f.num_spectra = r1 - r0;
		s.h.data = h.data + r0 * f.data_stride;
		s.h.mask = h.mask + r0 * f.mask_stride;
		s.h.coeff = h.coeff ? h.coeff + r0 * f.coeff_stride : nullptr;
		s.h.best_fit = h.best_fit ? h.best_fit + r0 * f.data_stride : nullptr;
		s.h.residual = h.residual ? h.residual + r0 * f.data_stride : nullptr;
		s.h.final_mask = h.final_mask + r0 * f.mask_stride;
		s.h.rms = h.rms + r0;
		s.h.status = h.status + r0;
		s.h.lsq = h.lsq + r0;
		s.h.boundary = h.boundary ? h.boundary + r0 * nb : nullptr;
		s.status = sakura_Status_kOK;
		s.kernel = "";
	}
	auto work = [](Shard *s) {
		cudaError_t const e = cudaSetDevice(s->device);
		if (e != cudaSuccess) {
			SetError("cudaSetDevice", e);
			s->status = sakura_Status_kUnknownError;
		} else {
			s->status = RunHostBatch(s->replica, s->f, s->h);
		}
		if (s->status != sakura_Status_kOK) {
			s->error = g_last_error;
		}
		s->kernel = g_last_kernel;
	};
	int previous = -1;
	(void) cudaGetDevice(&previous);
	if (G == 1) {
		work(&shards[0]);
	} else {
		std::vector<std::thread> threads;
		threads.reserve(G - 1);
		for (size_t g = 1; g < G; ++g) {
			threads.emplace_back(work, &shards[g]);
		}
		work(&shards[0]);
		for (auto &t : threads) {
			t.join();
		}
	}
	if (previous >= 0) {
		(void) cudaSetDevice(previous);
	}
	g_last_kernel = shards[0].kernel;
	for (size_t g = 0; g < G; ++g) {
		if (shards[g].status != sakura_Status_kOK) {
			snprintf(g_last_error, sizeof(g_last_error), "device %d: %s", shards[g].device,
					shards[g].error.c_str());
			return shards[g].status;
		}
	}
	return sakura_Status_kOK;
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

// ---------------------------------------------------------------- polynomial

namespace {
sakura_Status PolynomialBatchHost(struct sakura_LSQFitContextFloat const *context, uint16_t order,
		size_t num_spectra, size_t num_data, float const data[], size_t data_stride, bool const mask[],
		size_t mask_stride, float clip_threshold_sigma, uint16_t num_fitting_max, size_t num_coeff,
		double coeff[], float best_fit[], float residual[], bool final_mask[], float rms[],
		int32_t status[], int32_t lsqfit_status[], bool multi_device, size_t num_devices,
		int const device_ids[]) {
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
	if (multi_device) {
		return RunHostBatchMultiDevice(c, f, h, num_devices, device_ids);
	}
	return RunHostBatch(c, f, h);
}
} // namespace

sakura_Status sakura_LSQFitPolynomialFloatBatch(struct sakura_LSQFitContextFloat const *context,
		uint16_t order, size_t num_spectra, size_t num_data, float const data[], size_t data_stride,
		bool const mask[], size_t mask_stride, float clip_threshold_sigma, uint16_t num_fitting_max,
		size_t num_coeff, double coeff[], float best_fit[], float residual[], bool final_mask[],
		float rms[], int32_t status[], int32_t lsqfit_status[]) noexcept {
	return PolynomialBatchHost(context, order, num_spectra, num_data, data, data_stride, mask,
			mask_stride, clip_threshold_sigma, num_fitting_max, num_coeff, coeff, best_fit, residual,
			final_mask, rms, status, lsqfit_status, false, 0, nullptr);
}

sakura_Status sakura_LSQFitPolynomialFloatBatchMultiDevice(
		struct sakura_LSQFitContextFloat const *context, uint16_t order, size_t num_spectra,
		size_t num_data, float const data[], size_t data_stride, bool const mask[], size_t mask_stride,
		float clip_threshold_sigma, uint16_t num_fitting_max, size_t num_coeff, double coeff[],
		float best_fit[], float residual[], bool final_mask[], float rms[], int32_t status[],
		int32_t lsqfit_status[], size_t num_devices, int const device_ids[]) noexcept {
	return PolynomialBatchHost(context, order, num_spectra, num_data, data, data_stride, mask,
			mask_stride, clip_threshold_sigma, num_fitting_max, num_coeff, coeff, best_fit, residual,
			final_mask, rms, status, lsqfit_status, true, num_devices, device_ids);
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
	Context *c = const_cast<Context *>(context);
	{
		// residual / best_fit in place of data (the reference documents it, sakura.h:1755-1766): the
		// row-resident polynomial kernels read a row completely before they write any of it, so exact
		// aliasing of ONE of the two outputs is accepted where those kernels take the batch; the general
		// kernel re-reads data, and a partial overlap is never meaningful.
		size_t const span = (num_spectra == 0) ? 0 :
				((num_spectra - 1) * data_stride + num_data) * sizeof(float);
		auto overlaps = [&](void const *q) {
			if (q == nullptr || span == 0) {
				return false;
			}
			uintptr_t const a = reinterpret_cast<uintptr_t>(d_data), b = reinterpret_cast<uintptr_t>(q);
			return a < b + span && b < a + span;
		};
		bool const res_alias = overlaps(d_residual), best_alias = overlaps(d_best_fit);
		if (res_alias || best_alias) {
			bool const exact = (!res_alias || static_cast<void const *>(d_residual) == static_cast<void const *>(d_data))
					&& (!best_alias || static_cast<void const *>(d_best_fit) == static_cast<void const *>(d_data));
			bool const row_resident = (c->type == kFitPolynomial || (c->type == kFitChebyshev && num_coeff <= 6))
					&& cal == nullptr && PolyFastSupported(num_data, num_coeff, data_stride, mask_stride, d_data,
							d_mask, d_best_fit, d_residual, d_final_mask);
			CHECK_ARGS(exact && !(res_alias && best_alias) && row_resident);
		}
	}
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

namespace {
sakura_Status CubicSplineBatchHost(struct sakura_LSQFitContextFloat const *context, size_t num_pieces,
		size_t num_spectra, size_t num_data, float const data[], size_t data_stride, bool const mask[],
		size_t mask_stride, float clip_threshold_sigma, uint16_t num_fitting_max, double coeff[],
		float best_fit[], float residual[], bool final_mask[], float rms[], size_t boundary[],
		int32_t status[], int32_t lsqfit_status[], bool multi_device, size_t num_devices,
		int const device_ids[]) {
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
	if (multi_device) {
		return RunHostBatchMultiDevice(c, f, h, num_devices, device_ids);
	}
	return RunHostBatch(c, f, h);
}
} // namespace

sakura_Status sakura_LSQFitCubicSplineFloatBatch(struct sakura_LSQFitContextFloat const *context,
		size_t num_pieces, size_t num_spectra, size_t num_data, float const data[], size_t data_stride,
		bool const mask[], size_t mask_stride, float clip_threshold_sigma, uint16_t num_fitting_max,
		double coeff[], float best_fit[], float residual[], bool final_mask[], float rms[],
		size_t boundary[], int32_t status[], int32_t lsqfit_status[]) noexcept {
	return CubicSplineBatchHost(context, num_pieces, num_spectra, num_data, data, data_stride, mask,
			mask_stride, clip_threshold_sigma, num_fitting_max, coeff, best_fit, residual, final_mask,
			rms, boundary, status, lsqfit_status, false, 0, nullptr);
}

sakura_Status sakura_LSQFitCubicSplineFloatBatchMultiDevice(
		struct sakura_LSQFitContextFloat const *context, size_t num_pieces, size_t num_spectra,
		size_t num_data, float const data[], size_t data_stride, bool const mask[], size_t mask_stride,
		float clip_threshold_sigma, uint16_t num_fitting_max, double coeff[], float best_fit[],
		float residual[], bool final_mask[], float rms[], size_t boundary[], int32_t status[],
		int32_t lsqfit_status[], size_t num_devices, int const device_ids[]) noexcept {
	return CubicSplineBatchHost(context, num_pieces, num_spectra, num_data, data, data_stride, mask,
			mask_stride, clip_threshold_sigma, num_fitting_max, coeff, best_fit, residual, final_mask,
			rms, boundary, status, lsqfit_status, true, num_devices, device_ids);
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

namespace {
sakura_Status SinusoidBatchHost(struct sakura_LSQFitContextFloat const *context, size_t num_nwave,
		size_t const nwave[], size_t num_spectra, size_t num_data, float const data[],
		size_t data_stride, bool const mask[], size_t mask_stride, float clip_threshold_sigma,
		uint16_t num_fitting_max, size_t num_coeff, double coeff[], float best_fit[], float residual[],
		bool final_mask[], float rms[], int32_t status[], int32_t lsqfit_status[], bool multi_device,
		size_t num_devices, int const device_ids[]) {
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
	if (multi_device) {
		return RunHostBatchMultiDevice(c, f, h, num_devices, device_ids);
	}
	return RunHostBatch(c, f, h);
}
} // namespace

sakura_Status sakura_LSQFitSinusoidFloatBatch(struct sakura_LSQFitContextFloat const *context,
		size_t num_nwave, size_t const nwave[], size_t num_spectra, size_t num_data,
		float const data[], size_t data_stride, bool const mask[], size_t mask_stride,
		float clip_threshold_sigma, uint16_t num_fitting_max, size_t num_coeff, double coeff[],
		float best_fit[], float residual[], bool final_mask[], float rms[], int32_t status[],
		int32_t lsqfit_status[]) noexcept {
	return SinusoidBatchHost(context, num_nwave, nwave, num_spectra, num_data, data, data_stride, mask,
			mask_stride, clip_threshold_sigma, num_fitting_max, num_coeff, coeff, best_fit, residual,
			final_mask, rms, status, lsqfit_status, false, 0, nullptr);
}

sakura_Status sakura_LSQFitSinusoidFloatBatchMultiDevice(
		struct sakura_LSQFitContextFloat const *context, size_t num_nwave, size_t const nwave[],
		size_t num_spectra, size_t num_data, float const data[], size_t data_stride, bool const mask[],
		size_t mask_stride, float clip_threshold_sigma, uint16_t num_fitting_max, size_t num_coeff,
		double coeff[], float best_fit[], float residual[], bool final_mask[], float rms[],
		int32_t status[], int32_t lsqfit_status[], size_t num_devices, int const device_ids[]) noexcept {
	return SinusoidBatchHost(context, num_nwave, nwave, num_spectra, num_data, data, data_stride, mask,
			mask_stride, clip_threshold_sigma, num_fitting_max, num_coeff, coeff, best_fit, residual,
			final_mask, rms, status, lsqfit_status, true, num_devices, device_ids);
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

// ---------------------------------------------------------------- tuning

void sakura_b200_SetHostPipelineTuning(int64_t chunk_mib, int64_t packed_kib) noexcept {
	if (chunk_mib >= 1 && chunk_mib <= 4096) {
		g_chunk_mib.store(static_cast<size_t>(chunk_mib));
	}
	if (packed_kib >= 0 && packed_kib <= (1 << 20)) {
		g_packed_kib.store(static_cast<size_t>(packed_kib));
	}
}

sakura_Status sakura_b200_PackBoolMaskRows(size_t num_rows, size_t num_data, bool const mask[],
		size_t mask_stride, uint32_t bits[]) noexcept {
	CHECK_ARGS(num_data % 32 == 0 && mask_stride >= num_data);
	CHECK_ARGS(num_rows == 0 || (mask != nullptr && bits != nullptr));
	try {
		sakura_b200::PackMaskRows(reinterpret_cast<uint8_t const *>(mask), mask_stride, num_data, num_rows, bits);
	} catch (...) {
		return sakura_Status_kUnknownError;
	}
	return sakura_Status_kOK;
}

sakura_Status sakura_b200_UnpackBoolMaskRows(size_t num_rows, size_t num_data, uint32_t const bits[],
		bool mask[], size_t mask_stride, int32_t const row_flags[], int32_t flag_bit) noexcept {
	CHECK_ARGS(num_data % 32 == 0 && mask_stride >= num_data);
	CHECK_ARGS(num_rows == 0 || (mask != nullptr && bits != nullptr));
	try {
		sakura_b200::UnpackMaskRows(bits, num_data, num_rows, reinterpret_cast<uint8_t *>(mask), mask_stride,
				row_flags, flag_bit);
	} catch (...) {
		return sakura_Status_kUnknownError;
	}
	return sakura_Status_kOK;
}
