// See host_bits.h. Plain C++ (compiled by the host compiler, not nvcc); the AVX2 bodies are selected at
// run time, the portable ones work on 8 channels per 64-bit word.
#include "host_bits.h"

#include <immintrin.h>
#include <stdlib.h>
#include <string.h>

#include <atomic>
#include <condition_variable>
#include <mutex>
#include <thread>
#include <vector>

namespace sakura_b200 {

namespace {

// ---- one row ------------------------------------------------------------------------------------
void PackRowPortable(uint8_t const *src, size_t n, uint32_t *dst) {
	for (size_t w = 0; w < n / 32; ++w) {
		uint32_t out = 0;
		for (int q = 0; q < 4; ++q) {
			uint64_t v;
			memcpy(&v, src + 32 * w + 8 * q, 8);
			// any non-zero byte -> its bit 7 (carry-free: the low seven bits are added separately)
			uint64_t const nz = ((v & 0x7f7f7f7f7f7f7f7fULL) + 0x7f7f7f7f7f7f7f7fULL) | v;
			uint64_t const m = (nz >> 7) & 0x0101010101010101ULL;
			out |= static_cast<uint32_t>((m * 0x0102040810204080ULL) >> 56) << (8 * q);
		}
		dst[w] = out;
	}
}

void UnpackRowPortable(uint32_t const *src, size_t n, uint8_t *dst) {
	for (size_t w = 0; w < n / 32; ++w) {
		uint32_t const b = src[w];
		for (int k = 0; k < 32; ++k) {
			dst[32 * w + k] = static_cast<uint8_t>((b >> k) & 1u);
		}
	}
}

__attribute__((target("avx2"))) void PackRowAvx2(uint8_t const *src, size_t n, uint32_t *dst) {
	__m256i const zero = _mm256_setzero_si256();
	for (size_t w = 0; w < n / 32; ++w) {
		__m256i const v = _mm256_loadu_si256(reinterpret_cast<__m256i const *>(src + 32 * w));
		dst[w] = ~static_cast<uint32_t>(_mm256_movemask_epi8(_mm256_cmpeq_epi8(v, zero)));
	}
}

__attribute__((target("avx2"))) void UnpackRowAvx2(uint32_t const *src, size_t n, uint8_t *dst) {
	bool const stream = (reinterpret_cast<uintptr_t>(dst) % 32) == 0;
	__m256i const shuf = _mm256_setr_epi8(0, 0, 0, 0, 0, 0, 0, 0, 1, 1, 1, 1, 1, 1, 1, 1, 2, 2, 2, 2, 2, 2, 2, 2, 3, 3,
			3, 3, 3, 3, 3, 3);
	__m256i const bit = _mm256_set1_epi64x(static_cast<long long>(0x8040201008040201ULL));
	__m256i const one = _mm256_set1_epi8(1);
	for (size_t w = 0; w < n / 32; ++w) {
		__m256i v = _mm256_set1_epi32(static_cast<int>(src[w]));
		v = _mm256_shuffle_epi8(v, shuf);
		v = _mm256_and_si256(_mm256_cmpeq_epi8(_mm256_and_si256(v, bit), bit), one);
		if (stream) { // whole rows are written and not read again soon: no read-for-ownership of the lines
			_mm256_stream_si256(reinterpret_cast<__m256i *>(dst + 32 * w), v);
		} else {
			_mm256_storeu_si256(reinterpret_cast<__m256i *>(dst + 32 * w), v);
		}
	}
}

bool HaveAvx2() {
	static bool const have = __builtin_cpu_supports("avx2");
	return have;
}

// ---- worker pool ----------------------------------------------------------------------------------
class Pool {
public:
	Pool() {
		unsigned hw = std::thread::hardware_concurrency();
		unsigned want = hw >= 16 ? 6 : (hw >= 8 ? 3 : (hw >= 4 ? 1 : 0));
		// SAKURA_B200_HOST_THREADS = 0 .. 64: worker threads of the mask codec (0: the calling threads do it
		// themselves), for hosts whose callers bring their own thread pools
		if (char const *e = getenv("SAKURA_B200_HOST_THREADS")) {
			char *end = nullptr;
			long const v = strtol(e, &end, 10);
			if (end != e && v >= 0 && v <= 64) {
				want = static_cast<unsigned>(v);
			}
		}
		for (unsigned i = 0; i < want; ++i) {
			workers_.emplace_back([this] { Work(); });
		}
	}
	~Pool() {
		{
			std::lock_guard<std::mutex> lock(mu_);
			stop_ = true;
		}
		cv_.notify_all();
		for (auto &t : workers_) {
			t.join();
		}
	}
	size_t Workers() const {
		return workers_.size();
	}
	// one job at a time per caller; several callers (one pipeline thread per GPU) share the workers
	void Run(size_t count, size_t min_per_task, std::function<void(size_t, size_t)> const &fn) {
		size_t tasks = workers_.size() + 1;
		if (min_per_task > 0 && count / min_per_task < tasks) {
			tasks = count / min_per_task;
		}
		if (tasks <= 1) {
			fn(0, count);
			return;
		}
		Job job;
		job.fn = &fn;
		job.count = count;
		job.tasks = tasks;
		job.next.store(1);
		job.left.store(tasks);
		{
			std::lock_guard<std::mutex> lock(mu_);
			jobs_.push_back(&job);
		}
		cv_.notify_all();
		RunTask(job, 0);
		// help with the remaining tasks of the own job, then wait for those taken by workers
		for (;;) {
			size_t const t = job.next.fetch_add(1);
			if (t >= job.tasks) {
				break;
			}
			RunTask(job, t);
		}
		{
			std::unique_lock<std::mutex> lock(mu_);
			for (auto it = jobs_.begin(); it != jobs_.end(); ++it) {
				if (*it == &job) {
					jobs_.erase(it);
					break;
				}
			}
			done_cv_.wait(lock, [&] { return job.left.load() == 0; });
		}
	}

private:
	struct Job {
		std::function<void(size_t, size_t)> const *fn;
		size_t count;
		size_t tasks;
		std::atomic<size_t> next;
		std::atomic<size_t> left;
	};
	void RunTask(Job &job, size_t t) {
		size_t const lo = job.count * t / job.tasks;
		size_t const hi = job.count * (t + 1) / job.tasks;
		(*job.fn)(lo, hi);
		if (job.left.fetch_sub(1) == 1) {
			std::lock_guard<std::mutex> lock(mu_);
			done_cv_.notify_all();
		}
	}
	void Work() {
		std::unique_lock<std::mutex> lock(mu_);
		for (;;) {
			Job *job = nullptr;
			size_t t = 0;
			for (Job *j : jobs_) {
				size_t const cand = j->next.load();
				if (cand < j->tasks) {
					t = j->next.fetch_add(1);
					if (t < j->tasks) {
						job = j;
						break;
					}
				}
			}
			if (job != nullptr) {
				lock.unlock();
				RunTask(*job, t);
				lock.lock();
				continue;
			}
			if (stop_) {
				return;
			}
			cv_.wait(lock);
		}
	}
	std::mutex mu_;
	std::condition_variable cv_, done_cv_;
	std::vector<Job *> jobs_;
	std::vector<std::thread> workers_;
	bool stop_ = false;
};

Pool &ThePool() {
	static Pool *pool = new Pool(); // (leaked on purpose: no join at process exit)
	return *pool;
}

} // namespace

void HostParallelFor(size_t count, size_t min_per_task, std::function<void(size_t, size_t)> const &fn) {
	if (count == 0) {
		return;
	}
	ThePool().Run(count, min_per_task, fn);
}

void PackMaskRows(uint8_t const *bytes, size_t byte_stride, size_t n, size_t rows, uint32_t *bits) {
	bool const avx2 = HaveAvx2();
	size_t const words = n / 32;
	HostParallelFor(rows, 64, [&](size_t lo, size_t hi) {
		for (size_t r = lo; r < hi; ++r) {
			if (avx2) {
				PackRowAvx2(bytes + r * byte_stride, n, bits + r * words);
			} else {
				PackRowPortable(bytes + r * byte_stride, n, bits + r * words);
			}
		}
	});
}

void UnpackMaskRows(uint32_t const *bits, size_t n, size_t rows, uint8_t *bytes, size_t byte_stride,
		int32_t const *row_flags, int32_t flag_bit) {
	bool const avx2 = HaveAvx2();
	size_t const words = n / 32;
	HostParallelFor(rows, 64, [&](size_t lo, size_t hi) {
		for (size_t r = lo; r < hi; ++r) {
			if (row_flags != nullptr && (row_flags[r] & flag_bit) == 0) {
				continue;
			}
			if (avx2) {
				UnpackRowAvx2(bits + r * words, n, bytes + r * byte_stride);
			} else {
				UnpackRowPortable(bits + r * words, n, bytes + r * byte_stride);
			}
		}
		if (avx2) {
			_mm_sfence(); // streaming stores of this task visible before the call returns
		}
	});
}

} // namespace sakura_b200
