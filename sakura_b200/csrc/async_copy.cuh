// Bulk asynchronous global -> shared copies (cp.async.bulk, the 1-D TMA path; SASS UBLKCP) with
// mbarrier completion, and a two-stage producer/consumer ring shared by a pair of warps.
#ifndef SAKURA_B200_ASYNC_COPY_CUH_
#define SAKURA_B200_ASYNC_COPY_CUH_

#include <stdint.h>

namespace sakura_b200 {

__device__ __forceinline__ uint32_t smem_u32(void const *p) {
	return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}

__device__ __forceinline__ void mbar_init(uint64_t *bar, unsigned count) {
	asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(smem_u32(bar)), "r"(count) : "memory");
}

__device__ __forceinline__ void mbar_fence_init() {
	asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}

__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, unsigned bytes) {
	asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;"
			:: "r"(smem_u32(bar)), "r"(bytes) : "memory");
}

__device__ __forceinline__ void mbar_arrive(uint64_t *bar) {
	asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" :: "r"(smem_u32(bar)) : "memory");
}

__device__ __forceinline__ bool mbar_try_wait(uint64_t *bar, unsigned parity) {
	unsigned done;
	asm volatile("{\n\t"
			".reg .pred p;\n\t"
			"mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
			"selp.u32 %0, 1, 0, p;\n\t"
			"}" : "=r"(done) : "r"(smem_u32(bar)), "r"(parity) : "memory");
	return done != 0u;
}

__device__ __forceinline__ void mbar_wait(uint64_t *bar, unsigned parity) {
	while (!mbar_try_wait(bar, parity)) {
	}
}

/** Orders this thread's earlier generic-proxy accesses to shared memory before later async-proxy
 *  (bulk copy) accesses to the same locations. */
__device__ __forceinline__ void fence_proxy_async() {
	asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}

/** `bytes` (multiple of 16) from 16-byte aligned global `src` to 16-byte aligned shared `dst`;
 *  completion is signalled on `bar` as transaction bytes. */
__device__ __forceinline__ void bulk_copy_g2s(void *dst, void const *src, unsigned bytes, uint64_t *bar) {
	asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
			:: "r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}

// Two-stage ring of tiles for a pair of consumer warps; lane 0 of the pair's first warp produces.
// `it` counts the stage uses so far and advances identically in both warps. full[s] has one
// arrival (the producer's expect_tx) per use, empty[s] two (one per consumer warp).
struct PairPipe {
	uint64_t *full; // [2]
	uint64_t *empty; // [2]
	double *stage; // two stages of stage_doubles
	int stage_doubles;
	unsigned it;
	bool producer;
};

__device__ __forceinline__ void PipeInit(uint64_t *bars /*[4]*/, unsigned consumer_warps) {
	mbar_init(bars + 0, 1);
	mbar_init(bars + 1, 1);
	mbar_init(bars + 2, consumer_warps);
	mbar_init(bars + 3, consumer_warps);
}

/** Starts the copy for stage use `use` (the stage must have been released `use / 2` times). */
__device__ __forceinline__ void PipeIssue(PairPipe const &pp, unsigned use, void const *src,
		unsigned bytes, int lane) {
	if (pp.producer && lane == 0) {
		unsigned const st = use & 1u;
		unsigned const fill = use >> 1;
		if (fill > 0u) {
			mbar_wait(pp.empty + st, (fill - 1u) & 1u);
		}
		fence_proxy_async();
		mbar_expect_tx(pp.full + st, bytes);
		bulk_copy_g2s(pp.stage + st * pp.stage_doubles, src, bytes, pp.full + st);
	}
}

__device__ __forceinline__ double const *PipeWait(PairPipe const &pp, unsigned use) {
	unsigned const st = use & 1u;
	mbar_wait(pp.full + st, (use >> 1) & 1u);
	return pp.stage + st * pp.stage_doubles;
}

__device__ __forceinline__ void PipeRelease(PairPipe const &pp, unsigned use, int lane) {
	__syncwarp();
	if (lane == 0) {
		mbar_arrive(pp.empty + (use & 1u));
	}
}

} // namespace sakura_b200

#endif
