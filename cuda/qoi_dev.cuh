/*
 * qoi_dev.cuh -- the handful of sm_100a primitives the kernels are written on:
 * mbarrier + bulk async copies (TMA, SASS UBLKCP), acquire/release flag
 * accesses for the decoupled look-back, and byte helpers.
 *
 * When QOI_EMU is defined (tests/emu only) the same names are provided on top
 * of tests/emu/cuda_emu.h so the kernel source can be stepped on the CPU.
 */
#pragma once
#include <stdint.h>

#ifdef QOI_EMU
#include "cuda_emu.h"
#else
#include <cuda_runtime.h>
#endif

namespace qoidev {

#define QOI_FULL_MASK 0xFFFFFFFFu

#ifndef QOI_EMU
/* ------------------------------------------------------------------ real */
__device__ __forceinline__ uint32_t smem_addr(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t arrivals) {
	asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_addr(bar)), "r"(arrivals) : "memory");
}
/* make the initialised barriers visible to the async (TMA) proxy */
__device__ __forceinline__ void mbar_fence_init() {
	asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
	asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
/* one arrival that also announces `bytes` of bulk-copy traffic for the current phase */
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t *bar, uint32_t bytes) {
	asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_addr(bar)), "r"(bytes) : "memory");
}
/* the two halves of arrive.expect_tx: announce the bytes first (so that the bulk copy can be issued at once), arrive later
   (release: everything this thread -- and, after a __syncwarp, its warp -- wrote to shared memory before is visible to the waiters) */
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes) {
	asm volatile("mbarrier.expect_tx.relaxed.cta.shared::cta.b64 [%0], %1;" ::"r"(smem_addr(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t *bar) {
	asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_addr(bar)) : "memory");
}
/* pull one 128-byte line into L2 (no register result, no waiting) */
__device__ __forceinline__ void prefetch_l2(const void *p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }
__device__ __forceinline__ bool mbar_try_wait(uint64_t *bar, uint32_t parity) {
	uint32_t ok;
	asm volatile(
	    "{\n\t.reg .pred p;\n\t"
	    "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\t" /* %3: let the warp stay suspended up to this many ns instead of spinning (round 2: 13 % of the encoder's issued instructions were this loop) */
	    "selp.u32 %0, 1, 0, p;\n\t}"
	    : "=r"(ok)
	    : "r"(smem_addr(bar)), "r"(parity), "r"(200000u)
	    : "memory");
	return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
	while (!mbar_try_wait(bar, parity)) {
	}
}
/* global -> shared bulk copy (1-D TMA, tma_load_1d_hint below): both addresses 16-byte aligned, bytes a multiple of 16, bytes > 0.
   L2 eviction policy for data that is touched exactly once (pixels in, streams out): keeps the look-back records,
   the restart index and the instruction stream resident in L2 while gigabytes stream through */
__device__ __forceinline__ uint64_t l2_evict_first_policy() {
	uint64_t pol;
	asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
	return pol;
}
__device__ __forceinline__ void tma_load_1d_hint(void *dst_smem, const void *src_gmem, uint32_t bytes, uint64_t *bar, uint64_t pol) {
	asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;" ::"r"(smem_addr(dst_smem)),
	             "l"(src_gmem), "r"(bytes), "r"(smem_addr(bar)), "l"(pol)
	             : "memory");
}
/* streaming (evict-first) 16-byte global store (the encoder's copy-out) */
__device__ __forceinline__ void st_stream_u4(uint4 *p, uint4 v) {
	asm volatile("st.global.cs.v4.u32 [%0], {%1, %2, %3, %4};" ::"l"(p), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}
/* 256-bit accesses (sm_100: LDG.E.256 / STG.E.256): one lane moves a whole 32-byte sector with one instruction.  The decoders'
   lanes each stream their own segment and their own tile, so a warp-wide access touches 32 different lines whatever its width
   -- and costs the L1 data pipe one pass per line: twice the width is half the passes (profiles/r2_dec_lut: that pipe, at 70 %,
   is what bounds the decoder).  32-byte aligned addresses. */
__device__ __forceinline__ void ld_nc_256(const void *p, uint4 &a, uint4 &b) {
	asm volatile("ld.global.nc.v8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
	             : "=r"(a.x), "=r"(a.y), "=r"(a.z), "=r"(a.w), "=r"(b.x), "=r"(b.y), "=r"(b.z), "=r"(b.w)
	             : "l"(p));
}
__device__ __forceinline__ void ld_256(const void *p, uint4 &a, uint4 &b) {
	asm volatile("ld.global.v8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
	             : "=r"(a.x), "=r"(a.y), "=r"(a.z), "=r"(a.w), "=r"(b.x), "=r"(b.y), "=r"(b.z), "=r"(b.w)
	             : "l"(p)
	             : "memory");
}
__device__ __forceinline__ void st_256(void *p, uint4 a, uint4 b) {
	asm volatile("st.global.v8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"l"(p), "r"(a.x), "r"(a.y), "r"(a.z), "r"(a.w), "r"(b.x), "r"(b.y), "r"(b.z),
	             "r"(b.w)
	             : "memory");
}
__device__ __forceinline__ uint64_t ld_relaxed_u64(const uint64_t *p) {
	uint64_t v;
	asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
	return v;
}
__device__ __forceinline__ void st_relaxed_u64(uint64_t *p, uint64_t v) {
	asm volatile("st.relaxed.gpu.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
/* named barriers: producer/consumer hand-off between the scout warp and the compute warps.  The non-.aligned
   forms are used because lanes of a warp may reach them along different paths (deferred / patched pixels). */
__device__ __forceinline__ void bar_sync(uint32_t id, uint32_t nthreads) {
	asm volatile("barrier.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}
__device__ __forceinline__ void bar_arrive(uint32_t id, uint32_t nthreads) {
	__threadfence_block();
	asm volatile("barrier.arrive %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}
__device__ __forceinline__ void spin_pause() { __nanosleep(64); }
__device__ __forceinline__ uint32_t lane_id() { return threadIdx.x & 31u; }
__device__ __forceinline__ uint32_t lanemask_lt() {
	uint32_t m;
	asm("mov.u32 %0, %%lanemask_lt;" : "=r"(m));
	return m;
}
__device__ __forceinline__ uint32_t lanemask_gt() {
	uint32_t m;
	asm("mov.u32 %0, %%lanemask_gt;" : "=r"(m));
	return m;
}
#define QOI_DYNAMIC_SMEM(name) extern __shared__ __align__(1024) unsigned char name[]

#else
/* ------------------------------------------------------------------ emulation */
struct EmuBar {
	uint8_t phase, count;
	int16_t arrivals_left;
	int32_t tx;
};
static_assert(sizeof(EmuBar) == 8, "");
static inline void emubar_check(EmuBar *b) {
	if (b->arrivals_left == 0 && b->tx == 0) {
		b->phase ^= 1;
		b->arrivals_left = b->count;
		qemu::progress++;
	}
}
static inline void mbar_init(uint64_t *bar, uint32_t arrivals) {
	EmuBar *b = (EmuBar *)bar;
	b->phase = 0;
	b->count = (uint8_t)arrivals;
	b->arrivals_left = (int16_t)arrivals;
	b->tx = 0;
}
static inline void mbar_fence_init() {}
static inline void mbar_arrive_expect_tx(uint64_t *bar, uint32_t bytes) {
	EmuBar *b = (EmuBar *)bar;
	b->tx += (int32_t)bytes;
	b->arrivals_left--;
	emubar_check(b);
}
static inline void mbar_expect_tx(uint64_t *bar, uint32_t bytes) { ((EmuBar *)bar)->tx += (int32_t)bytes; }
static inline void mbar_arrive(uint64_t *bar) {
	EmuBar *b = (EmuBar *)bar;
	b->arrivals_left--;
	emubar_check(b);
}
static inline void prefetch_l2(const void *) {}
static inline bool mbar_try_wait(uint64_t *bar, uint32_t parity) { return ((EmuBar *)bar)->phase != (uint8_t)parity; }
static inline void mbar_wait(uint64_t *bar, uint32_t parity) {
	while (!mbar_try_wait(bar, parity)) qemu::yield("mbarrier wait");
}
static inline void emu_check_bulk(const void *a, const void *b, uint32_t bytes) {
	if (((uintptr_t)a & 15) || ((uintptr_t)b & 15) || (bytes & 15) || bytes == 0) {
		fprintf(stderr, "cuda_emu: bulk copy violates the 16-byte rules: %p %p %u\n", a, b, bytes);
		abort();
	}
}
static inline void tma_load_1d(void *dst_smem, const void *src_gmem, uint32_t bytes, uint64_t *bar) {
	emu_check_bulk(dst_smem, src_gmem, bytes);
	memcpy(dst_smem, src_gmem, bytes);
	EmuBar *b = (EmuBar *)bar;
	b->tx -= (int32_t)bytes;
	emubar_check(b);
}
static inline uint64_t l2_evict_first_policy() { return 0; }
static inline void tma_load_1d_hint(void *d, const void *s_, uint32_t n, uint64_t *bar, uint64_t) { tma_load_1d(d, s_, n, bar); }
static inline void st_stream_u4(uint4 *p, uint4 v) { *p = v; }
static inline void ld_nc_256(const void *p, uint4 &a, uint4 &b) { a = ((const uint4 *)p)[0]; b = ((const uint4 *)p)[1]; }
static inline void ld_256(const void *p, uint4 &a, uint4 &b) { a = ((const uint4 *)p)[0]; b = ((const uint4 *)p)[1]; }
static inline void st_256(void *p, uint4 a, uint4 b) { ((uint4 *)p)[0] = a; ((uint4 *)p)[1] = b; }
static inline uint64_t ld_relaxed_u64(const uint64_t *p) { return *p; }
static inline void st_relaxed_u64(uint64_t *p, uint64_t v) { *p = v; qemu::progress++; }
static inline void bar_sync(uint32_t id, uint32_t nthreads) { qemu::named_barrier(id, nthreads, true); }
static inline void bar_arrive(uint32_t id, uint32_t nthreads) { qemu::named_barrier(id, nthreads, false); }
static inline void spin_pause() { qemu::yield("spin wait"); }
static inline uint32_t lane_id() { return threadIdx.x & 31u; }
static inline uint32_t lanemask_lt() { return (1u << lane_id()) - 1u; }
static inline uint32_t lanemask_gt() { return lane_id() == 31 ? 0u : ~((2u << lane_id()) - 1u); }
#define QOI_DYNAMIC_SMEM(name) unsigned char *name = qemu::cur->blk->smem
#endif

/* ------------------------------------------------------------------ common helpers */

/* QOI_COLOR_HASH & 63 (reference qoi.h:322,430,577) on a packed r|g<<8|b<<16|a<<24 pixel: one IDP4A */
__device__ __forceinline__ uint32_t qoi_slot(uint32_t px) { return __dp4a(px, 0x0B070503u, 0u) & 63u; }

}  // namespace qoidev
