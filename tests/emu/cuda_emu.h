/*
 * cuda_emu.h -- a small CPU emulation of the CUDA subset used by cuda/*.cuh.
 *
 * TEST INFRASTRUCTURE ONLY.  The build container has no GPU; this header lets
 * tests/emu/emu_main.cpp compile the *same kernel source* (cuda/qoi_enc.cuh,
 * cuda/qoi_dec.cuh) with g++ and run it thread-by-thread on the CPU, so that
 * indexing, scan and decoupled-look-back logic (and deadlocks) can be debugged
 * before a GPU run.  It is never linked into libqoi_gpu.so and the product has
 * no CPU path: qoi_b200 / the C ABI fail if the CUDA library or device is missing.
 *
 * Model: every CUDA thread is a ucontext fiber on ONE OS thread.  Fibers switch
 * only at collectives (__syncthreads, warp-wide intrinsics) and in spin waits
 * (__nanosleep, mbarrier waits), in a seeded random order, so a missing barrier
 * or a reader running before its writer shows up as a wrong result, and a wait
 * that can never be satisfied is reported as a deadlock instead of hanging.
 * Restrictions (the kernels respect them): warp intrinsics are called with the
 * full mask by all 32 lanes; block size is a multiple of 32; threads do not exit
 * before their block's last barrier.
 */
#pragma once
#include <ucontext.h>

#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <vector>

#define __global__
#define __device__
#define __host__
#define __forceinline__ inline
#define __launch_bounds__(...)
#define __restrict__
#define __align__(n) __attribute__((aligned(n)))

struct uint2 { uint32_t x, y; };
struct __attribute__((aligned(16))) uint4 { uint32_t x, y, z, w; };
static inline uint4 make_uint4(uint32_t x, uint32_t y, uint32_t z, uint32_t w) { uint4 r = {x, y, z, w}; return r; }
static inline uint2 make_uint2(uint32_t x, uint32_t y) { uint2 r = {x, y}; return r; }

namespace qemu {

struct Idx { unsigned x, y, z; };

struct Warp {
	uint64_t slot[32];
	uint64_t result[32];
	uint32_t arrived = 0;
	uint32_t gen = 0;
};

struct Block {
	std::vector<unsigned char> smem_store;
	unsigned char *smem = nullptr;
	unsigned nthreads = 0;
	unsigned bar_count = 0;
	unsigned bar_gen = 0;
	std::vector<Warp> warps;
	Idx bidx{0, 0, 0};
	unsigned nb_count[16] = {0};
	unsigned nb_gen[16] = {0};
};

/* Context switch.  glibc's swapcontext makes a sigprocmask system call per switch, which was a third of the emulator's
   run time; on x86-64 a fiber switch is six callee-saved registers and the stack pointer.  Other hosts keep ucontext. */
#if defined(__x86_64__) && !defined(QEMU_USE_UCONTEXT)
#define QEMU_FAST_SWITCH 1
extern "C" void qemu_ctx_switch(void **save_sp, void *load_sp);
asm(".text\n"
    ".globl qemu_ctx_switch\n"
    ".type qemu_ctx_switch,@function\n"
    "qemu_ctx_switch:\n"
    "  pushq %rbp\n  pushq %rbx\n  pushq %r12\n  pushq %r13\n  pushq %r14\n  pushq %r15\n"
    "  movq %rsp, (%rdi)\n"
    "  movq %rsi, %rsp\n"
    "  popq %r15\n  popq %r14\n  popq %r13\n  popq %r12\n  popq %rbx\n  popq %rbp\n"
    "  ret\n"
    ".size qemu_ctx_switch, .-qemu_ctx_switch\n");
struct Ctx { void *sp = nullptr; };
#else
struct Ctx { ucontext_t uc; };
#endif

struct Fiber {
	Ctx ctx;
	void *stack = nullptr;
	bool done = false;
	Block *blk = nullptr;
	unsigned tid = 0;
	Idx tidx{0, 0, 0};
	const char *waiting = "";
	unsigned nb_phase[16] = {0}; /* how often this thread has arrived at each named barrier */
	unsigned nap = 0;            /* scheduler rounds this thread still sits out (models a warp that is not scheduled for a while) */
};

inline Fiber *cur = nullptr;
inline Ctx sched_ctx;
inline Idx g_blockdim{1, 1, 1}, g_griddim{1, 1, 1};
inline uint64_t progress = 0;
inline const std::function<void()> *g_body = nullptr;
inline uint64_t g_switches = 0;

inline void ctx_switch(Ctx &from, Ctx &to) {
#ifdef QEMU_FAST_SWITCH
	qemu_ctx_switch(&from.sp, to.sp);
#else
	swapcontext(&from.uc, &to.uc);
#endif
}

inline void yield(const char *why) {
	cur->waiting = why;
	g_switches++;
	ctx_switch(cur->ctx, sched_ctx);
	cur->waiting = "";
}

inline void trampoline() {
	(*g_body)();
	cur->done = true;
	progress++;
	ctx_switch(cur->ctx, sched_ctx); /* never resumed */
	abort();
}

inline const uint64_t *collect(unsigned mask, uint64_t v, const char *why) {
	Fiber *f = cur;
	Warp &w = f->blk->warps[f->tid / 32];
	unsigned lane = f->tid % 32;
	if (mask != 0xFFFFFFFFu) {
		fprintf(stderr, "cuda_emu: warp collective '%s' with partial mask %08x is not supported\n", why, mask);
		abort();
	}
	w.slot[lane] = v;
	w.arrived |= 1u << lane;
	if (w.arrived == mask) {
		memcpy(w.result, w.slot, sizeof w.slot);
		w.arrived = 0;
		w.gen++;
		progress++;
	} else {
		unsigned g = w.gen;
		while (w.gen == g) yield(why);
	}
	return w.result;
}

inline void block_barrier() {
	Block *b = cur->blk;
	b->bar_count++;
	if (b->bar_count == b->nthreads) {
		b->bar_count = 0;
		b->bar_gen++;
		progress++;
	} else {
		unsigned g = b->bar_gen;
		while (b->bar_gen == g) yield("__syncthreads");
	}
}

/* bar.sync / bar.arrive on barrier `id` with `nthreads` participants per phase */
inline void named_barrier(unsigned id, unsigned nthreads, bool wait) {
	Block *b = cur->blk;
	if (id >= 16) { fprintf(stderr, "cuda_emu: barrier id %u out of range\n", id); abort(); }
	unsigned g = b->nb_gen[id];
	/* a thread may only arrive for phase k once phase k-1 has completed: otherwise its arrival is counted towards the
	   unfinished phase and releases it early (a real hardware hazard, found the hard way) */
	if (cur->nb_phase[id] != g) {
		fprintf(stderr, "cuda_emu: BARRIER PHASE RACE on barrier %u: block %u thread %u arrives for phase %u while phase %u is still open\n",
		        id, b->bidx.x, cur->tid, cur->nb_phase[id], g);
		abort();
	}
	cur->nb_phase[id]++;
	b->nb_count[id]++;
	if (b->nb_count[id] > nthreads) { fprintf(stderr, "cuda_emu: barrier %u over-subscribed (%u > %u)\n", id, b->nb_count[id], nthreads); abort(); }
	if (b->nb_count[id] == nthreads) {
		b->nb_count[id] = 0;
		b->nb_gen[id]++;
		progress++;
		return;
	}
	if (wait)
		while (b->nb_gen[id] == g) yield("named barrier");
}

struct LaunchStats { uint64_t switches; };

/* run `body` as a grid of `grid` blocks x `block` threads with `smem_bytes` dynamic shared memory each */
inline LaunchStats launch(unsigned grid, unsigned block, size_t smem_bytes, const std::function<void()> &body,
                          unsigned seed = 1) {
	if (block % 32) { fprintf(stderr, "cuda_emu: block size must be a multiple of 32\n"); abort(); }
	const size_t STACK = 96 * 1024;
	std::vector<Block> blocks(grid);
	std::vector<Fiber> fibers((size_t)grid * block);
	g_blockdim = {block, 1, 1};
	g_griddim = {grid, 1, 1};
	g_body = &body;
	g_switches = 0;
	for (unsigned b = 0; b < grid; b++) {
		blocks[b].smem_store.assign(smem_bytes + 1024, 0xCD); /* poison: uninitialised shared memory is not zero */
		uintptr_t p = (uintptr_t)blocks[b].smem_store.data();
		blocks[b].smem = (unsigned char *)((p + 1023) & ~(uintptr_t)1023);
		blocks[b].nthreads = block;
		blocks[b].warps.resize(block / 32);
		blocks[b].bidx = {b, 0, 0};
		for (unsigned t = 0; t < block; t++) {
			Fiber &f = fibers[(size_t)b * block + t];
			f.blk = &blocks[b];
			f.tid = t;
			f.tidx = {t, 0, 0};
			f.stack = malloc(STACK);
#ifdef QEMU_FAST_SWITCH
			{ /* initial frame: six zeroed callee-saved registers, then `ret` into trampoline with rsp % 16 == 8 as after a call */
				void **sp = (void **)(((uintptr_t)f.stack + STACK) & ~(uintptr_t)15);
				*--sp = nullptr;              /* the return address trampoline would see (it never returns) */
				*--sp = (void *)&trampoline;
				for (int r = 0; r < 6; r++) *--sp = nullptr;
				f.ctx.sp = (void *)sp;
			}
#else
			getcontext(&f.ctx.uc);
			f.ctx.uc.uc_stack.ss_sp = f.stack;
			f.ctx.uc.uc_stack.ss_size = STACK;
			f.ctx.uc.uc_link = nullptr;
			makecontext(&f.ctx.uc, (void (*)())trampoline, 0);
#endif
		}
	}
	std::vector<size_t> order(fibers.size());
	for (size_t i = 0; i < order.size(); i++) order[i] = i;
	uint32_t rs = seed * 2654435761u + 12345u;
	size_t live = fibers.size();
	int idle_rounds = 0;
	while (live) {
		/* new random order each round (seed 0: plain order, useful for reproducing) */
		if (seed) {
			for (size_t i = order.size() - 1; i > 0; i--) {
				rs ^= rs << 13; rs ^= rs >> 17; rs ^= rs << 5;
				size_t j = rs % (i + 1);
				std::swap(order[i], order[j]);
			}
		}
		uint64_t before = progress;
		bool napping = false;
		for (size_t k = 0; k < order.size(); k++) {
			Fiber &f = fibers[order[k]];
			if (f.done) continue;
			if (f.nap) { /* arbitrary skew between threads: protocols must not rely on "everybody runs soon" */
				f.nap--;
				napping = true;
				continue;
			}
			cur = &f;
			ctx_switch(sched_ctx, f.ctx);
			if (f.done) live--;
			else if (seed) {
				rs ^= rs << 13; rs ^= rs >> 17; rs ^= rs << 5;
				if ((rs & 127u) == 0u) f.nap = 1 + (rs >> 8) % 48;
			}
		}
		if (progress == before) {
			if (napping) {
				for (auto &f : fibers) f.nap = 0; /* nothing moved: wake everybody before calling it a deadlock */
			} else if (++idle_rounds >= 4) {
				fprintf(stderr, "cuda_emu: DEADLOCK -- no thread made progress. Waiting threads:\n");
				int shown = 0;
				for (auto &f : fibers)
					if (!f.done && shown++ < 24)
						fprintf(stderr, "  block %u thread %u waits in %s\n", f.blk->bidx.x, f.tid, f.waiting);
				abort();
			}
		} else {
			idle_rounds = 0;
		}
	}
	for (auto &f : fibers) free(f.stack);
	cur = nullptr;
	return LaunchStats{g_switches};
}

}  // namespace qemu

#define threadIdx (qemu::cur->tidx)
#define blockIdx (qemu::cur->blk->bidx)
#define blockDim (qemu::g_blockdim)
#define gridDim (qemu::g_griddim)

/* ---- barriers and warp-wide intrinsics ---- */
static inline void __syncthreads() { qemu::block_barrier(); }
static inline void __syncwarp(unsigned mask = 0xFFFFFFFFu) { qemu::collect(mask, 0, "__syncwarp"); }

template <class T> static inline T __shfl_sync(unsigned mask, T v, int src) {
	static_assert(sizeof(T) <= 8, "");
	uint64_t raw = 0;
	memcpy(&raw, &v, sizeof(T));
	const uint64_t *all = qemu::collect(mask, raw, "__shfl_sync");
	T out;
	memcpy(&out, &all[src & 31], sizeof(T));
	return out;
}
template <class T> static inline T __shfl_up_sync(unsigned mask, T v, unsigned delta) {
	uint64_t raw = 0;
	memcpy(&raw, &v, sizeof(T));
	const uint64_t *all = qemu::collect(mask, raw, "__shfl_up_sync");
	unsigned lane = qemu::cur->tid % 32;
	T out = v;
	if (lane >= delta) memcpy(&out, &all[lane - delta], sizeof(T));
	return out;
}
template <class T> static inline T __shfl_down_sync(unsigned mask, T v, unsigned delta) {
	uint64_t raw = 0;
	memcpy(&raw, &v, sizeof(T));
	const uint64_t *all = qemu::collect(mask, raw, "__shfl_down_sync");
	unsigned lane = qemu::cur->tid % 32;
	T out = v;
	if (lane + delta < 32) memcpy(&out, &all[lane + delta], sizeof(T));
	return out;
}
template <class T> static inline T __shfl_xor_sync(unsigned mask, T v, int x) {
	uint64_t raw = 0;
	memcpy(&raw, &v, sizeof(T));
	const uint64_t *all = qemu::collect(mask, raw, "__shfl_xor_sync");
	unsigned lane = qemu::cur->tid % 32;
	T out;
	memcpy(&out, &all[(lane ^ x) & 31], sizeof(T));
	return out;
}
static inline unsigned __ballot_sync(unsigned mask, int pred) {
	const uint64_t *all = qemu::collect(mask, pred ? 1 : 0, "__ballot_sync");
	unsigned r = 0;
	for (int i = 0; i < 32; i++) r |= (unsigned)(all[i] & 1) << i;
	return r;
}
static inline unsigned __match_any_sync(unsigned mask, unsigned v) {
	const uint64_t *all = qemu::collect(mask, v, "__match_any_sync");
	unsigned r = 0;
	for (int i = 0; i < 32; i++) r |= (unsigned)((uint32_t)all[i] == v) << i;
	return r;
}
static inline unsigned __reduce_or_sync(unsigned mask, unsigned v) {
	const uint64_t *all = qemu::collect(mask, v, "__reduce_or_sync");
	unsigned r = 0;
	for (int i = 0; i < 32; i++) r |= (unsigned)all[i];
	return r;
}
static inline unsigned __reduce_add_sync(unsigned mask, unsigned v) {
	const uint64_t *all = qemu::collect(mask, v, "__reduce_add_sync");
	unsigned r = 0;
	for (int i = 0; i < 32; i++) r += (unsigned)all[i];
	return r;
}
static inline int __reduce_max_sync(unsigned mask, int v) {
	const uint64_t *all = qemu::collect(mask, (uint64_t)(uint32_t)v, "__reduce_max_sync");
	int r = (int)(uint32_t)all[0];
	for (int i = 1; i < 32; i++) r = (int)(uint32_t)all[i] > r ? (int)(uint32_t)all[i] : r;
	return r;
}
static inline int __reduce_min_sync(unsigned mask, int v) {
	const uint64_t *all = qemu::collect(mask, (uint64_t)(uint32_t)v, "__reduce_min_sync");
	int r = (int)(uint32_t)all[0];
	for (int i = 1; i < 32; i++) r = (int)(uint32_t)all[i] < r ? (int)(uint32_t)all[i] : r;
	return r;
}

/* ---- scalar intrinsics ---- */
static inline int __popc(unsigned v) { return __builtin_popcount(v); }
static inline int __popcll(unsigned long long v) { return __builtin_popcountll(v); }
static inline int __clz(int v) { return v ? __builtin_clz((unsigned)v) : 32; }
static inline int __ffs(int v) { return __builtin_ffs(v); }
static inline int __ffsll(long long v) { return __builtin_ffsll(v); }
static inline unsigned __byte_perm(unsigned a, unsigned b, unsigned s) {
	uint64_t both = ((uint64_t)b << 32) | a;
	unsigned r = 0;
	for (int i = 0; i < 4; i++) {
		unsigned sel = (s >> (4 * i)) & 0xF;
		unsigned byte = (unsigned)(both >> (8 * (sel & 7))) & 0xFF;
		if (sel & 8) byte = (byte & 0x80) ? 0xFF : 0x00;
		r |= byte << (8 * i);
	}
	return r;
}
static inline unsigned __funnelshift_r(unsigned lo, unsigned hi, unsigned sh) {
	uint64_t both = ((uint64_t)hi << 32) | lo;
	return (unsigned)(both >> (sh & 31));
}
static inline unsigned __dp4a(unsigned a, unsigned b, unsigned c) {
	for (int i = 0; i < 4; i++) c += ((a >> (8 * i)) & 0xFF) * ((b >> (8 * i)) & 0xFF);
	return c;
}
static inline int __dp4a(int a, int b, int c) { /* signed bytes x signed bytes */
	for (int i = 0; i < 4; i++) c += (int)(int8_t)(a >> (8 * i)) * (int)(int8_t)(b >> (8 * i));
	return c;
}
static inline unsigned __vsub4(unsigned a, unsigned b) {
	unsigned r = 0;
	for (int i = 0; i < 4; i++) r |= ((((a >> (8 * i)) & 0xFF) - ((b >> (8 * i)) & 0xFF)) & 0xFF) << (8 * i);
	return r;
}
static inline unsigned __vadd4(unsigned a, unsigned b) {
	unsigned r = 0;
	for (int i = 0; i < 4; i++) r |= ((((a >> (8 * i)) & 0xFF) + ((b >> (8 * i)) & 0xFF)) & 0xFF) << (8 * i);
	return r;
}
static inline unsigned __umulhi(unsigned a, unsigned b) { return (unsigned)(((uint64_t)a * b) >> 32); }

/* ---- memory ---- */
template <class T> static inline T __ldcg(const T *p) { return *p; }
template <class T> static inline T __ldg(const T *p) { return *p; }
template <class T> static inline void __stcg(T *p, T v) { *p = v; qemu::progress++; }
static inline void __threadfence() {}
static inline void __threadfence_block() {}
static inline void __nanosleep(unsigned) { qemu::yield("spin wait"); }

static inline unsigned atomicAdd(unsigned *p, unsigned v) { unsigned o = *p; *p = o + v; qemu::progress++; return o; }
static inline int atomicAdd(int *p, int v) { int o = *p; *p = o + v; qemu::progress++; return o; }
static inline unsigned atomicOr(unsigned *p, unsigned v) { unsigned o = *p; *p = o | v; qemu::progress++; return o; }
static inline unsigned atomicExch(unsigned *p, unsigned v) { unsigned o = *p; *p = v; qemu::progress++; return o; }
static inline int atomicMin(int *p, int v) { int o = *p; if (v < o) *p = v; qemu::progress++; return o; }
static inline int atomicMax(int *p, int v) { int o = *p; if (v > o) *p = v; qemu::progress++; return o; }
