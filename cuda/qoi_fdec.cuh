/*
 * qoi_fdec.cuh -- parallel decode of ARBITRARY ("foreign") QOI streams on sm_100a.
 *
 * The reference decoder (qoi.h:540-587) chains four things through the stream: the read cursor p, the
 * current colour px, the run counter and index[64].  For a stream we did not write there is no restart
 * index, so the state at any cut has to be *found*:
 *
 *   1. parse (exact, parallel).  The stream is cut into segments of QOI_FSEG_BYTES bytes.  An op's
 *      length depends on its first byte only (0xFE:4, 0xFF:5, 10xxxxxx:2, else 1; qoi.h:547-575), so for
 *      each of the 5 possible distances from the segment start to its first op the segment maps
 *      "entry skip -> (exit skip, pixels produced)".  fparse_kernel tabulates that map; fchain_kernel
 *      composes the maps along each stream and leaves every segment's true first op and first pixel.
 *   2. speculate (parallel).  fdecode_kernel decodes every segment from a GUESSED state (px, index[])
 *      and records what the result actually depended on: which index[] slots were read before the
 *      segment itself wrote them, and whether the incoming colour / alpha was used before being replaced.
 *   3. verify (sequential per stream, cheap: no pixels, 64 lanes' worth of state).  fverify_kernel walks
 *      the segments in order with the EXACT state.  A segment whose guesses agree with the exact state on
 *      everything it depended on is final, and its writes advance the exact state.  At the first
 *      disagreement that segment gets the exact state as its new guess; behind it the chain continues in
 *      "guess mode" so that later segments get better guesses (a purely relative segment's output colour
 *      is re-based: out = new_in + (old_out - old_in)).  Only segments whose guess changed on something
 *      they depend on are decoded again.
 *   4. fix up (exact).  After QOI_FDEC_ROUNDS speculative launches ffix_kernel makes the same in-order walk
 *      once more, but decodes the segments whose guess still does not hold on the spot (one lane, the
 *      reference loop), so the chain state stays exact to the end of the stream.  The result is therefore
 *      exact for every input; the sequential work is proportional to what speculation could not settle
 *      (streams whose colours are almost all INDEX / relative ops degrade to the sequential decoder).
 *
 * Every kernel cites the reference rule it restates; nothing here reads outside [stream, stream+size).
 */
#pragma once
#include "qoi_common.h"
#include "qoi_dev.cuh"

namespace qoifdec {
using namespace qoidev;

#define QOI_FSEG_BYTES 4096
#define QOI_FDEC_ROUNDS 6 /* speculative decode launches (with a verify pass between them) before the exact fix-up */
#define QOI_FDEC_THREADS 128
#define QOI_FCHAIN_SMEM ((32 * 5 + 32 * 2 + 32 + 32 + 4) * 4)
#define QOI_FDEC_SMEM ((64 + 32) * QOI_FDEC_THREADS * 4) /* per lane: index[64] + 64 source half-words */
/* Where a channel group (r,g,b or alpha) of a colour came from, as tracked by the speculative decode: values 0..63 = slot
   of the guessed index[] that was read before the segment wrote it (r,g,b possibly plus DIFF/LUMA deltas) */
#define QOI_SRC_IN 64u    /* the colour that entered the segment (r,g,b possibly plus deltas) */
#define QOI_SRC_OWN 65u   /* bytes of the segment itself: independent of every guess */
#define QOI_SRC_NONE 0xFFu

struct QoiFState {
	/* geometry (host written) */
	const uint32_t *seg_base; /* [n_images + 1] first global segment of each image (capacity, not live count) */
	/* parse */
	uint32_t *px_cnt;         /* [segs][5] pixels produced when entered with skip s */
	uint8_t *exit_skip;       /* [segs][8] (5 used) */
	/* chain result */
	uint8_t *entry;           /* [segs] distance from the segment start to its first op */
	uint32_t *px_start;       /* [segs] first pixel of the segment */
	uint32_t *n_live;         /* [n_images] segments that hold at least one op of a pixel < npx */
	/* speculation */
	uint32_t *in_px;          /* [segs] guessed colour before the segment */
	uint32_t *in_tab;         /* [segs][64] guessed index[] */
	uint32_t *out_px;         /* [segs] */
	uint32_t *out_tab;        /* [segs][64] values written by the segment (last writer per slot) */
	uint64_t *out_written;    /* [segs] */
	uint64_t *rbw;            /* [segs] slots read before the segment wrote them */
	uint16_t *out_src;        /* [segs][64] where the value a slot holds after the segment came from: low byte r,g,b, high byte
	                             alpha (QOI_SRC_*: a slot of the GUESSED index[], the incoming colour, or the segment's own
	                             bytes); 0xFFFF = slot not written */
	uint32_t *dep;            /* [segs] bit0: a pixel used the incoming r,g,b; bit1: ... the incoming alpha;
	                             bits 8-15 / 16-23: QOI_SRC_* of the outgoing colour's r,g,b / alpha */
	uint8_t *redo;            /* [segs] decode (again) in the next fdecode launch */
	/* verification */
	uint32_t *verified;       /* [n_images] segments [0, verified) are final */
	uint32_t *ex_px;          /* [n_images] exact colour after the verified prefix */
	uint32_t *ex_tab;         /* [n_images][64] exact index[] after the verified prefix */
};

struct QoiFParams {
	const uint8_t *streams;
	const int32_t *lens;
	uint8_t *out;
	int32_t *status;
	const QoiImgMeta *imgs;
	uint32_t n_images;
	uint32_t n_segs; /* capacity over all images */
	int32_t out_channels;
	QoiFState f;
};

__device__ __forceinline__ uint32_t add_bytes(uint32_t a, uint32_t b) {
	return ((a & 0x7F7F7F7Fu) + (b & 0x7F7F7F7Fu)) ^ ((a ^ b) & 0x80808080u);
}
__device__ __forceinline__ uint32_t sub_bytes(uint32_t a, uint32_t b) {
	return ((a | 0x80808080u) - (b & 0x7F7F7F7Fu)) ^ ((a ^ ~b) & 0x80808080u);
}
__device__ __forceinline__ uint32_t be32(const uint8_t *p) {
	return ((uint32_t)p[0] << 24) | ((uint32_t)p[1] << 16) | ((uint32_t)p[2] << 8) | (uint32_t)p[3];
}

/* which image owns global segment g (binary search in seg_base) */
__device__ __forceinline__ uint32_t seg_image(const QoiFParams &P, uint32_t g) {
	uint32_t lo = 0, hi = P.n_images - 1;
	while (lo < hi) {
		const uint32_t mid = (lo + hi + 1) >> 1;
		if (P.f.seg_base[mid] <= g) lo = mid;
		else hi = mid - 1;
	}
	return lo;
}

/* header against the plan, as qoi_decode_kernel does (qoi.h:513-521 is enforced by the host entry points) */
__device__ __forceinline__ bool stream_ok(const QoiImgMeta &m, const uint8_t *s, int32_t size) {
	if (size < QOI_HEADER_BYTES + QOI_END_BYTES) return false;
	return s[0] == 'q' && s[1] == 'o' && s[2] == 'i' && s[3] == 'f' && be32(s + 4) == m.width && be32(s + 8) == m.height &&
	       s[12] == m.channels && s[13] == m.colorspace;
}

/* op length from its first byte (qoi.h:547,552,566; everything else is one byte) and the pixels it yields (a RUN op
   yields run+1, qoi.h:573-575,541-543) */
__device__ __forceinline__ int32_t op_len(uint32_t tag) { return (tag == 0xFEu) ? 4 : (tag == 0xFFu) ? 5 : ((tag >> 6) == 2u ? 2 : 1); }
__device__ __forceinline__ uint32_t op_px(uint32_t tag) { return ((tag >> 6) == 3u && tag < 0xFEu) ? (tag & 0x3Fu) + 1u : 1u; }

/* ------------------------------------------------------------------ 1a. parse: one thread per segment */
__global__ void __launch_bounds__(QOI_FDEC_THREADS) fparse_kernel(const QoiFParams P) {
	const uint32_t g = blockIdx.x * QOI_FDEC_THREADS + threadIdx.x;
	if (g >= P.n_segs) return;
	const uint32_t img = seg_image(P, g);
	const QoiImgMeta m = P.imgs[img];
	const uint8_t *s = P.streams + m.stream_off;
	const int32_t size = P.lens[img];
	if (!stream_ok(m, s, size)) return;
	const int32_t stop = size - QOI_END_BYTES; /* qoi.h:539 */
	const int32_t begin = QOI_HEADER_BYTES + (int32_t)(g - P.f.seg_base[img]) * QOI_FSEG_BYTES;
	if (begin >= stop) return; /* beyond the stream */
	const int32_t end = begin + QOI_FSEG_BYTES < stop ? begin + QOI_FSEG_BYTES : stop;
	/* Walk 0 is made in full.  Walks 1..4 usually fall onto an op boundary of walk 0 within a few ops and are the same
	   from there on: walk 0 remembers its op starts inside the first 128 bytes (bitmap), the others stop as soon as they
	   land on one and take walk 0's tail.  A walk that stays out of phase (e.g. a stream of two-byte ops entered at
	   odd and even offsets) is simply made in full as well. */
	const int32_t near = begin + 128 < end ? begin + 128 : end;
	uint64_t seen_lo = 0, seen_hi = 0;
	uint32_t total0 = 0;
	int32_t p0 = begin;
	while (p0 < end) {
		if (p0 < near) {
			const uint32_t off = (uint32_t)(p0 - begin);
			if (off < 64u) seen_lo |= 1ull << off;
			else seen_hi |= 1ull << (off - 64u);
		}
		const uint32_t tag = s[p0];
		p0 += op_len(tag);
		total0 += op_px(tag);
	}
	const uint32_t exit0 = (uint32_t)(p0 - end);
	P.f.px_cnt[(size_t)g * 5] = total0;
	P.f.exit_skip[(size_t)g * 8] = (uint8_t)exit0;
#pragma unroll 1
	for (int sk = 1; sk < 5; sk++) {
		int32_t p = begin + sk;
		uint32_t cnt = 0;
		bool merged = false;
		while (p < near) {
			const uint32_t off = (uint32_t)(p - begin);
			if (((off < 64u ? seen_lo : seen_hi) >> (off & 63u)) & 1ull) { merged = true; break; }
			const uint32_t tag = s[p];
			p += op_len(tag);
			cnt += op_px(tag);
		}
		uint32_t ex;
		if (merged) {
			/* pixels walk 0 produced before the meeting point */
			uint32_t before = 0;
			for (int32_t q = begin; q < p;) {
				const uint32_t tag = s[q];
				q += op_len(tag);
				before += op_px(tag);
			}
			cnt += total0 - before;
			ex = exit0;
		} else {
			while (p < end) {
				const uint32_t tag = s[p];
				p += op_len(tag);
				cnt += op_px(tag);
			}
			ex = p >= end ? (uint32_t)(p - end) : 0u; /* an entry skip beyond a very short last segment */
		}
		P.f.px_cnt[(size_t)g * 5 + sk] = cnt;
		P.f.exit_skip[(size_t)g * 8 + sk] = (uint8_t)ex;
	}
}

/* ------------------------------------------------------------------ 1b. chain: one warp per stream */
/* The composition itself is sequential (skip -> exit skip), but it only touches two small words per segment: the
   lanes fetch the maps of 32 segments at a time into shared memory, lane 0 composes them there, the lanes write the
   results back coalesced. */
__global__ void __launch_bounds__(32) fchain_kernel(const QoiFParams P) {
	QOI_DYNAMIC_SMEM(smem); /* (32*5 + 32*2 + 32 + 32 + 4) words */
	uint32_t *s_cnt = (uint32_t *)smem, *s_exit = s_cnt + 32 * 5, *s_entry = s_exit + 32 * 2, *s_px = s_entry + 32;
	uint32_t *s_state = s_px + 32; /* skip, px, live */
	const uint32_t img = blockIdx.x;
	const uint32_t lane = threadIdx.x;
	if (img >= P.n_images) return;
	const QoiImgMeta m = P.imgs[img];
	const uint8_t *s = P.streams + m.stream_off;
	const int32_t size = P.lens[img];
	const bool ok = stream_ok(m, s, size);
	if (lane == 0 && P.status) P.status[img] = ok ? 0 : 4; /* QOI_GPU_E_STREAM */
	if (lane == 0) { s_state[0] = 0; s_state[1] = 0; s_state[2] = 0; }
	__syncwarp();
	if (ok) {
		const int32_t stop = size - QOI_END_BYTES;
		const uint32_t base = P.f.seg_base[img];
		const uint32_t nseg = stop > QOI_HEADER_BYTES ? (uint32_t)((stop - QOI_HEADER_BYTES + QOI_FSEG_BYTES - 1) / QOI_FSEG_BYTES) : 0u;
		for (uint32_t j0 = 0; j0 < nseg; j0 += 32) {
			const uint32_t nb = nseg - j0 < 32u ? nseg - j0 : 32u;
			if (lane < nb) {
				const uint32_t g = base + j0 + lane;
#pragma unroll
				for (int k = 0; k < 5; k++) s_cnt[lane * 5 + k] = P.f.px_cnt[(size_t)g * 5 + k];
				const uint2 e = *(const uint2 *)(P.f.exit_skip + (size_t)g * 8);
				s_exit[lane * 2] = e.x;
				s_exit[lane * 2 + 1] = e.y;
			}
			__syncwarp();
			if (lane == 0) {
				uint32_t skip = s_state[0], px = s_state[1], live = s_state[2];
				for (uint32_t k = 0; k < nb; k++) {
					if (px >= m.npx) { s_entry[k] = 0xFFu; continue; } /* all pixels already produced: dead segment */
					s_entry[k] = skip;
					s_px[k] = px;
					px += s_cnt[k * 5 + skip];
					skip = (skip < 4 ? (s_exit[k * 2] >> (8 * skip)) : s_exit[k * 2 + 1]) & 0xFFu;
					live = j0 + k + 1;
				}
				s_state[0] = skip; s_state[1] = px; s_state[2] = live;
			}
			__syncwarp();
			if (lane < nb && s_entry[lane] != 0xFFu) {
				const uint32_t g = base + j0 + lane;
				P.f.entry[g] = (uint8_t)s_entry[lane];
				P.f.px_start[g] = s_px[lane];
				P.f.redo[g] = 1;
				P.f.in_px[g] = 0xFF000000u; /* round-0 guess: the reference's initial state (qoi.h:533-537) */
			}
			__syncwarp();
			if (s_state[1] >= m.npx) break;
		}
	}
	if (lane == 0) {
		P.f.n_live[img] = s_state[2];
		P.f.verified[img] = 0;
		P.f.ex_px[img] = 0xFF000000u;
	}
	P.f.ex_tab[(size_t)img * 64 + lane] = 0u;
	P.f.ex_tab[(size_t)img * 64 + lane + 32] = 0u;
}

/* round-0 guess of index[]: zeros (qoi.h:533) */
__global__ void __launch_bounds__(QOI_FDEC_THREADS) fzero_kernel(const QoiFParams P) {
	const uint32_t g = blockIdx.x * QOI_FDEC_THREADS + threadIdx.x;
	if (g >= P.n_segs) return;
	const uint32_t img = seg_image(P, g);
	if (g - P.f.seg_base[img] >= P.f.n_live[img]) return;
	uint4 *t = (uint4 *)(P.f.in_tab + (size_t)g * 64);
#pragma unroll
	for (int k = 0; k < 16; k++) t[k] = make_uint4(0u, 0u, 0u, 0u);
}

/* ------------------------------------------------------------------ 2. speculative decode: one lane per segment */
template <int OUT_CH>
__global__ void __launch_bounds__(QOI_FDEC_THREADS) fdecode_kernel(const QoiFParams P) {
	QOI_DYNAMIC_SMEM(smem);
	uint32_t *tab = (uint32_t *)smem + threadIdx.x; /* tab[slot * QOI_FDEC_THREADS] */
	const uint32_t g = blockIdx.x * QOI_FDEC_THREADS + threadIdx.x;
	if (g >= P.n_segs) return;
	const uint32_t img = seg_image(P, g);
	const uint32_t j = g - P.f.seg_base[img];
	const uint32_t live = P.f.n_live[img];
	if (j >= live || j < P.f.verified[img] || !P.f.redo[g]) return;
	const QoiImgMeta m = P.imgs[img];
	const uint8_t *s = P.streams + m.stream_off;
	const int32_t stop = P.lens[img] - QOI_END_BYTES;
	const int32_t begin = QOI_HEADER_BYTES + (int32_t)j * QOI_FSEG_BYTES;
	const int32_t end = begin + QOI_FSEG_BYTES < stop ? begin + QOI_FSEG_BYTES : stop;
	const uint32_t px_begin = P.f.px_start[g];
	/* the last live segment also owns the pixels a truncated stream leaves over (they repeat the last colour, qoi.h:544) */
	const uint32_t px_end = (j + 1 < live) ? P.f.px_start[g + 1] : m.npx;

	uint32_t px = P.f.in_px[g];
	{
		const uint32_t *src = P.f.in_tab + (size_t)g * 64;
#pragma unroll 8
		for (int k = 0; k < 64; k++) tab[k * QOI_FDEC_THREADS] = src[k];
	}
	/* Per slot: where the value it holds now came from (QOI_SRC_*; r,g,b in the low byte, alpha in the high byte,
	   0xFFFF = not written by this segment); two slots to a word, words interleaved over the threads (conflict free).
	   `src` is the same for the current colour. */
	uint16_t *sst = (uint16_t *)((uint32_t *)smem + 64 * QOI_FDEC_THREADS + threadIdx.x); /* sst[(slot >> 1) * 2 * THREADS + (slot & 1)] */
#pragma unroll
	for (int k = 0; k < 32; k++) *(uint32_t *)(sst + k * 2 * QOI_FDEC_THREADS) = 0xFFFFFFFFu;
	uint64_t rbw = 0;
	uint32_t src = QOI_SRC_IN | (QOI_SRC_IN << 8);
	bool dep_rgb = false, dep_a = false;
	int32_t p = begin + (int32_t)P.f.entry[g];
	uint32_t run = 0;
	uint8_t *o = P.out + m.dec_off + (size_t)px_begin * OUT_CH;
	for (uint32_t i = px_begin; i < px_end; i++) {
		if (run > 0) {
			run--;
		} else if (p < end) {
			const uint32_t tag = s[p++];
			if (tag == 0xFEu) { /* qoi.h:547-551: r,g,b replaced, alpha kept */
				px = (px & 0xFF000000u) | (uint32_t)s[p] | ((uint32_t)s[p + 1] << 8) | ((uint32_t)s[p + 2] << 16);
				p += 3;
				src = (src & 0xFF00u) | QOI_SRC_OWN;
			} else if (tag == 0xFFu) { /* qoi.h:552-557 */
				px = (uint32_t)s[p] | ((uint32_t)s[p + 1] << 8) | ((uint32_t)s[p + 2] << 16) | ((uint32_t)s[p + 3] << 24);
				p += 4;
				src = QOI_SRC_OWN | (QOI_SRC_OWN << 8);
			} else if ((tag & 0xC0u) == 0x00u) { /* qoi.h:558-560 */
				const uint32_t st = sst[(tag >> 1) * 2 * QOI_FDEC_THREADS + (tag & 1u)];
				if (st == 0xFFFFu) {
					rbw |= 1ull << tag;
					src = tag | (tag << 8);
				} else {
					src = st;
				}
				px = tab[tag * QOI_FDEC_THREADS];
			} else if ((tag & 0xC0u) == 0x40u) { /* qoi.h:561-565 */
				const uint32_t sel = ((tag >> 4) & 3u) | (((tag >> 2) & 3u) << 4) | ((tag & 3u) << 8) | 0x4000u;
				px = add_bytes(px, __byte_perm(0x0100FFFEu, 0u, sel));
			} else if ((tag & 0xC0u) == 0x80u) { /* qoi.h:566-572 */
				const uint32_t b2 = s[p++];
				const uint32_t vg = (tag & 0x3Fu) + 224u;
				const uint32_t rb = vg + 248u;
				px = add_bytes(px, ((rb + ((b2 >> 4) & 0xFu)) & 0xFFu) | ((vg & 0xFFu) << 8) | (((rb + (b2 & 0xFu)) & 0xFFu) << 16));
			} else { /* qoi.h:573-575 */
				run = tag & 0x3Fu;
			}
			dep_rgb = dep_rgb || (src & 0xFFu) == QOI_SRC_IN;
			dep_a = dep_a || (src >> 8) == QOI_SRC_IN;
			const uint32_t h = qoi_slot(px);
			tab[h * QOI_FDEC_THREADS] = px; /* qoi.h:577 */
			sst[(h >> 1) * 2 * QOI_FDEC_THREADS + (h & 1u)] = (uint16_t)src;
		} else { /* stream exhausted: the pixel repeats the current colour */
			dep_rgb = dep_rgb || (src & 0xFFu) == QOI_SRC_IN;
			dep_a = dep_a || (src >> 8) == QOI_SRC_IN;
		}
		if (OUT_CH == 4) {
			*(uint32_t *)o = px;
		} else {
			o[0] = (uint8_t)px; o[1] = (uint8_t)(px >> 8); o[2] = (uint8_t)(px >> 16);
		}
		o += OUT_CH;
	}
	uint64_t written = 0;
	{
		uint32_t *dsrc = (uint32_t *)(P.f.out_src + (size_t)g * 64);
#pragma unroll 4
		for (int k = 0; k < 32; k++) {
			const uint32_t wv = *(const uint32_t *)(sst + k * 2 * QOI_FDEC_THREADS);
			if ((wv & 0xFFFFu) != 0xFFFFu) written |= 1ull << (2 * k);
			if ((wv >> 16) != 0xFFFFu) written |= 2ull << (2 * k);
			dsrc[k] = wv;
		}
	}
	P.f.out_px[g] = px;
	P.f.out_written[g] = written;
	P.f.rbw[g] = rbw;
	P.f.dep[g] = (dep_rgb ? 1u : 0u) | (dep_a ? 2u : 0u) | (src << 8);
	P.f.redo[g] = 0;
	{
		uint32_t *dst = P.f.out_tab + (size_t)g * 64;
#pragma unroll 8
		for (int k = 0; k < 64; k++) dst[k] = tab[k * QOI_FDEC_THREADS];
	}
}

/* what the in-order walks (fverify_kernel, ffix_kernel) need to know about one segment; loaded a few segments ahead of
   use because none of it depends on the chain state -- the walk itself is then shared-memory / register work only */
struct SegRec {
	uint64_t rbw, wr;
	uint32_t s0, s1; /* QOI_SRC_* pairs of slots lane and lane + 32 */
	uint32_t dep, g_px, o_px, g0, g1, o0, o1, redo;
};
#define QOI_FWALK_BATCH 4
__device__ __forceinline__ void load_rec(const QoiFParams &P, uint32_t g, uint32_t lane, SegRec &r) {
	r.rbw = P.f.rbw[g];
	r.wr = P.f.out_written[g];
	r.s0 = P.f.out_src[(size_t)g * 64 + lane];
	r.s1 = P.f.out_src[(size_t)g * 64 + lane + 32];
	r.dep = P.f.dep[g];
	r.redo = P.f.redo[g];
	r.g_px = P.f.in_px[g];
	r.o_px = P.f.out_px[g];
	r.g0 = P.f.in_tab[(size_t)g * 64 + lane];
	r.g1 = P.f.in_tab[(size_t)g * 64 + lane + 32];
	r.o0 = P.f.out_tab[(size_t)g * 64 + lane];
	r.o1 = P.f.out_tab[(size_t)g * 64 + lane + 32];
}

/* ------------------------------------------------------------------ 3. verify + re-guess: one warp per stream */
__global__ void __launch_bounds__(32) fverify_kernel(const QoiFParams P) {
	const uint32_t img = blockIdx.x;
	const uint32_t lane = threadIdx.x;
	if (img >= P.n_images) return;
	const uint32_t live = P.f.n_live[img];
	uint32_t ver = P.f.verified[img];
	if (ver >= live) return;
	const uint32_t base = P.f.seg_base[img];
	/* chain state: lane holds slots lane and lane+32 */
	uint32_t t0 = P.f.ex_tab[(size_t)img * 64 + lane], t1 = P.f.ex_tab[(size_t)img * 64 + lane + 32];
	uint32_t px = P.f.ex_px[img];
	bool exact = true;
	SegRec recs[QOI_FWALK_BATCH], ahead[QOI_FWALK_BATCH]; /* the batch being walked, the batch in flight */
#pragma unroll
	for (int b = 0; b < QOI_FWALK_BATCH; b++)
		if (ver + b < live) load_rec(P, base + ver + b, lane, recs[b]);
	for (uint32_t j0 = ver; j0 < live; j0 += QOI_FWALK_BATCH) {
#pragma unroll
		for (int b = 0; b < QOI_FWALK_BATCH; b++)
			if (j0 + QOI_FWALK_BATCH + b < live) load_rec(P, base + j0 + QOI_FWALK_BATCH + b, lane, ahead[b]);
#pragma unroll
		for (int b = 0; b < QOI_FWALK_BATCH; b++) {
			const uint32_t j = j0 + b;
			if (j >= live) break;
			const uint32_t g = base + j;
			const SegRec &r = recs[b];
			const uint32_t dep = r.dep, g_px = r.g_px, o_px = r.o_px;
			/* does the guess this segment was decoded with agree with the chain state on everything it used? */
			const uint32_t pmask = ((dep & 1u) ? 0x00FFFFFFu : 0u) | ((dep & 2u) ? 0xFF000000u : 0u);
			/* which slots of the guess equal the chain state */
			const uint64_t agree = (uint64_t)__ballot_sync(QOI_FULL_MASK, r.g0 == t0) | ((uint64_t)__ballot_sync(QOI_FULL_MASK, r.g1 == t1) << 32);
			const bool differs = (r.rbw & ~agree) != 0ull || ((g_px ^ px) & pmask) != 0u;
			uint64_t apply = r.wr;
			if (differs) {
				/* new guess for this segment = the chain state (the exact state while `exact` still holds) */
				P.f.in_tab[(size_t)g * 64 + lane] = t0;
				P.f.in_tab[(size_t)g * 64 + lane + 32] = t1;
				if (lane == 0) {
					P.f.in_px[g] = px;
					P.f.redo[g] = 1;
				}
				if (exact) {
					exact = false;
					ver = j; /* everything before j is final */
				}
				/* Keep chaining for the segments behind it.  Of its stale writes only those are trusted whose source was
				   guessed right (a slot that agrees, the incoming colour if it agrees, or the segment's own bytes); the other
				   slots keep the older value (right when the table is stable, e.g. a palette).  The outgoing colour: a
				   channel whose source was wrong is re-based on the chain's value of that source (r,g,b are source + deltas,
				   alpha is the source's alpha). */
				const bool in_rgb_ok = ((g_px ^ px) & 0x00FFFFFFu) == 0u, in_a_ok = ((g_px ^ px) & 0xFF000000u) == 0u;
				const uint32_t lo = 1u << lane;
				uint64_t ap = 0;
				{
					const uint32_t c0 = r.s0 & 0xFFu, a0 = r.s0 >> 8, c1 = r.s1 & 0xFFu, a1 = r.s1 >> 8;
					const bool ok0 = r.s0 != 0xFFFFu && (c0 == QOI_SRC_OWN || (c0 == QOI_SRC_IN ? in_rgb_ok : ((agree >> c0) & 1ull) != 0ull)) &&
					                 (a0 == QOI_SRC_OWN || (a0 == QOI_SRC_IN ? in_a_ok : ((agree >> a0) & 1ull) != 0ull));
					const bool ok1 = r.s1 != 0xFFFFu && (c1 == QOI_SRC_OWN || (c1 == QOI_SRC_IN ? in_rgb_ok : ((agree >> c1) & 1ull) != 0ull)) &&
					                 (a1 == QOI_SRC_OWN || (a1 == QOI_SRC_IN ? in_a_ok : ((agree >> a1) & 1ull) != 0ull));
					ap = (ok0 ? (uint64_t)lo : 0ull) | (ok1 ? (uint64_t)lo << 32 : 0ull);
				}
				apply = ap;
				const uint32_t sc = (dep >> 8) & 0xFFu, sa = (dep >> 16) & 0xFFu;
				/* the guessed and the chain value of the source slots (held by lane slot & 31) */
				const uint32_t gc0 = __shfl_sync(QOI_FULL_MASK, r.g0, sc & 31u), gc1 = __shfl_sync(QOI_FULL_MASK, r.g1, sc & 31u);
				const uint32_t tc0 = __shfl_sync(QOI_FULL_MASK, t0, sc & 31u), tc1 = __shfl_sync(QOI_FULL_MASK, t1, sc & 31u);
				const uint32_t ta0 = __shfl_sync(QOI_FULL_MASK, t0, sa & 31u), ta1 = __shfl_sync(QOI_FULL_MASK, t1, sa & 31u);
				uint32_t rgb = o_px, al = o_px;
				if (sc == QOI_SRC_IN) rgb = add_bytes(px, sub_bytes(o_px, g_px));
				else if (sc < 64u) rgb = add_bytes((sc & 32u) ? tc1 : tc0, sub_bytes(o_px, (sc & 32u) ? gc1 : gc0));
				if (sa == QOI_SRC_IN) al = px;
				else if (sa < 64u) al = (sa & 32u) ? ta1 : ta0;
				px = (rgb & 0x00FFFFFFu) | (al & 0xFF000000u);
			} else {
				if (exact) ver = j + 1;
				px = o_px; /* agreed on everything it used: the decode is what the chain state produces */
			}
			if ((apply >> lane) & 1ull) t0 = r.o0;
			if ((apply >> (lane + 32)) & 1ull) t1 = r.o1;
			if (exact) { /* remember the exact state after the verified prefix (the fix-up and the next round start here) */
				P.f.ex_tab[(size_t)img * 64 + lane] = t0;
				P.f.ex_tab[(size_t)img * 64 + lane + 32] = t1;
				if (lane == 0) P.f.ex_px[img] = px;
			}
		}
#pragma unroll
		for (int b = 0; b < QOI_FWALK_BATCH; b++) recs[b] = ahead[b];
	}
	if (lane == 0) P.f.verified[img] = ver;
}

/* ------------------------------------------------------------------ 4. exact fix-up: one warp per stream */
/* Walks the not yet verified segments in order with the exact state, like fverify_kernel, but a segment whose guess
   does not hold is decoded right here (lane 0, the reference loop qoi.h:540-587) so the state stays exact.  Valid
   segments cost a 64-entry compare and merge; the sequential work is proportional to the segments that speculation
   could not settle.  After this kernel every pixel is final. */
template <int OUT_CH>
__global__ void __launch_bounds__(32) ffix_kernel(const QoiFParams P) {
	QOI_DYNAMIC_SMEM(smem);
	uint32_t *tab = (uint32_t *)smem; /* exact index[] */
	const uint32_t img = blockIdx.x;
	const uint32_t lane = threadIdx.x;
	if (img >= P.n_images) return;
	const uint32_t live = P.f.n_live[img];
	const uint32_t ver = P.f.verified[img];
	if (ver >= live) return; /* everything verified (also: rejected or empty stream) */
	const QoiImgMeta m = P.imgs[img];
	const uint8_t *s = P.streams + m.stream_off;
	const int32_t stop = P.lens[img] - QOI_END_BYTES;
	const uint32_t base = P.f.seg_base[img];
	tab[lane] = P.f.ex_tab[(size_t)img * 64 + lane];
	tab[lane + 32] = P.f.ex_tab[(size_t)img * 64 + lane + 32];
	uint32_t px = P.f.ex_px[img];
	__syncwarp();
	SegRec recs[QOI_FWALK_BATCH], ahead[QOI_FWALK_BATCH];
#pragma unroll
	for (int b = 0; b < QOI_FWALK_BATCH; b++)
		if (ver + b < live) load_rec(P, base + ver + b, lane, recs[b]);
	for (uint32_t j0 = ver; j0 < live; j0 += QOI_FWALK_BATCH) {
#pragma unroll
	for (int b = 0; b < QOI_FWALK_BATCH; b++)
		if (j0 + QOI_FWALK_BATCH + b < live) load_rec(P, base + j0 + QOI_FWALK_BATCH + b, lane, ahead[b]);
#pragma unroll
	for (int b = 0; b < QOI_FWALK_BATCH; b++) {
		const uint32_t j = j0 + b;
		if (j >= live) break;
		const uint32_t g = base + j;
		const SegRec &r = recs[b];
		const uint32_t dep = r.dep, g_px = r.g_px;
		const uint32_t pmask = ((dep & 1u) ? 0x00FFFFFFu : 0u) | ((dep & 2u) ? 0xFF000000u : 0u);
		bool bad = ((r.rbw >> lane) & 1ull) && r.g0 != tab[lane];
		bad = bad || (((r.rbw >> (lane + 32)) & 1ull) && r.g1 != tab[lane + 32]);
		bad = bad || (lane == 0 && ((g_px ^ px) & pmask) != 0u);
		bad = bad || r.redo != 0; /* a new guess was issued but never decoded */
		if (__ballot_sync(QOI_FULL_MASK, bad) == 0u) {
			px = r.o_px;
			if ((r.wr >> lane) & 1ull) tab[lane] = r.o0;
			if ((r.wr >> (lane + 32)) & 1ull) tab[lane + 32] = r.o1;
		} else if (lane == 0) {
			const int32_t begin = QOI_HEADER_BYTES + (int32_t)j * QOI_FSEG_BYTES;
			const int32_t end = begin + QOI_FSEG_BYTES < stop ? begin + QOI_FSEG_BYTES : stop;
			const uint32_t px_begin = P.f.px_start[g];
			const uint32_t px_end = (j + 1 < live) ? P.f.px_start[g + 1] : m.npx;
			int32_t p = begin + (int32_t)P.f.entry[g];
			uint32_t run = 0;
			uint8_t *o = P.out + m.dec_off + (size_t)px_begin * OUT_CH;
			for (uint32_t i = px_begin; i < px_end; i++) {
				if (run > 0) {
					run--;
				} else if (p < end) {
					const uint32_t tag = s[p++];
					if (tag == 0xFEu) {
						px = (px & 0xFF000000u) | (uint32_t)s[p] | ((uint32_t)s[p + 1] << 8) | ((uint32_t)s[p + 2] << 16);
						p += 3;
					} else if (tag == 0xFFu) {
						px = (uint32_t)s[p] | ((uint32_t)s[p + 1] << 8) | ((uint32_t)s[p + 2] << 16) | ((uint32_t)s[p + 3] << 24);
						p += 4;
					} else if ((tag & 0xC0u) == 0x00u) {
						px = tab[tag];
					} else if ((tag & 0xC0u) == 0x40u) {
						const uint32_t sel = ((tag >> 4) & 3u) | (((tag >> 2) & 3u) << 4) | ((tag & 3u) << 8) | 0x4000u;
						px = add_bytes(px, __byte_perm(0x0100FFFEu, 0u, sel));
					} else if ((tag & 0xC0u) == 0x80u) {
						const uint32_t b2 = s[p++];
						const uint32_t vg = (tag & 0x3Fu) + 224u;
						const uint32_t rb = vg + 248u;
						px = add_bytes(px, ((rb + ((b2 >> 4) & 0xFu)) & 0xFFu) | ((vg & 0xFFu) << 8) | (((rb + (b2 & 0xFu)) & 0xFFu) << 16));
					} else {
						run = tag & 0x3Fu;
					}
					tab[qoi_slot(px)] = px;
				}
				if (OUT_CH == 4) {
					*(uint32_t *)o = px;
				} else {
					o[0] = (uint8_t)px; o[1] = (uint8_t)(px >> 8); o[2] = (uint8_t)(px >> 16);
				}
				o += OUT_CH;
			}
		}
		__syncwarp();
		px = __shfl_sync(QOI_FULL_MASK, px, 0);
	}
#pragma unroll
	for (int b = 0; b < QOI_FWALK_BATCH; b++) recs[b] = ahead[b];
	}
	if (lane == 0) P.f.verified[img] = live;
}

/* binds the arrays of a QoiFState onto one allocation laid out by qoi_fstate_layout() (qoi_plan.h) */
static inline void fstate_bind(QoiFState &f, unsigned char *base, const size_t *off) {
	f.seg_base = (const uint32_t *)(base + off[0]);
	f.px_cnt = (uint32_t *)(base + off[1]);
	f.exit_skip = (uint8_t *)(base + off[2]);
	f.entry = (uint8_t *)(base + off[3]);
	f.px_start = (uint32_t *)(base + off[4]);
	f.n_live = (uint32_t *)(base + off[5]);
	f.in_px = (uint32_t *)(base + off[6]);
	f.in_tab = (uint32_t *)(base + off[7]);
	f.out_px = (uint32_t *)(base + off[8]);
	f.out_tab = (uint32_t *)(base + off[9]);
	f.out_written = (uint64_t *)(base + off[10]);
	f.rbw = (uint64_t *)(base + off[11]);
	f.out_src = (uint16_t *)(base + off[17]);
	f.dep = (uint32_t *)(base + off[12]);
	f.redo = (uint8_t *)(base + off[13]);
	f.verified = (uint32_t *)(base + off[14]);
	f.ex_px = (uint32_t *)(base + off[15]);
	f.ex_tab = (uint32_t *)(base + off[16]);
}

}  // namespace qoifdec
