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
#include "qoi_dec.cuh" /* the chunked stream reader (qoidec::Reader) */

namespace qoifdec {
using namespace qoidev;

#define QOI_FSEG_BYTES 4096
#define QOI_FDEC_ROUNDS 10 /* speculative decode launches (with a verify pass between them) before the exact fix-up */
#define QOI_FDEC_THREADS 128
#define QOI_FFIX_SMEM ((64 + QOI_FSEG_BYTES / 4 + 8) * 4) /* index[], one staged segment */
#define QOI_FCHAIN_SMEM ((32 * 5 + 32 * 2 + 32 + 32 + 4) * 4)
#define QOI_FDEC_SMEM ((64 + 32 + QOI_DEC_RING_WORDS) * QOI_FDEC_THREADS * 4) /* per lane: index[64], 64 source half-words, stream ring */
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
	uint32_t *redo_list;      /* [segs] the segments to decode in the next fdecode launch (rounds >= 1), in no particular order */
	uint32_t *n_redo;         /* [64] length of that list per round */
	/* double buffered by pass parity: a launch reads [fsel_rd], the verify pass writes [fsel_wr] */
	uint32_t *verified;       /* [2][n_images] segments [0, verified) are final */
	uint32_t *ex_px;          /* [2][n_images] exact colour after the verified prefix */
	uint32_t *ex_tab;         /* [2][n_images][64] exact index[] after the verified prefix */
	uint32_t *bstate;         /* [2][n_images][QOI_FVER_CHUNKS][65] chain state at the start of each verify range */
	uint32_t *vcount;         /* [3][n_images] per verify pass (mod 3): rejected guesses + ranges whose start state was not the
	                             state their predecessor ended with; 0 = the pass found the whole stream consistent */
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
	uint32_t round; /* speculative round of this launch (0 = every live segment, from the reference's initial state) */
	QoiFState f;
};

#define QOI_FVER_CHUNKS 8  /* verify warps (ranges) per stream */
#define QOI_FVER_BWORDS 65 /* index[64] + colour */
#ifdef QOI_EMU /* tests/emu: small streams must be able to exercise the multi-range path */
static inline uint32_t qoi_fver_min_range() { const char *e = getenv("QOI_EMU_FVER_MIN_RANGE"); return e ? (uint32_t)atoi(e) : 1024u; }
#define QOI_FVER_MIN_RANGE qoi_fver_min_range()
#else
#define QOI_FVER_MIN_RANGE 1024u /* segments per range at least */
#endif
/* which copy of the pass-buffered state (verified, ex_px, ex_tab, bstate) a launch reads / the verify pass writes */
__device__ __forceinline__ uint32_t fsel_rd(const QoiFParams &P) { return (P.round + 1u) & 1u; }
__device__ __forceinline__ uint32_t fsel_wr(const QoiFParams &P) { return P.round & 1u; }

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
		P.f.verified[img] = P.f.verified[P.n_images + img] = 0; /* both parities of the pass-buffered state */
		P.f.vcount[img] = P.f.vcount[P.n_images + img] = P.f.vcount[2u * P.n_images + img] = 0;
		P.f.ex_px[img] = P.f.ex_px[P.n_images + img] = 0xFF000000u;
	}
	P.f.ex_tab[(size_t)img * 64 + lane] = 0u;
	P.f.ex_tab[(size_t)img * 64 + lane + 32] = 0u;
	P.f.ex_tab[((size_t)P.n_images + img) * 64 + lane] = 0u;
	P.f.ex_tab[((size_t)P.n_images + img) * 64 + lane + 32] = 0u;
}

/* round-0 guess of index[]: zeros (qoi.h:533) */
__global__ void __launch_bounds__(QOI_FDEC_THREADS) fzero_kernel(const QoiFParams P) {
	const uint32_t g = blockIdx.x * QOI_FDEC_THREADS + threadIdx.x;
	if (g < 64u) P.f.n_redo[g] = 0u;
	if (g >= P.n_segs) return;
	const uint32_t img = seg_image(P, g);
	if (g - P.f.seg_base[img] >= P.f.n_live[img]) return;
	uint4 *t = (uint4 *)(P.f.in_tab + (size_t)g * 64);
#pragma unroll
	for (int k = 0; k < 16; k++) t[k] = make_uint4(0u, 0u, 0u, 0u);
}

/* ------------------------------------------------------------------ 2. speculative decode: one lane per segment */
template <int OUT_CH, bool UNI>
__global__ void __launch_bounds__(QOI_FDEC_THREADS) fdecode_kernel(const QoiFParams P) {
	QOI_DYNAMIC_SMEM(smem);
	uint32_t *tab = (uint32_t *)smem + threadIdx.x; /* tab[slot * QOI_FDEC_THREADS] */
	/* round 0: lane k takes segment k; later rounds: the k-th entry of the list the last verify pass left, so that the
	   few segments that are decoded again fill whole warps */
	uint32_t g = blockIdx.x * QOI_FDEC_THREADS + threadIdx.x;
	if (P.round > 0u) {
		if (g >= P.f.n_redo[P.round]) return;
		g = P.f.redo_list[g];
	}
	if (g >= P.n_segs) return;
	const uint32_t img = seg_image(P, g);
	const uint32_t j = g - P.f.seg_base[img];
	const uint32_t live = P.f.n_live[img];
	if (j >= live || j < P.f.verified[fsel_rd(P) * P.n_images + img] || !P.f.redo[g]) return;
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
	constexpr int B = QOI_FDEC_THREADS;
	uint16_t *sst = (uint16_t *)((uint32_t *)smem + 64 * B + threadIdx.x); /* sst[(slot >> 1) * 2 * B + (slot & 1)] */
	uint32_t *ring = (uint32_t *)smem + (64 + 32) * B + threadIdx.x;        /* ring[word * B] */
#pragma unroll
	for (int k = 0; k < 32; k++) *(uint32_t *)(sst + k * 2 * B) = 0xFFFFFFFFu;
	uint64_t rbw = 0;
	uint32_t src = QOI_SRC_IN | (QOI_SRC_IN << 8);
	bool dep_rgb = false, dep_a = false;
	const int32_t p0 = begin + (int32_t)P.f.entry[g];
	uint32_t run = 0;

	/* stream bytes: the reader of the indexed decoder (qoi_dec.cuh) -- 32-byte chunks one ahead, a per-lane ring in
	   shared memory, a 64-bit window; an op that straddles `end` is read to its last byte (it lies inside the stream:
	   the end marker follows `stop`), words outside the stream read as zero and are never touched */
	qoidec::Reader rd;
	{
		const uintptr_t a = (uintptr_t)(s + p0);
		rd.lo = (uintptr_t)s & ~(uintptr_t)3;
		rd.hi = ((uintptr_t)s + (uintptr_t)P.lens[img] + 3u) & ~(uintptr_t)3;
		rd.gp = (const uint4 *)(a & ~(uintptr_t)31);
		rd.p = p0;
#pragma unroll
		for (int half = 0; half < 2; half++) {
			qoidec::load_chunk(rd);
			uint32_t *r = ring + half * 8 * B;
			r[0 * B] = rd.pa.x; r[1 * B] = rd.pa.y; r[2 * B] = rd.pa.z; r[3 * B] = rd.pa.w;
			r[4 * B] = rd.pb.x; r[5 * B] = rd.pb.y; r[6 * B] = rd.pb.z; r[7 * B] = rd.pb.w;
		}
		qoidec::load_chunk(rd);
		rd.wi = (uint32_t)((a & 31u) >> 2);
		rd.buf = 0;
		rd.avail = 0;
	}
	auto next_word = [&]() -> uint32_t {
		const uint32_t w = ring[(rd.wi & (QOI_DEC_RING_WORDS - 1)) * B];
		rd.wi++;
		if (!UNI && (rd.wi & 7u) == 0u) { /* lane-driven refill: the half just read is free */
			uint32_t *r = ring + ((rd.wi - 8u) & (QOI_DEC_RING_WORDS - 1)) * B;
			r[0 * B] = rd.pa.x; r[1 * B] = rd.pa.y; r[2 * B] = rd.pa.z; r[3 * B] = rd.pa.w;
			r[4 * B] = rd.pb.x; r[5 * B] = rd.pb.y; r[6 * B] = rd.pb.z; r[7 * B] = rd.pb.w;
			qoidec::load_chunk(rd);
		}
		return w;
	};
	/* Warp-uniform refill (UNI): called once per loop trip (at most 4 pixels = 20 stream bytes) by all lanes together.
	   Chunk `filled` is parked in registers; it goes into ring half (filled & 1) as soon as the reader has left chunk
	   filled - 2, and the next chunk is put in flight.  Doing this at the same instruction for all lanes matters: the
	   scoreboard tracks the parked registers per warp, so with lane-driven refills every lane's refill waited for the
	   load another lane had just issued (37 % of the decoder's stall samples in round 1). */
	uint32_t filled = 2;
	auto refill = [&]() {
		if (UNI && (rd.wi >> 3) + 1u >= filled) {
			uint32_t *r = ring + ((filled & 1u) * 8u) * B;
			r[0 * B] = rd.pa.x; r[1 * B] = rd.pa.y; r[2 * B] = rd.pa.z; r[3 * B] = rd.pa.w;
			r[4 * B] = rd.pb.x; r[5 * B] = rd.pb.y; r[6 * B] = rd.pb.z; r[7 * B] = rd.pb.w;
			filled++;
			qoidec::load_chunk(rd);
		}
	};
	{
		const uint32_t mis = (uint32_t)((uintptr_t)(s + p0) & 3u);
		rd.buf = (uint64_t)(next_word() >> (8u * mis));
		rd.avail = 4u - mis;
	}

	/* one pixel (qoi.h:541-578), evaluated branch-free like qoi_decode_kernel, plus the bookkeeping of where the
	   colour came from */
	auto step = [&]() -> uint32_t {
		const bool parse = (run == 0u) && (rd.p < end);
		if (parse && rd.avail < 5u) {
			rd.buf |= (uint64_t)next_word() << (8u * rd.avail);
			rd.avail += 4u;
			if (rd.avail < 5u) {
				rd.buf |= (uint64_t)next_word() << (8u * rd.avail);
				rd.avail += 4u;
			}
		}
		const uint32_t tag = (uint32_t)rd.buf & 0xFFu;
		const uint32_t pay = (uint32_t)(rd.buf >> 8);
		const uint32_t top = tag >> 6, t6 = tag & 63u;
		const bool is_rgb = tag == 0xFEu, is_rgba = tag == 0xFFu;               /* qoi.h:547-557 */
		const uint32_t idxv = tab[t6 * B];                                       /* qoi.h:558-560 */
		const uint32_t ist = sst[(t6 >> 1) * 2 * B + (t6 & 1u)];                 /* ... and where that slot's value came from */
		const uint32_t sel = ((tag >> 4) & 3u) | (((tag >> 2) & 3u) << 4) | ((tag & 3u) << 8) | 0x4000u;
		const uint32_t dadd = __byte_perm(0x0100FFFEu, 0u, sel);                 /* qoi.h:561-565 */
		const uint32_t vg = t6 + 224u;
		const uint32_t rb = vg + 248u;
		const uint32_t ladd = ((rb + ((pay >> 4) & 0xFu)) & 0xFFu) | ((vg & 0xFFu) << 8) | (((rb + (pay & 0xFu)) & 0xFFu) << 16); /* qoi.h:566-572 */
		const uint32_t rel = add_bytes(px, top == 1u ? dadd : ladd);
		uint32_t npx = top == 0u ? idxv : rel;
		npx = top == 3u ? px : npx;                                              /* qoi.h:573-575 */
		npx = is_rgb ? ((px & 0xFF000000u) | (pay & 0x00FFFFFFu)) : npx;
		npx = is_rgba ? pay : npx;
		const bool fresh = top == 0u && ist == 0xFFFFu; /* INDEX of a slot this segment has not written: the guess is used */
		uint32_t nsrc = top == 0u ? (fresh ? (t6 | (t6 << 8)) : ist) : src;
		nsrc = is_rgb ? ((src & 0xFF00u) | QOI_SRC_OWN) : nsrc; /* r,g,b replaced, alpha kept */
		nsrc = is_rgba ? (QOI_SRC_OWN | (QOI_SRC_OWN << 8)) : nsrc;
		const uint32_t len = is_rgba ? 5u : (is_rgb ? 4u : (top == 2u ? 2u : 1u));
		const uint32_t nrun = (top == 3u && tag < 0xFEu) ? t6 : 0u;
		if (parse) {
			px = npx;
			src = nsrc;
			if (fresh) rbw |= 1ull << t6;
			rd.buf >>= 8u * len;
			rd.avail -= len;
			rd.p += (int32_t)len;
			const uint32_t h = qoi_slot(px);
			tab[h * B] = px; /* qoi.h:577 */
			sst[(h >> 1) * 2 * B + (h & 1u)] = (uint16_t)src;
			run = nrun;
		} else if (run > 0u) {
			run--;
		}
		/* every pixel that leaves while a channel still is the incoming one makes the segment depend on it (also the
		   pixels a truncated stream leaves over, qoi.h:544) */
		dep_rgb = dep_rgb || (src & 0xFFu) == QOI_SRC_IN;
		dep_a = dep_a || (src >> 8) == QOI_SRC_IN;
		return px;
	};

	uint8_t *obase = P.out + m.dec_off;
	uint32_t i = px_begin;
	if (OUT_CH == 4) {
		uint32_t *o32 = (uint32_t *)obase;
		const bool vec_ok = ((uintptr_t)obase & 15u) == 0;
		while (i < px_end && (!vec_ok || (i & 3u))) {
			o32[i] = step();
			i++;
			refill();
		}
		while (i + 4 <= px_end) {
			const uint32_t a = step(), b = step(), c = step(), d = step();
			*(uint4 *)(o32 + i) = make_uint4(a, b, c, d);
			i += 4;
			refill();
		}
		while (i < px_end) {
			o32[i] = step();
			i++;
			refill();
		}
	} else {
		while (i < px_end) {
			const uint32_t a = step();
			uint8_t *o = obase + (size_t)i * 3;
			o[0] = (uint8_t)a; o[1] = (uint8_t)(a >> 8); o[2] = (uint8_t)(a >> 16);
			i++;
			refill();
		}
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

/* What the in-order walks (fverify_kernel, ffix_kernel) need to know about one segment.  None of it depends on the
   chain state, so the records are streamed QOI_FWALK_RING segments ahead of the walk into a ring in shared memory
   with per-thread asynchronous copies; the walk itself is then shared-memory / register work only. */
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

/* Calls body(j, record of segment j) for j = first .. last-1 in order.  Two register batches: while one is walked the
   other is in flight (no copies between them: a register move would wait for its load).  Measured alternatives
   (round 1, B200): records packed to six words per lane and eight segments per batch, and a 16-entry cp.async ring in
   shared memory, were 3-5x SLOWER than this plain form -- see DESIGN.md section 7. */
template <typename Body>
__device__ __forceinline__ void walk_records(const QoiFParams &P, uint32_t base, uint32_t first, uint32_t last, uint32_t lane, Body body) {
	constexpr uint32_t B = QOI_FWALK_BATCH;
	SegRec ra[B], rb[B];
#pragma unroll
	for (uint32_t b = 0; b < B; b++)
		if (first + b < last) load_rec(P, base + first + b, lane, ra[b]);
	uint32_t j0 = first;
	while (j0 < last) {
#pragma unroll
		for (uint32_t b = 0; b < B; b++)
			if (j0 + B + b < last) load_rec(P, base + j0 + B + b, lane, rb[b]);
#pragma unroll
		for (uint32_t b = 0; b < B; b++)
			if (j0 + b < last) body(j0 + b, ra[b]);
		j0 += B;
		if (j0 >= last) break;
#pragma unroll
		for (uint32_t b = 0; b < B; b++)
			if (j0 + B + b < last) load_rec(P, base + j0 + B + b, lane, ra[b]);
#pragma unroll
		for (uint32_t b = 0; b < B; b++)
			if (j0 + b < last) body(j0 + b, rb[b]);
		j0 += B;
	}
}

/* ------------------------------------------------------------------ 3. verify + re-guess: QOI_FVER_CHUNKS warps per stream */
/* The live segments of a stream are cut into QOI_FVER_CHUNKS equal ranges, one warp each.  The warp whose range holds
   the first unverified segment starts there with the EXACT state and is the only one that can verify; the warps of
   the ranges behind it start at their range with the chain state the previous pass had at that point (bstate) and only
   produce guesses.  A range start is therefore one pass stale -- just another guess error; exactness never depends on
   it (the prefix logic and the fix-up do not use it).  All cross-pass state is double buffered by pass parity (a pass
   reads what the pass before wrote), so no warp reads something another warp of the same launch writes. */

__global__ void __launch_bounds__(32) fverify_kernel(const QoiFParams P) {
	const uint32_t img = blockIdx.x / QOI_FVER_CHUNKS, chunk = blockIdx.x % QOI_FVER_CHUNKS;
	const uint32_t lane = threadIdx.x;
	if (img >= P.n_images) return;
	const uint32_t rd = fsel_rd(P) * P.n_images + img, wr = fsel_wr(P) * P.n_images + img;
	const uint32_t live = P.f.n_live[img];
	uint32_t ver = P.f.verified[rd];
	/* ranges only where the walk is long: content whose index[] is stable (UI graphics: few, long segments) converges
	   best with one in-order walk, and there it is short anyway */
	const uint32_t nch = live / QOI_FVER_MIN_RANGE < 1u ? 1u : (live / QOI_FVER_MIN_RANGE > QOI_FVER_CHUNKS ? QOI_FVER_CHUNKS : live / QOI_FVER_MIN_RANGE);
	if (chunk >= nch) return;
	const uint32_t lo = (uint32_t)(((uint64_t)live * chunk) / nch), hi = (uint32_t)(((uint64_t)live * (chunk + 1u)) / nch);
	const bool owner = ver < live ? (ver >= lo && ver < hi) : (chunk == 0u); /* of the exact state */
	if (owner && ver >= live) { /* nothing left: carry the state over to this pass's buffer */
		P.f.ex_tab[(size_t)wr * 64 + lane] = P.f.ex_tab[(size_t)rd * 64 + lane];
		P.f.ex_tab[(size_t)wr * 64 + lane + 32] = P.f.ex_tab[(size_t)rd * 64 + lane + 32];
		if (lane == 0) { P.f.ex_px[wr] = P.f.ex_px[rd]; P.f.verified[wr] = ver; }
		return;
	}
	/* Fixed point: the previous pass rejected no guess and every range started with the state its predecessor ended
	   with, so all its walks together were one exact walk from the verified prefix to the end: everything is final.
	   (The counter was completed by the previous launch; nothing read here is written by this one.) */
	if (P.round >= 1u && ver < live && P.f.vcount[((P.round + 2u) % 3u) * P.n_images + img] == 0u) {
		if (chunk == 0u && lane == 0) P.f.verified[wr] = live;
		return;
	}
	if (chunk == 0u && lane == 0) P.f.vcount[((P.round + 1u) % 3u) * P.n_images + img] = 0u; /* the next pass's counter */
	if (ver >= live || hi <= ver || lo >= hi) return; /* verified range, or empty range */
	const uint32_t base = P.f.seg_base[img];
	uint32_t nbad = 0;
	/* chain state: lane holds slots lane and lane+32 */
	uint32_t t0, t1, px;
	bool exact = owner;
	uint32_t first;
	if (owner) {
		first = ver;
		t0 = P.f.ex_tab[(size_t)rd * 64 + lane];
		t1 = P.f.ex_tab[(size_t)rd * 64 + lane + 32];
		px = P.f.ex_px[rd];
		P.f.ex_tab[(size_t)wr * 64 + lane] = t0; /* stays the exact state if nothing more can be verified */
		P.f.ex_tab[(size_t)wr * 64 + lane + 32] = t1;
		if (lane == 0) P.f.ex_px[wr] = px;
	} else {
		first = lo;
		if (P.round == 0u) {
			/* No earlier pass.  The state must not AGREE with the round-0 guess (zeros, opaque black) by accident, or the
			   garbage a segment decoded from that guess would be trusted -- e.g. alpha 0 from an INDEX of an empty slot,
			   which nothing in an opaque image ever replaces.  0xFFFFFFFF cannot be the content of any slot but 38 (a
			   value sits in the slot it hashes to); opaque alpha is the prior that survives. */
			t0 = t1 = 0xFFFFFFFFu;
			px = 0xFF7F7F7Fu;
		} else {
			const uint32_t *bs = P.f.bstate + ((size_t)rd * QOI_FVER_CHUNKS + chunk) * QOI_FVER_BWORDS;
			t0 = bs[lane];
			t1 = bs[lane + 32];
			px = bs[64];
		}
	}
	{
		auto body = [&](uint32_t j, const SegRec &r) {
			const uint32_t g = base + j;
			const uint32_t dep = r.dep, g_px = r.g_px, o_px = r.o_px;
			/* does the guess this segment was decoded with agree with the chain state on everything it used? */
			const uint32_t pmask = ((dep & 1u) ? 0x00FFFFFFu : 0u) | ((dep & 2u) ? 0xFF000000u : 0u);
			/* which slots of the guess equal the chain state */
			const uint64_t agree = (uint64_t)__ballot_sync(QOI_FULL_MASK, r.g0 == t0) | ((uint64_t)__ballot_sync(QOI_FULL_MASK, r.g1 == t1) << 32);
			/* (a guess that was issued but never decoded -- redo still set -- is never accepted: its record is the old decode's) */
			const bool differs = (r.rbw & ~agree) != 0ull || ((g_px ^ px) & pmask) != 0u || r.redo != 0u;
#ifdef QOI_EMU_FDEBUG /* tests/emu only: why a segment's guess was rejected */
			if (lane == 0 && getenv("QOI_EMU_FDEBUG"))
				fprintf(stderr, "  pass %u chunk %u seg %u %s differs=%d px %08x guess %08x pmask %08x badslots %016llx\n", P.round, chunk, j,
				        exact ? "EXACT" : "guess", (int)differs, px, g_px, pmask, (unsigned long long)(r.rbw & ~agree));
#endif
			uint64_t apply = r.wr;
			nbad += differs ? 1u : 0u;
			if (differs) {
				/* new guess for this segment = the chain state (the exact state while `exact` still holds) */
				P.f.in_tab[(size_t)g * 64 + lane] = t0;
				P.f.in_tab[(size_t)g * 64 + lane + 32] = t1;
				if (lane == 0) {
					P.f.in_px[g] = px;
					P.f.redo[g] = 1;
					P.f.redo_list[atomicAdd(&P.f.n_redo[P.round + 1u], 1u)] = g;
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
				P.f.ex_tab[(size_t)wr * 64 + lane] = t0;
				P.f.ex_tab[(size_t)wr * 64 + lane + 32] = t1;
				if (lane == 0) P.f.ex_px[wr] = px;
			}
		};
		walk_records(P, base, first, hi, lane, body);
	}
	if (owner && lane == 0) P.f.verified[wr] = ver;
	if (chunk + 1u < nch) {
		/* did the next range start (in this pass) with the state this one ends with? */
		const uint32_t *bo = P.f.bstate + ((size_t)rd * QOI_FVER_CHUNKS + chunk + 1u) * QOI_FVER_BWORDS;
		const bool same = P.round >= 1u && bo[lane] == t0 && bo[lane + 32] == t1 && bo[64] == px;
		if (__ballot_sync(QOI_FULL_MASK, same) != QOI_FULL_MASK) nbad++;
		/* where the next range starts in the next pass */
		uint32_t *bs = P.f.bstate + ((size_t)wr * QOI_FVER_CHUNKS + chunk + 1u) * QOI_FVER_BWORDS;
		bs[lane] = t0;
		bs[lane + 32] = t1;
		if (lane == 0) bs[64] = px;
	}
	if (lane == 0 && nbad) atomicAdd(&P.f.vcount[(P.round % 3u) * P.n_images + img], nbad);
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
	uint32_t *stage = tab + 64;       /* one segment of stream words (QOI_FFIX_SMEM) */
	const uint32_t img = blockIdx.x;
	const uint32_t lane = threadIdx.x;
	if (img >= P.n_images) return;
	const uint32_t live = P.f.n_live[img];
	const uint32_t rd = fsel_rd(P) * P.n_images + img;
	const uint32_t ver = P.f.verified[rd];
	const QoiImgMeta m = P.imgs[img];
	const uint8_t *s = P.streams + m.stream_off;
	if (live == 0u) {
		/* a valid header with no chunk bytes at all: the reference's loop (qoi.h:540-587) never parses an op, every pixel
		   is the start colour {0,0,0,255} (qoi.h:534-537).  A rejected header leaves the output untouched. */
		if (stream_ok(m, s, P.lens[img])) {
			uint8_t *o = P.out + m.dec_off;
			if (OUT_CH == 4) {
				for (uint32_t i = lane; i < m.npx; i += 32u) ((uint32_t *)o)[i] = 0xFF000000u;
			} else {
				for (uint64_t i = lane; i < (uint64_t)m.npx * 3u; i += 32u) o[i] = 0;
			}
		}
		return;
	}
	if (ver >= live) return; /* everything verified */
	const int32_t stop = P.lens[img] - QOI_END_BYTES;
	const uintptr_t s_lo = (uintptr_t)s & ~(uintptr_t)3, s_hi = ((uintptr_t)s + (uintptr_t)P.lens[img] + 3u) & ~(uintptr_t)3;
	const uint32_t base = P.f.seg_base[img];
	tab[lane] = P.f.ex_tab[(size_t)rd * 64 + lane];
	tab[lane + 32] = P.f.ex_tab[(size_t)rd * 64 + lane + 32];
	uint32_t px = P.f.ex_px[rd];
	__syncwarp();
	{
	auto body = [&](uint32_t j, const SegRec &r) {
		const uint32_t g = base + j;
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
		} else {
			const int32_t begin = QOI_HEADER_BYTES + (int32_t)j * QOI_FSEG_BYTES;
			const int32_t end = begin + QOI_FSEG_BYTES < stop ? begin + QOI_FSEG_BYTES : stop;
			/* all lanes copy the segment's words (and the few bytes an op may reach past `end`) into shared memory,
			   then lane 0 runs the reference loop on them; words outside the stream are never touched */
			const uintptr_t w0 = (uintptr_t)(s + begin) & ~(uintptr_t)3;
			const uintptr_t w1 = ((uintptr_t)(s + end) + 8u + 3u) & ~(uintptr_t)3;
			const uint32_t nw = (uint32_t)((w1 - w0) >> 2);
			for (uint32_t k = lane; k < nw; k += 32) {
				const uintptr_t wa = w0 + 4u * k;
				stage[k] = (wa >= s_lo && wa + 4 <= s_hi) ? __ldg((const uint32_t *)wa) : 0u;
			}
			__syncwarp();
			if (lane == 0) {
				const uint8_t *stage_b = (const uint8_t *)stage;
				const int32_t d = (int32_t)((intptr_t)(uintptr_t)s - (intptr_t)w0); /* stream byte q sits at stage_b[q + d] */
				const uint32_t px_begin = P.f.px_start[g];
				const uint32_t px_end = (j + 1 < live) ? P.f.px_start[g + 1] : m.npx;
				int32_t p = begin + (int32_t)P.f.entry[g];
				uint32_t run = 0;
				uint8_t *o = P.out + m.dec_off + (size_t)px_begin * OUT_CH;
				for (uint32_t i = px_begin; i < px_end; i++) {
					if (run > 0) {
						run--;
					} else if (p < end) {
						const uint32_t tag = stage_b[(p++) + d];
						if (tag == 0xFEu) {
							px = (px & 0xFF000000u) | (uint32_t)stage_b[p + d] | ((uint32_t)stage_b[p + 1 + d] << 8) | ((uint32_t)stage_b[p + 2 + d] << 16);
							p += 3;
						} else if (tag == 0xFFu) {
							px = (uint32_t)stage_b[p + d] | ((uint32_t)stage_b[p + 1 + d] << 8) | ((uint32_t)stage_b[p + 2 + d] << 16) | ((uint32_t)stage_b[p + 3 + d] << 24);
							p += 4;
						} else if ((tag & 0xC0u) == 0x00u) {
							px = tab[tag];
						} else if ((tag & 0xC0u) == 0x40u) {
							const uint32_t sel = ((tag >> 4) & 3u) | (((tag >> 2) & 3u) << 4) | ((tag & 3u) << 8) | 0x4000u;
							px = add_bytes(px, __byte_perm(0x0100FFFEu, 0u, sel));
						} else if ((tag & 0xC0u) == 0x80u) {
							const uint32_t b2 = stage_b[(p++) + d];
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
		}
		__syncwarp();
		px = __shfl_sync(QOI_FULL_MASK, px, 0);
	};
	walk_records(P, base, ver, live, lane, body);
	}
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
	f.redo_list = (uint32_t *)(base + off[18]);
	f.n_redo = (uint32_t *)(base + off[19]);
	f.verified = (uint32_t *)(base + off[14]);
	f.ex_px = (uint32_t *)(base + off[15]);
	f.ex_tab = (uint32_t *)(base + off[16]);
	f.bstate = (uint32_t *)(base + off[20]);
	f.vcount = (uint32_t *)(base + off[21]);
}

}  // namespace qoifdec
