/*
 * qoi_fdec.cuh -- parallel decode of ARBITRARY ("foreign") QOI streams on sm_100a (round-2 rewrite).
 *
 * The reference decoder (qoi.h:540-587) chains four things through the stream: the read cursor p, the current colour px,
 * the run counter and index[64].  A stream this library wrote comes with a restart index (the state at every 2048-pixel
 * tile, qoi_enc.cuh) and decodes one lane per tile (qoi_dec.cuh).  For any other stream that index does not exist -- so
 * this file BUILDS it, in parallel, and then runs the same lane-per-tile decoder:
 *
 *   1. parse (exact, parallel).  The stream is cut into segments of QOI_FSEG_BYTES bytes.  An op's length depends on its
 *      first byte only (0xFE:4, 0xFF:5, 10xxxxxx:2, else 1; qoi.h:547-575), so a segment maps "bytes to skip at entry ->
 *      (bytes to skip at exit, pixels produced)" for the 5 possible entry skips.  fparse_kernel tabulates the map,
 *      fchain_kernel composes the maps along each stream: every segment's true first op and first pixel.
 *   2. state rounds (parallel, O(ops), no pixels).  fstate_kernel walks a segment's ops from a GUESSED state (colour,
 *      index[]; round 0: index[] zeroed as the reference starts, the colour replayed from the tail of the segment before
 *      -- fzero_kernel) and records the state it ends with (last value written per slot, the colour) and what that result
 *      depended on: which guessed slots were read before the segment wrote them, whether the incoming r,g,b / alpha mattered.
 *      Whenever an op starts at or behind a 2048-pixel tile boundary it drops a restart record (pixel, stream offset,
 *      colour before, index[]) into the plan's index -- the same record the encoder leaves.
 *   3. scan + check (parallel).  "State before segment j" is a last-writer scan over the segments' results (per slot: the
 *      nearest earlier segment that wrote it; the colour: the previous segment's): fscan_a / fscan_b, 64 segments per
 *      chunk.  fverify_kernel compares every segment's guess with the scanned state on what the segment depended on,
 *      gives the segments that disagree the scanned state as their new guess and lists them for the next round.  A round in which nothing disagrees is a proof by induction that every result is exact (segment 0 starts
 *      from the reference's initial state).  One more round then re-walks the segments whose guess differs from the exact
 *      state in ANY slot, so that the restart records they left are exact in all 64 slots.
 *   4. fall back (exact, sequential): streams that have not settled after QOI_FROUNDS rounds (content whose colours are one
 *      long chain of DIFF / LUMA / INDEX ops without any absolute op) are finished by ffall_kernel, one walk per stream
 *      from the first unsettled segment with the exact state.
 *   5. pixels: qoi_decode_kernel with the index, one lane per 2048-pixel tile, for every kind of content (a stream of
 *      nothing but RUN ops has 254 000 pixels per 4096-byte segment; one lane per SEGMENT writing them was the slowest
 *      case of round 1).  The index stays in the plan: it can be exported as a side-car (qoi_gpu_index_export).
 *
 * Round 1 verified with an in-order walk per stream (13 000 segments x 10 passes, 0.8 ms each whatever the batch) and
 * tracked the source of every value to re-base stale results; the scans replace the walks, a stale result is simply redone.
 * Every kernel cites the reference rule it restates; nothing here reads outside [stream, stream + size).
 */
#pragma once
#include "qoi_common.h"
#include "qoi_dev.cuh"
#include "qoi_dec.cuh" /* the chunked stream reader (qoidec::Reader) */

namespace qoifdec {
using namespace qoidev;

#define QOI_FSEG_BYTES 1024
#define QOI_FROUNDS 16  /* state rounds before the sequential fall-back */
#define QOI_FCHUNK 64   /* segments per scan chunk */
#define QOI_FDEC_THREADS 64
#define QOI_FSCAN_THREADS 96 /* 64 slot threads + the colour thread (thread 64); a multiple of 32 */
#define QOI_FSTATE_THREADS 128 /* state rounds: the op table (qoidec::op_lut_entry, 4 KB) is shared by more lanes */
#define QOI_FDEC_SMEM (QOI_DEC_LUT_BYTES + (64 + 16 + QOI_DEC_RING_WORDS) * QOI_FSTATE_THREADS * 4) /* op table; per lane: index[64], 64 flag bytes, stream ring */
#define QOI_FFALL_SMEM ((64 + QOI_FSEG_BYTES / 4 + 8) * 4) /* index[], one staged segment */

/* flag[] bits (fverify_kernel) */
#define QOI_FCONT_LIST 4096u /* a round with at most this many listed segments lets the walks continue into the next segment ... */
#ifndef QOI_FCONT_SHIFT
#define QOI_FCONT_SHIFT 5    /* ... or at most n_segs >> this many: the limit has to grow with the batch.  (A fixed 4096 was right for
                                32 x 4096^2 and a cliff for anything larger: 32 x 8192^2 and 1024 x 1080p kept lists of a few
                                thousand segments per round, never continued, ran out of rounds and fell back to the sequential
                                walk -- 2.1 s and 142 ms, profiles/r2_state/bench_other_configs.txt.) */
#endif
#define QOI_FCONT_MAX 48u    /* ... this many segments at most */
#define QOI_FF_DEP 1u  /* the guess disagrees with the scanned state on something the segment depended on */
#define QOI_FF_TRUST 4u /* the segment's last walk started from a guess that held on everything it depended on: what it derived from
                           the guess can be trusted (segment 0 always: its "guess" is the reference's initial state) */
#define QOI_FF_CGUESS 8u /* round 0 only: the guessed colour is the replayed tail of the segment before (fzero_kernel), not the default */

struct QoiFState {
	/* geometry (host written) */
	const uint32_t *seg_base;   /* [n_images + 1] first global segment of each image (capacity, not live count) */
	const uint32_t *chunk_base; /* [n_images + 1] first scan chunk of each image */
	/* parse */
	uint32_t *px_cnt;   /* [segs][5] pixels produced when entered with skip s */
	uint8_t *exit_skip; /* [segs][8] (5 used) */
	uint32_t *last_rel; /* [segs][5] pixels produced before the segment's LAST op starts */
	uint16_t *last_abs; /* [segs] offset of the last RGB / RGBA op that starts in the segment (walk 0), 0xFFFF: none */
	/* chain result */
	uint8_t *entry;     /* [segs] distance from the segment start to its first op */
	uint32_t *px_start; /* [segs] first pixel of the segment */
	uint32_t *tb0;      /* [segs] the first tile boundary no op of an earlier segment starts behind: the segment's ops own the restart
	                       records from there on */
	uint32_t *n_live;   /* [n_images] segments that hold at least one op of a pixel < npx */
	/* state rounds */
	uint32_t *g_px;     /* [segs] guessed colour before the segment */
	uint32_t *g_tab;    /* [segs][64] guessed index[] */
	uint32_t *o_px;     /* [segs] colour after the segment */
	uint32_t *o_tab;    /* [segs][64] index[] after the segment (meaningful where o_wr has the bit) */
	uint64_t *o_wr;     /* [segs] slots the segment wrote */
	uint64_t *o_taint;  /* [segs] ... with a value that derives from the guess (incoming colour or a guessed slot) */
	uint64_t *rbw;      /* [segs] slots of the guess read before the segment wrote them */
	uint32_t *dep;      /* [segs] bit 0: the result depends on the incoming r,g,b; bit 1: ... on the incoming alpha;
	                       bits 2,3: r,g,b / alpha of the colour after the segment derive from the guess;
	                       bits 4,5: ... and do so as "incoming colour + deltas" (DIFF / LUMA / RUN ops only);
	                       bit 6: the walk read a guessed slot that can only be an initial zero (see fstate_kernel) */
	uint8_t *flag;      /* [segs] QOI_FF_* of the last check */
	uint32_t *list;     /* [segs] segments to walk in the next round */
	uint32_t *n_list;   /* [64] length of that list per round */
	uint32_t *active;   /* [64] streams a round's scan still looked at (0: every later launch returns at once) */
	/* scan */
	uint32_t *a_tab;    /* [chunks][64] last value written per slot inside the chunk */
	uint64_t *a_wr;     /* [chunks] slots written inside the chunk */
	uint64_t *a_px;     /* [chunks] the chunk's colour map (seg_colour_map composed over its segments) */
	uint32_t *p_tab;    /* [chunks][64] index[] before the chunk */
	uint32_t *p_px;     /* [chunks] colour before the chunk */
	uint32_t *bad_cnt;  /* [4][n_images]: 0,1: segments the check of an even / odd round could not prove; 2: the stream has settled;
	                       3: the last round that looked at the stream */
	uint32_t *first_bad;/* [2][n_images] first unproven segment of that check (by round parity): where the fall-back starts */
	/* per tile of the plan: which slots of the restart record's index[] the segment had written when it dropped the record
	   (the others are filled from the segment's exact incoming state at the end), and which segment that was */
	uint64_t *rec_wr;
	uint32_t *rs_seg;
	uint32_t *claim;    /* [segs] (epoch << 8 | round + 1) of the last walk that took the segment over by continuing into it */
	uint32_t *x_state;  /* [n_images][65] exact index[] and colour before the first unproven segment (for the fall-back) */
};

struct QoiFParams {
	const uint8_t *streams;
	const int32_t *lens;
	int32_t *status;
	const QoiImgMeta *imgs;
	uint32_t n_images;
	uint32_t n_segs;   /* capacity over all images */
	uint32_t n_chunks; /* capacity */
	uint32_t round;    /* state round of this launch (0 = every live segment, from the reference's initial state) */
	uint32_t epoch;    /* counts the decode calls on the plan (claim[] is never cleared) */
	QoiFState f;
	QoiTileState st;   /* the plan's restart index: rs_px / rs_byte / rs_prev / ent are (re)written here */
};

__device__ __forceinline__ uint32_t add_bytes(uint32_t a, uint32_t b) {
	return ((a & 0x7F7F7F7Fu) + (b & 0x7F7F7F7Fu)) ^ ((a ^ b) & 0x80808080u);
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
/* (Measured alternatives for the byte fetch of walk 0, 32 x 4096^2 photographs, plain byte loads = 1.22 ms: 16-byte pieces held
   in registers with the byte picked by selects 1.80 ms; the decoders' chunk ring in shared memory (256-bit loads, one
   shared-memory read per op) 1.61 ms -- the byte loads mostly hit L1 and cost fewer instructions than either.) */
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
	uint32_t total0 = 0, last0 = 0; /* last0: pixels before the last op */
	uint32_t abs0 = 0xFFFFu;        /* where the segment's last RGB / RGBA op starts (fguess_kernel) */
	int32_t p0 = begin;
	while (p0 < end) {
		last0 = total0;
		if (p0 < near) {
			const uint32_t off = (uint32_t)(p0 - begin);
			if (off < 64u) seen_lo |= 1ull << off;
			else seen_hi |= 1ull << (off - 64u);
		}
		const uint32_t tag = s[p0];
		if (tag >= 0xFEu) abs0 = (uint32_t)(p0 - begin);
		p0 += op_len(tag);
		total0 += op_px(tag);
	}
	P.f.last_abs[g] = (uint16_t)abs0;
	const uint32_t exit0 = (uint32_t)(p0 - end);
	P.f.px_cnt[(size_t)g * 5] = total0;
	P.f.last_rel[(size_t)g * 5] = last0;
	P.f.exit_skip[(size_t)g * 8] = (uint8_t)exit0;
#pragma unroll 1
	for (int sk = 1; sk < 5; sk++) {
		int32_t p = begin + sk;
		uint32_t cnt = 0, lastk = 0;
		bool merged = false;
		while (p < near) {
			const uint32_t off = (uint32_t)(p - begin);
			if (((off < 64u ? seen_lo : seen_hi) >> (off & 63u)) & 1ull) { merged = true; break; }
			const uint32_t tag = s[p];
			p += op_len(tag);
			lastk = cnt;
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
			lastk = cnt - (total0 - last0); /* from the meeting point on it is walk 0: the same last op */
			ex = exit0;
		} else {
			while (p < end) {
				const uint32_t tag = s[p];
				p += op_len(tag);
				lastk = cnt;
				cnt += op_px(tag);
			}
			ex = p >= end ? (uint32_t)(p - end) : 0u; /* an entry skip beyond a very short last segment */
		}
		P.f.px_cnt[(size_t)g * 5 + sk] = cnt;
		P.f.last_rel[(size_t)g * 5 + sk] = lastk;
		P.f.exit_skip[(size_t)g * 8 + sk] = (uint8_t)ex;
	}
}

/* which image owns scan chunk c (binary search in chunk_base) */
__device__ __forceinline__ uint32_t chunk_image(const QoiFParams &P, uint32_t c) {
	uint32_t lo = 0, hi = P.n_images - 1;
	while (lo < hi) {
		const uint32_t mid = (lo + hi + 1) >> 1;
		if (P.f.chunk_base[mid] <= c) lo = mid;
		else hi = mid - 1;
	}
	return lo;
}

/* ------------------------------------------------------------------ 1b. chain: one CTA per stream */
/* Composing the maps (skip -> exit skip, + pixels) is associative: thread t composes the maps of its block of consecutive
   segments for all 5 entry skips, thread 0 chains the QOI_FCHAIN_THREADS block maps from the stream's start (skip 0,
   pixel 0), then every thread walks its block again from its now known start and writes the per-segment results.  (Round
   2's first version chained 32 segments at a time through one lane: 1.3 ms for a 4096x4096 photograph, more than the
   state round it feeds.)  Also resets what the state rounds accumulate per image and marks every tile of the image as
   "no op starts here" (the state rounds overwrite the tiles in which one does). */
#define QOI_FCHAIN_THREADS 1024
#define QOI_FCHAIN_SMEM ((QOI_FCHAIN_THREADS * 8 + 32 * 8 + 8) * 4)
__global__ void __launch_bounds__(QOI_FCHAIN_THREADS) fchain_kernel(const QoiFParams P) {
	QOI_DYNAMIC_SMEM(smem); /* QOI_FCHAIN_SMEM */
	uint32_t *s_map = (uint32_t *)smem;                  /* [thread][8]: exit skip per entry skip (packed), 5 pixel counts; then: start skip, start pixel */
	uint32_t *s_grp = s_map + QOI_FCHAIN_THREADS * 8;    /* [32][8]: the same for groups of 32 threads */
	uint32_t *s_live = s_grp + 32 * 8;                   /* live segments */
	const uint32_t img = blockIdx.x;
	const uint32_t t = threadIdx.x;
	if (img >= P.n_images) return;
	const QoiImgMeta m = P.imgs[img];
	const uint8_t *s = P.streams + m.stream_off;
	const int32_t size = P.lens[img];
	const bool ok = stream_ok(m, s, size);
	if (t == 0 && P.status) P.status[img] = ok ? 0 : 4; /* QOI_GPU_E_STREAM */
	if (t == 0) s_live[0] = 0;
	const uint32_t tiles = (m.npx + QOI_TILE_PX - 1) / QOI_TILE_PX;
	for (uint32_t k = t; k < tiles; k += QOI_FCHAIN_THREADS) P.st.rs_px[m.first_tile + k] = -1;
	const int32_t stop = size - QOI_END_BYTES;
	const uint32_t base = P.f.seg_base[img];
	const uint32_t nseg = (ok && stop > QOI_HEADER_BYTES) ? (uint32_t)((stop - QOI_HEADER_BYTES + QOI_FSEG_BYTES - 1) / QOI_FSEG_BYTES) : 0u;
	const uint32_t per = (nseg + QOI_FCHAIN_THREADS - 1) / QOI_FCHAIN_THREADS;
	const uint32_t j_lo = t * per < nseg ? t * per : nseg, j_hi = j_lo + per < nseg ? j_lo + per : nseg;
	{ /* my block's map */
		uint32_t sk[5] = {0u, 1u, 2u, 3u, 4u}, px[5] = {0u, 0u, 0u, 0u, 0u};
#pragma unroll 4
		for (uint32_t j = j_lo; j < j_hi; j++) { /* all five counts are fetched whatever the skips turn out to be: the loads of several segments overlap */
			const uint32_t g = base + j;
			const uint2 e = *(const uint2 *)(P.f.exit_skip + (size_t)g * 8);
			uint32_t c[5];
#pragma unroll
			for (int k = 0; k < 5; k++) c[k] = P.f.px_cnt[(size_t)g * 5 + k];
#pragma unroll
			for (int k = 0; k < 5; k++) {
				const uint32_t q = sk[k];
				px[k] += q == 0u ? c[0] : q == 1u ? c[1] : q == 2u ? c[2] : q == 3u ? c[3] : c[4];
				sk[k] = (q < 4u ? (e.x >> (8u * q)) : e.y) & 0xFFu;
			}
		}
		s_map[t * 8] = sk[0] | (sk[1] << 4) | (sk[2] << 8) | (sk[3] << 12) | (sk[4] << 16);
#pragma unroll
		for (int k = 0; k < 5; k++) s_map[t * 8 + 1 + k] = px[k];
	}
	__syncthreads();
	if (t < 32u) { /* the maps of 32 consecutive threads, composed */
		uint32_t sk[5] = {0u, 1u, 2u, 3u, 4u}, px[5] = {0u, 0u, 0u, 0u, 0u};
		for (uint32_t q = 0; q < 32u; q++) {
			const uint32_t *mp = s_map + (t * 32u + q) * 8;
			const uint32_t e = mp[0];
#pragma unroll
			for (int k = 0; k < 5; k++) {
				px[k] += mp[1 + sk[k]];
				sk[k] = (e >> (4u * sk[k])) & 0xFu;
			}
		}
		s_grp[t * 8] = sk[0] | (sk[1] << 4) | (sk[2] << 8) | (sk[3] << 12) | (sk[4] << 16);
#pragma unroll
		for (int k = 0; k < 5; k++) s_grp[t * 8 + 1 + k] = px[k];
	}
	__syncthreads();
	if (t == 0) { /* the stream starts at an op (skip 0), pixel 0 */
		uint32_t skip = 0, px = 0;
		for (uint32_t q = 0; q < 32u; q++) {
			const uint32_t e = s_grp[q * 8], c = s_grp[q * 8 + 1 + skip];
			s_grp[q * 8 + 6] = skip;
			s_grp[q * 8 + 7] = px;
			px += c;
			skip = (e >> (4u * skip)) & 0xFu;
		}
	}
	__syncthreads();
	if (t < 32u) {
		uint32_t skip = s_grp[t * 8 + 6], px = s_grp[t * 8 + 7];
		for (uint32_t q = 0; q < 32u; q++) {
			uint32_t *mp = s_map + (t * 32u + q) * 8;
			const uint32_t e = mp[0], c = mp[1 + skip];
			mp[6] = skip;
			mp[7] = px;
			px += c;
			skip = (e >> (4u * skip)) & 0xFu;
		}
	}
	__syncthreads();
	{
		uint32_t skip = s_map[t * 8 + 6], px = s_map[t * 8 + 7];
		uint32_t live = 0;
		/* The tile boundaries behind a segment's last op start belong to later ops: every segment writes tb0 of the segment
		   BEHIND it (the stream's very first op owns tile 0's record: 0 for segment 0). */
		if (t == 0 && nseg) P.f.tb0[base] = 0u;
#pragma unroll 4
		for (uint32_t j = j_lo; j < j_hi; j++) {
			const uint32_t g = base + j;
			const uint2 e = *(const uint2 *)(P.f.exit_skip + (size_t)g * 8);
			uint32_t c5[5], l5[5];
#pragma unroll
			for (int k = 0; k < 5; k++) { c5[k] = P.f.px_cnt[(size_t)g * 5 + k]; l5[k] = P.f.last_rel[(size_t)g * 5 + k]; }
			if (px >= m.npx) break; /* all pixels already produced: this and every later segment is dead */
			const uint32_t c = skip == 0u ? c5[0] : skip == 1u ? c5[1] : skip == 2u ? c5[2] : skip == 3u ? c5[3] : c5[4];
			const uint32_t lr = skip == 0u ? l5[0] : skip == 1u ? l5[1] : skip == 2u ? l5[2] : skip == 3u ? l5[3] : l5[4];
			P.f.entry[g] = (uint8_t)skip;
			P.f.px_start[g] = px;
			P.f.g_px[g] = 0xFF000000u; /* round-0 guess: the reference's initial state (qoi.h:533-537); exact for segment 0 */
			P.f.flag[g] = j == 0u ? (uint8_t)QOI_FF_TRUST : (uint8_t)0;
			/* a segment without an op start (only the last one can be that short) leaves the boundary where it was: nothing behind it reads it */
			if (j + 1u < nseg) P.f.tb0[g + 1u] = c ? ((px + lr) / QOI_TILE_PX + 1u) * QOI_TILE_PX : 0u;
			px += c;
			skip = (skip < 4u ? (e.x >> (8u * skip)) : e.y) & 0xFFu;
			live = j + 1u;
		}
		if (live) atomicMax((int *)&s_live[0], (int)live);
	}
	__syncthreads();
	if (t == 0) {
		const uint32_t live = s_live[0];
		if (ok && live == 0u) {
			/* a valid header and not one op: the reference's loop never parses, every pixel is the start colour
			   (qoi.h:534-544).  One restart record at pixel 0 makes the tile decoder produce exactly that. */
			P.st.rs_px[m.first_tile] = 0;
			P.st.rs_byte[m.first_tile] = QOI_HEADER_BYTES;
			P.st.rs_prev[m.first_tile] = 0xFF000000u;
		}
		P.f.n_live[img] = live;
		P.f.bad_cnt[img] = P.f.bad_cnt[P.n_images + img] = P.f.bad_cnt[2u * P.n_images + img] = P.f.bad_cnt[3u * P.n_images + img] = 0;
		P.f.first_bad[img] = P.f.first_bad[P.n_images + img] = 0x7FFFFFFFu;
	}
}

/* round-0 guess of index[]: zeros (qoi.h:533); list counters */
__global__ void __launch_bounds__(QOI_FDEC_THREADS) fzero_kernel(const QoiFParams P) {
	const uint32_t g = blockIdx.x * QOI_FDEC_THREADS + threadIdx.x;
	if (g < 64u) P.f.n_list[g] = P.f.active[g] = 0u;
	if (g >= P.n_segs) return;
	const uint32_t img = seg_image(P, g);
	if (g - P.f.seg_base[img] >= P.f.n_live[img]) return;
	uint32_t *t = P.f.g_tab + (size_t)g * 64;
#pragma unroll
	for (int k = 0; k < 8; k++) st_256(t + 8 * k, make_uint4(0u, 0u, 0u, 0u), make_uint4(0u, 0u, 0u, 0u));
	/* Round-0 guess of the colour that enters the segment.  The reference's start colour is right for segment 0 only; for
	   the others the colour that LEAVES the segment before is usually known from that segment's own bytes: replay its tail
	   from its last RGB / RGBA op (fparse_kernel noted where it starts) -- DIFF, LUMA and RUN ops only add to it; an INDEX
	   op on the way, or no absolute op at all, leaves the default.  r,g,b are then exact and alpha is whatever an RGBA op
	   said, else opaque; it is a guess like any other (a walk of the segment before that does not pass through that op
	   start makes it wrong, the check finds out), but most walks of round 0 now start on the right colour and what they
	   derive from it survives the first check. */
	const uint32_t j = g - P.f.seg_base[img];
	if (j == 0u) return;
	const uint32_t off = P.f.last_abs[g - 1u];
	if (off == 0xFFFFu) return;
	const QoiImgMeta m = P.imgs[img];
	const uint8_t *s = P.streams + m.stream_off;
	const int32_t mine = QOI_HEADER_BYTES + (int32_t)j * QOI_FSEG_BYTES; /* the ops that start before this belong to the segment before */
	int32_t p = mine - QOI_FSEG_BYTES + (int32_t)off;
	uint32_t tag = s[p];
	uint32_t px = (uint32_t)s[p + 1] | ((uint32_t)s[p + 2] << 8) | ((uint32_t)s[p + 3] << 16) | (tag == 0xFFu ? (uint32_t)s[p + 4] << 24 : 0xFF000000u);
	p += op_len(tag);
	bool known = true;
	while (p < mine && known) {
		tag = s[p];
		const uint32_t top = tag >> 6;
		if (top == 0u || tag >= 0xFEu) known = false; /* INDEX: the colour comes out of index[]; (an absolute op cannot follow the last one) */
		else if (top == 1u) px = add_bytes(px, ((((tag >> 4) & 3u) - 2u) & 0xFFu) | (((((tag >> 2) & 3u) - 2u) & 0xFFu) << 8) | ((((tag & 3u) - 2u) & 0xFFu) << 16)); /* qoi.h:561-565 */
		else if (top == 2u) { /* qoi.h:566-572 */
			const uint32_t b2 = s[p + 1], vg = (tag & 0x3Fu) - 32u;
			px = add_bytes(px, ((vg - 8u + ((b2 >> 4) & 0xFu)) & 0xFFu) | ((vg & 0xFFu) << 8) | (((vg - 8u + (b2 & 0xFu)) & 0xFFu) << 16));
		}
		p += op_len(tag);
	}
	if (known) {
		P.f.g_px[g] = px;
		P.f.flag[g] = (uint8_t)QOI_FF_CGUESS;
	}
}

/* proven prefix of a stream, as a global segment index: what the check of the round before established */
__device__ __forceinline__ uint32_t stream_proven_end(const QoiFParams &P, uint32_t img) {
	const uint32_t base = P.f.seg_base[img];
	if (P.round == 0u) return base + 1u;
	const uint32_t key = P.f.first_bad[((P.round - 1u) & 1u) * P.n_images + img]; /* 2 * segment + (not a disagreement) */
	const uint32_t pe = (key >> 1) + (key & 1u);
	return base + (pe > 1u ? pe : 1u);
}
/* a lane's index[] (shared memory, tab[slot * B]) <-> 64 consecutive words in global memory, eight slots per 256-bit access:
   the lanes of a warp work on 32 different segments, so every access costs the L1 data pipe one pass per lane whatever its
   width -- word by word, the 512 bytes a walk reads and leaves per segment were more passes than its ops */
template <int B>
__device__ __forceinline__ void tab_load_256(uint32_t *tab, const uint32_t *src) {
#pragma unroll
	for (int k = 0; k < 8; k++) {
		uint4 a, b;
		ld_256(src + 8 * k, a, b);
		uint32_t *t = tab + (8 * k) * B;
		t[0 * B] = a.x; t[1 * B] = a.y; t[2 * B] = a.z; t[3 * B] = a.w;
		t[4 * B] = b.x; t[5 * B] = b.y; t[6 * B] = b.z; t[7 * B] = b.w;
	}
}
template <int B>
__device__ __forceinline__ void tab_store_256(uint32_t *dst, const uint32_t *tab) {
#pragma unroll
	for (int k = 0; k < 8; k++) {
		const uint32_t *t = tab + (8 * k) * B;
		st_256(dst + 8 * k, make_uint4(t[0 * B], t[1 * B], t[2 * B], t[3 * B]), make_uint4(t[4 * B], t[5 * B], t[6 * B], t[7 * B]));
	}
}

/* ------------------------------------------------------------------ 2. state round: one lane per segment, one trip per OP */
/* index[] in shared memory [slot][thread] like qoi_decode_kernel; stream bytes through the same chunked reader. */
__device__ __forceinline__ void fstate_walk(const QoiFParams &P, uint32_t g, const uint4 *lut, uint32_t *tab, uint32_t *flg, uint32_t *ring) {
	constexpr int B = QOI_FSTATE_THREADS;
	const uint32_t img = seg_image(P, g);
	uint32_t j = g - P.f.seg_base[img];
	const uint32_t live = P.f.n_live[img];
	if (j >= live) return;
	const QoiImgMeta m = P.imgs[img];
	const uint8_t *s = P.streams + m.stream_off;
	const int32_t stop = P.lens[img] - QOI_END_BYTES;
	int32_t begin = QOI_HEADER_BYTES + (int32_t)j * QOI_FSEG_BYTES;
	int32_t end = begin + QOI_FSEG_BYTES < stop ? begin + QOI_FSEG_BYTES : stop;
	/* Late rounds list a handful of segments, typically the front of a chain: a smooth region whose colours all hang on the
	   colour that entered it settles one segment per round, and a round costs three more launches.  So when the list is
	   short a walk does not stop at its segment's end: while the state it arrives with disagrees with the NEXT segment's
	   guess on what that segment depended on, it takes the next segment over (new guess = the state it holds) and goes on.
	   claim[] keeps two walks of one round off the same segment. */
	const uint32_t claim_tag = (P.epoch << 8) | (P.round + 1u);
	const uint32_t cont_limit = (P.n_segs >> QOI_FCONT_SHIFT) > QOI_FCONT_LIST ? (P.n_segs >> QOI_FCONT_SHIFT) : QOI_FCONT_LIST;
	/* Only the walk of the stream's FIRST UNPROVEN segment continues: everything before it is proven, so its guess is the exact
	   state and what it arrives with is exact too.  Any other walk may have started from a wrong guess; letting it overwrite
	   its successors' guesses dragged wrong state down the stream 48 segments per round with the correction 48 segments
	   behind -- an opaque stream (every segment depends on the incoming alpha) that once guessed alpha 0 never settled:
	   6 of 512 1080p photographs and every 8192^2 one went to the sequential fall-back (profiles/r2_state). */
	const bool cont = P.round >= 2u && P.f.n_list[P.round] <= cont_limit && g == stream_proven_end(P, img);
	if (cont && atomicExch(&P.f.claim[g], claim_tag) == claim_tag) return; /* a walk that started earlier got here first */

	uint32_t px = P.f.g_px[g];
	tab_load_256<B>(tab, P.f.g_tab + (size_t)g * 64);
	uint64_t rbw;
	/* What the current colour derives from, one bit each (cf), r,g,b and alpha separately (an RGB op replaces only r,g,b):
	     T_*: from the guess at all (the colour that entered the segment, or a guessed index[] slot read before the segment
	          wrote it);
	     C_*: from the guessed COLOUR in particular -- only that makes the segment depend on the incoming colour (a colour
	          read out of a guessed slot is covered by rbw, and a stream of INDEX ops must not wait for its predecessor's last
	          colour);
	     D_*: ... through index[], i.e. not as "incoming colour + deltas of DIFF / LUMA / RUN ops", which the scan can
	          re-base exactly.
	   Every slot the segment writes keeps W | T_* | C_* of the colour stored there in a byte of flg[], so that an INDEX op
	   of a slot the segment filled itself inherits them. */
	constexpr uint32_t F_W = 1u, F_TRGB = 2u, F_TA = 4u, F_CRGB = 8u, F_CA = 16u, F_DRGB = 32u, F_DA = 64u;
	uint32_t cf, depacc;
	bool suspect;
	auto flag_byte = [&](uint32_t slot) -> uint8_t * { return (uint8_t *)(flg + (slot >> 2) * B) + (slot & 3u); };
	/* the W bits / (T_RGB | T_A) bits of all 64 slots as masks */
	auto gather = [&](uint64_t &wr, uint64_t &taint) {
		wr = 0; taint = 0;
#pragma unroll 4
		for (int q = 0; q < 16; q++) {
			const uint32_t w = flg[q * B];
			const uint32_t wb = w & 0x01010101u;                                 /* W of slots 4q .. 4q+3 at bits 0,8,16,24 */
			const uint32_t tb4 = ((w >> 1) | (w >> 2)) & 0x01010101u;
			const uint32_t wn = (wb | (wb >> 7) | (wb >> 14) | (wb >> 21)) & 0xFu; /* ... at bits 0..3 */
			const uint32_t tn = (tb4 | (tb4 >> 7) | (tb4 >> 14) | (tb4 >> 21)) & 0xFu;
			wr |= (uint64_t)wn << (4 * q);
			taint |= (uint64_t)tn << (4 * q);
		}
	};
	uint32_t pos = P.f.px_start[g];                                          /* pixel the next op starts at */
	/* The restart record of a tile belongs to the first op that starts at or behind the tile's first pixel.  `tb` = the next
	   tile boundary no op has started behind yet; fchain_kernel knows where every segment's last op starts, so the first
	   boundary that is this segment's to take is exact (tb0) and no two lanes ever write the same record. */
	uint32_t tb = P.f.tb0[g];
	const int32_t p0 = begin + (int32_t)P.f.entry[g];

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
	uint32_t filled = 2;
	auto next_word = [&]() -> uint32_t {
		const uint32_t w = ring[(rd.wi & (QOI_DEC_RING_WORDS - 1)) * B];
		rd.wi++;
		return w;
	};
	/* warp-uniform refill of the ring, once per trip (see qoi_dec.cuh) */
	auto refill = [&]() {
		if ((rd.wi >> 3) + 1u >= filled) {
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

	/* the tile's restart record = the decoder's state right before the op at (pos, rd.p).  index[] goes out as it is in this
	   lane -- exact in the slots the segment has written so far (rec_wr); the other slots still hold the GUESS, and only the
	   ones the segment depends on are known to be right: fcompose_kernel overwrites them with the exact state before the
	   segment once that is known. */
	auto drop_record = [&]() {
		const uint32_t T = m.first_tile + tb / QOI_TILE_PX;
		P.st.rs_px[T] = (int32_t)pos;
		P.st.rs_byte[T] = rd.p;
		P.st.rs_prev[T] = px;
		{
			uint64_t wr, taint;
			gather(wr, taint);
			P.f.rec_wr[T] = wr;
		}
		P.f.rs_seg[T] = g;
		if (tb) { /* qoi_decode_kernel reads index[] of tile T from the record of tile T - 1; tile 0 starts from zeros */
			uint64_t *dst = P.st.ent + (size_t)(T - 1) * 64;
#pragma unroll 4
			for (int q = 0; q < 16; q++) /* value << 32 | 1, four slots per store */
				st_256(dst + 4 * q, make_uint4(1u, tab[(4 * q) * B], 1u, tab[(4 * q + 1) * B]), make_uint4(1u, tab[(4 * q + 2) * B], 1u, tab[(4 * q + 3) * B]));
		}
		tb += QOI_TILE_PX;
	};
	for (uint32_t steps = 0;; steps++) {
	rbw = 0;
	cf = F_TRGB | F_TA | F_CRGB | F_CA;
	depacc = 0;
	suspect = false;
#pragma unroll
	for (int q = 0; q < 16; q++) flg[q * B] = 0u;
	while (rd.p < end && pos < m.npx) { /* qoi.h:540,544: no op is parsed once every pixel is there */
		if (rd.avail < 5u) { /* an op needs at most 5 bytes */
			rd.buf |= (uint64_t)next_word() << (8u * rd.avail);
			rd.avail += 4u;
			if (rd.avail < 5u) {
				rd.buf |= (uint64_t)next_word() << (8u * rd.avail);
				rd.avail += 4u;
			}
		}
		if (pos >= tb) drop_record(); /* an op starts at or behind a tile boundary */
		const uint32_t tag = (uint32_t)rd.buf & 0xFFu;
		const uint32_t pay = (uint32_t)(rd.buf >> 8);
		const uint32_t top = tag >> 6, t6 = tag & 63u;
		const bool is_rgb = tag == 0xFEu, is_rgba = tag == 0xFFu;   /* qoi.h:547-557 are tested before the 2-bit tags */
		const bool is_idx = top == 0u;
		const uint4 L = lut[tag];                                    /* what the first byte says (qoidec::op_lut_entry) */
		const uint32_t idxv = tab[t6 * B];                           /* qoi.h:558-560 */
		const uint32_t npx = qoidec::op_colour(tag, pay, L, idxv, px); /* qoi.h:547-575 */
		const uint32_t len = __dp4a(L.y, 0x00000100u, 0u);
		const uint32_t npix = __dp4a(L.y, 0x00010000u, 0u) + 1u;     /* a RUN op yields run + 1 pixels (qoi.h:541-543,573-575) */
		/* where the new colour comes from */
		uint8_t *fb_t = flag_byte(t6);
		const uint32_t fbv = *fb_t;
		const bool own_slot = (fbv & F_W) != 0u;
		if (is_idx && !own_slot) rbw |= 1ull << t6; /* the guessed index[] is used */
		/* index[t] only ever holds a colour that hashes to t, or the initial zero (qoi.h:533,577): a guessed slot whose value
		   does not hash to its own number is an initial zero -- which an encoder never refers to (only a hand-made stream
		   does).  Almost certainly the guess is wrong there, and {0,0,0,0} would put alpha 0 into everything this segment
		   derives from it: mark the walk, so that it is not trusted on a mere re-walk (seg_trusted). */
		suspect = suspect || (is_idx && !own_slot && (qoi_slot(idxv) != t6 || idxv == 0u)); /* slot 0's zero hashes to 0: as much a guess as the others */
		/* INDEX of a slot the segment filled itself: the stored colour's provenance (from the incoming colour -> still
		   incoming colour + deltas; from a guessed slot -> D); of a guessed slot: D */
		const uint32_t own_t = fbv & (F_TRGB | F_TA), own_c = (fbv & (F_CRGB | F_CA)) >> 2;
		const uint32_t cf_idx = own_slot ? ((fbv & (F_TRGB | F_TA | F_CRGB | F_CA)) | ((own_t & ~own_c) << 4)) : (F_TRGB | F_TA | F_DRGB | F_DA);
		uint32_t ncf = is_rgb ? (cf & (F_TA | F_CA | F_DA)) : cf; /* RGB: r,g,b from the stream, alpha kept; DIFF / LUMA / RUN: unchanged */
		ncf = is_rgba ? 0u : ncf;
		cf = is_idx ? cf_idx : ncf;
		/* every op yields at least one pixel of the new colour; while that colour derives from the guessed COLOUR the result
		   depends on it */
		depacc |= cf;
		px = npx;
		rd.buf >>= 8u * len;
		rd.avail -= len;
		rd.p += (int32_t)len;
		const uint32_t h = qoi_slot(px);
		tab[h * B] = px; /* qoi.h:577 */
		*flag_byte(h) = (uint8_t)(F_W | (cf & (F_TRGB | F_TA | F_CRGB | F_CA)));
		pos += npix;
		refill();
	}
	/* the colour that leaves the segment (and the pixels a truncated stream leaves over, qoi.h:544) may still be the incoming one */
	depacc |= cf;
	{
		uint64_t wr, taint;
		gather(wr, taint);
		P.f.o_wr[g] = wr;
		P.f.o_taint[g] = taint & wr;
	}
	P.f.o_px[g] = px;
	P.f.rbw[g] = rbw;
	/* bits 0,1: the result depends on the incoming r,g,b / alpha; bits 2,3: r,g,b / alpha of the final colour derive from the
	   guess; bits 4,5: ... as incoming colour + deltas only */
	P.f.dep[g] = ((depacc & F_CRGB) ? 1u : 0u) | ((depacc & F_CA) ? 2u : 0u) | ((cf & F_TRGB) ? 4u : 0u) | ((cf & F_TA) ? 8u : 0u) |
	             (((cf & (F_TRGB | F_DRGB)) == F_TRGB) ? 16u : 0u) | (((cf & (F_TA | F_DA)) == F_TA) ? 32u : 0u) | (suspect ? 64u : 0u);
	tab_store_256<B>(P.f.o_tab + (size_t)g * 64, tab);
	if (!cont || steps + 1u >= QOI_FCONT_MAX || j + 1u >= live) break;
	{ /* does the next segment's last walk stand with the state this one arrives with? */
		const uint32_t gn = g + 1u;
		const uint32_t dn = P.f.dep[gn];
		const uint32_t pmask = ((dn & 1u) ? 0x00FFFFFFu : 0u) | ((dn & 2u) ? 0xFF000000u : 0u);
		bool differs = ((P.f.g_px[gn] ^ px) & pmask) != 0u;
		uint64_t need = P.f.rbw[gn];
		while (need && !differs) {
			const uint32_t q = (uint32_t)__ffsll((long long)need) - 1u;
			need &= need - 1ull;
			differs = P.f.g_tab[(size_t)gn * 64 + q] != tab[q * B];
		}
		if (!differs) break;
		if (atomicExch(&P.f.claim[gn], claim_tag) == claim_tag) break;
		g = gn;
		j++;
		begin += QOI_FSEG_BYTES; /* the reader, pos and tb simply go on: this segment's first op is where the last one's ended */
		end = begin + QOI_FSEG_BYTES < stop ? begin + QOI_FSEG_BYTES : stop;
		P.f.g_px[g] = px;
		tab_store_256<B>(P.f.g_tab + (size_t)g * 64, tab);
	}
	}
}

/* round 0: lane k takes segment k; later rounds: the k-th entry of the list the last fverify_kernel left (late rounds are
   launched with a small grid and stride over it) */
__global__ void __launch_bounds__(QOI_FSTATE_THREADS) fstate_kernel(const QoiFParams P) {
	QOI_DYNAMIC_SMEM(smem);
	constexpr int B = QOI_FSTATE_THREADS;
	uint4 *lut = (uint4 *)smem;
	for (uint32_t k = threadIdx.x; k < 256u; k += B) lut[k] = qoidec::op_lut_entry(k);
	__syncthreads();
	uint32_t *tab = (uint32_t *)(smem + QOI_DEC_LUT_BYTES) + threadIdx.x; /* tab[slot * B] */
	uint32_t *flg = tab + 64 * B;                   /* flg[word * B]: one byte per slot */
	uint32_t *ring = flg + 16 * B;                  /* ring[word * B] */
	const uint32_t n = P.round == 0u ? P.n_segs : P.f.n_list[P.round];
	for (uint32_t k = blockIdx.x * B + threadIdx.x; k < n; k += gridDim.x * B) fstate_walk(P, P.round == 0u ? k : P.f.list[k], lut, tab, flg, ring);
}

/* ------------------------------------------------------------------ 3. last-writer scan over the segments of every stream */
/* Per stream and round: bad_cnt planes 0,1 = segments the check of an even / odd round could not prove, plane 2 = the stream
   has settled (a check proved every segment: exact by induction from segment 0), plane 3 = the last round that looked at
   the stream; first_bad planes 0,1 = the first unproven segment of that check.  A settled stream is skipped from then on,
   so extra rounds cost the slow streams' time only. */
__device__ __forceinline__ bool stream_settled(const QoiFParams &P, uint32_t img) {
	if (P.f.bad_cnt[2u * P.n_images + img] != 0u) return true;
	return P.round > 0u && P.f.bad_cnt[((P.round - 1u) & 1u) * P.n_images + img] == 0u;
}
/* Does segment g determine word t of the state behind it (t < 64: index[t], t == 64: the colour)?  What the segment wrote from
   its own bytes always does; what it derived from its guess only if that guess held (QOI_FF_TRUST) -- a stream of nothing but
   INDEX ops copies its guess around, and untrusted copies would push every wrong guess one segment further per round. */
/* A walk that started from a guess which held on everything it depended on is trusted; a fresh re-walk is trusted ahead of
   its check (QOI_FF_TRUST) -- unless it read an initial-zero slot (dep bit 6): then only once it lies inside the proven
   prefix of its stream (g < gpe; segment 0 always does). */
__device__ __forceinline__ bool seg_trusted(const QoiFParams &P, uint32_t g, uint32_t gpe) {
	const uint32_t fl = P.f.flag[g], d = P.f.dep[g];
	if ((fl & QOI_FF_TRUST) != 0u && (!(d & 64u) || g < gpe)) return true;
	/* Round 0: a walk that depended on nothing but a replayed colour (no guessed slot read, no initial zero) is as good as
	   one whose guess a check has confirmed -- its colour came from the stream's own bytes.  Trusting it lets the first
	   scan see its writes, so that the segments listed for round 1 get guesses that already hold (without it the same
	   segments were walked in round 1 AND round 2).  Sound: the check still compares its guess with the scanned state. */
	return P.round == 0u && (fl & QOI_FF_CGUESS) != 0u && P.f.rbw[g] == 0ull && !(d & 64u);
}
/* The colour is not "last writer": a segment of DIFF / LUMA / RUN ops adds a known delta to whatever colour enters it.  So
   the colour's scan composes maps  colour -> (r,g,b: constant c | + d) x (alpha: constant c | unchanged), one per segment.
   A map word: value bytes (constant or delta per channel), bit 32: r,g,b constant, bit 33: alpha constant. */
__device__ __forceinline__ uint64_t seg_colour_map(const QoiFParams &P, uint32_t g, uint32_t gpe) {
	const uint32_t d = P.f.dep[g], o = P.f.o_px[g], gi = P.f.g_px[g];
	const bool trust = seg_trusted(P, g, gpe);
	uint64_t m = 0;
	uint32_t v = 0;
	/* r,g,b: from the segment's own bytes, or derived through index[] and trusted -> constant; incoming + deltas -> add;
	   derived through index[] from a guess that did not hold -> unknown: pass the incoming colour on (delta 0) */
	if (!(d & 4u) || (!(d & 16u) && trust)) { m |= 1ull << 32; v |= o & 0x00FFFFFFu; }
	else if (d & 16u) v |= (((o | 0x80808080u) - (gi & 0x7F7F7F7Fu)) ^ ((o ^ ~gi) & 0x80808080u)) & 0x00FFFFFFu; /* o - guess, per byte */
	/* alpha: no op changes it relatively */
	if (!(d & 8u) || (!(d & 32u) && trust)) { m |= 1ull << 33; v |= o & 0xFF000000u; }
	return m | v;
}
__device__ __forceinline__ uint32_t colour_apply(uint64_t map, uint32_t px) {
	const uint32_t v = (uint32_t)map;
	const uint32_t rgb = (map >> 32) & 1ull ? (v & 0x00FFFFFFu) : (add_bytes(px, v & 0x00FFFFFFu) & 0x00FFFFFFu);
	const uint32_t a = (map >> 33) & 1ull ? (v & 0xFF000000u) : (px & 0xFF000000u);
	return rgb | a;
}
/* first `a`, then `b` */
__device__ __forceinline__ uint64_t colour_compose(uint64_t a, uint64_t b) {
	const uint32_t va = (uint32_t)a, vb = (uint32_t)b;
	uint64_t m = 0;
	uint32_t v = 0;
	if ((b >> 32) & 1ull) { m |= 1ull << 32; v |= vb & 0x00FFFFFFu; }
	else { m |= a & (1ull << 32); v |= add_bytes(va, vb & 0x00FFFFFFu) & 0x00FFFFFFu; }
	if ((b >> 33) & 1ull) { m |= 1ull << 33; v |= vb & 0xFF000000u; }
	else { m |= a & (1ull << 33); v |= va & 0xFF000000u; }
	return m | v;
}
/* What the walks over a chunk need to know about each of its (up to 64) segments, gathered once per chunk by one thread per
   segment: which slots the segment determines, its colour map, which words of its guess it depended on.  The walks
   themselves then fetch only index[] words (eight segments ahead of the chain of selects that consumes them). */
struct ChunkMeta {
	uint64_t gives[QOI_FCHUNK]; /* slots the segment determines (written, and from its own bytes or trusted) */
	uint64_t cmap[QOI_FCHUNK];  /* colour map */
	uint64_t rbw[QOI_FCHUNK];   /* guessed slots it depended on */
	uint32_t gpx[QOI_FCHUNK];   /* guessed colour */
	uint32_t need[QOI_FCHUNK];  /* ... and the bits of it the segment depended on */
	uint32_t stale[QOI_FCHUNK]; /* (fverify) 1: was not trusted when gathered */
};
#define QOI_FMETA_SMEM ((int)sizeof(qoifdec::ChunkMeta))
__device__ __forceinline__ void chunk_meta(const QoiFParams &P, ChunkMeta *M, uint32_t g0, uint32_t n, uint32_t gpe, uint32_t t) {
	if (t < n) {
		const uint32_t g = g0 + t;
		const bool trust = seg_trusted(P, g, gpe);
		const uint64_t wr = P.f.o_wr[g], ta = P.f.o_taint[g];
		const uint32_t d = P.f.dep[g];
		M->gives[t] = trust ? wr : (wr & ~ta);
		M->cmap[t] = seg_colour_map(P, g, gpe);
		M->rbw[t] = P.f.rbw[g];
		M->gpx[t] = P.f.g_px[g];
		M->need[t] = ((d & 1u) ? 0x00FFFFFFu : 0u) | ((d & 2u) ? 0xFF000000u : 0u);
		M->stale[t] = trust ? 0u : 1u;
	}
	__syncthreads();
}
#define QOI_FWALK_BATCH 8
/* every round after the first looks at the streams the check before it left unsettled; active[r] counts them */
__device__ __forceinline__ bool round_idle(const QoiFParams &P) { return P.round > 0u && P.f.active[P.round - 1u] == 0u; }

/* a: per chunk of QOI_FCHUNK segments, thread = slot (thread 64 = the colour): the chunk's last value per slot */
__device__ __forceinline__ void fscan_a_chunk(const QoiFParams &P, ChunkMeta *M, uint32_t c, uint32_t t) {
	const uint32_t img = chunk_image(P, c);
	if (stream_settled(P, img)) return;
	const uint32_t lc = c - P.f.chunk_base[img];
	const uint32_t live = P.f.n_live[img];
	const uint32_t j0 = lc * QOI_FCHUNK;
	if (j0 >= live) return; /* beyond the stream: never looked at */
	const uint32_t n = (j0 + QOI_FCHUNK < live ? j0 + QOI_FCHUNK : live) - j0;
	const uint32_t g0 = P.f.seg_base[img] + j0;
	chunk_meta(P, M, g0, n, stream_proven_end(P, img), t);
	if (t < 64u) {
		uint32_t val = 0;
		bool wrote = false;
		for (uint32_t jb = 0; jb < n; jb += QOI_FWALK_BATCH) {
			uint32_t v[QOI_FWALK_BATCH];
#pragma unroll
			for (int k = 0; k < QOI_FWALK_BATCH; k++)
				if (jb + k < n) v[k] = P.f.o_tab[(size_t)(g0 + jb + k) * 64 + t];
#pragma unroll
			for (int k = 0; k < QOI_FWALK_BATCH; k++)
				if (jb + k < n && ((M->gives[jb + k] >> t) & 1ull)) { val = v[k]; wrote = true; }
		}
		P.f.a_tab[(size_t)c * 64 + t] = val;
		const uint32_t b = __ballot_sync(QOI_FULL_MASK, wrote);
		if ((t & 31u) == 0u) ((uint32_t *)(P.f.a_wr + c))[t >> 5] = b; /* low / high half of the 64-bit mask */
	} else if (t == 64u) {
		uint64_t cm = 0; /* identity: + 0, alpha unchanged */
		for (uint32_t j = 0; j < n; j++) cm = colour_compose(cm, M->cmap[j]);
		P.f.a_px[c] = cm;
	}
	__syncthreads(); /* M is reused by the next chunk of this CTA */
}
__global__ void __launch_bounds__(QOI_FSCAN_THREADS) fscan_a_kernel(const QoiFParams P) {
	QOI_DYNAMIC_SMEM(smem); /* QOI_FMETA_SMEM */
	if (round_idle(P)) return;
	for (uint32_t c = blockIdx.x; c < P.n_chunks; c += gridDim.x) fscan_a_chunk(P, (ChunkMeta *)smem, c, threadIdx.x);
}

/* b: per stream: exclusive prefix over its chunks, from the reference's initial state (qoi.h:533-537).  A 4096x4096
   photograph has ~800 chunks: QOI_FSCANB_GROUPS groups of 64 slot threads each take a run of consecutive chunks (last value
   inside the run, then the groups before it, then the run again writing the prefixes); the last warp does the same for the
   colour maps with 32 runs. */
#define QOI_FSCANB_GROUPS 15
#define QOI_FSCANB_THREADS (QOI_FSCANB_GROUPS * 64 + 32)
#define QOI_FSCANB_SMEM ((QOI_FSCANB_GROUPS * 64 * 2 + 32 * 2 + 8) * 4)
__global__ void __launch_bounds__(QOI_FSCANB_THREADS) fscan_b_kernel(const QoiFParams P) {
	QOI_DYNAMIC_SMEM(smem);
	uint32_t *s_val = (uint32_t *)smem, *s_wr = s_val + QOI_FSCANB_GROUPS * 64;
	uint64_t *s_map = (uint64_t *)(s_wr + QOI_FSCANB_GROUPS * 64);
	const uint32_t img = blockIdx.x, t = threadIdx.x;
	if (img >= P.n_images || round_idle(P)) return;
	if (stream_settled(P, img)) { /* uniform over the CTA: every thread reads the same words */
		if (t == 0) P.f.bad_cnt[2u * P.n_images + img] = 1u;
		return;
	}
	if (t == 0) { /* what this round's check accumulates */
		atomicAdd(&P.f.active[P.round], 1u);
		P.f.bad_cnt[(P.round & 1u) * P.n_images + img] = 0u;
		P.f.first_bad[(P.round & 1u) * P.n_images + img] = 0x7FFFFFFFu;
		P.f.bad_cnt[3u * P.n_images + img] = P.round;
	}
	const uint32_t live = P.f.n_live[img];
	const uint32_t nch = (live + QOI_FCHUNK - 1) / QOI_FCHUNK;
	const uint32_t c0 = P.f.chunk_base[img];
	if (t < QOI_FSCANB_GROUPS * 64u) {
		const uint32_t sl = t & 63u, grp = t >> 6;
		const uint32_t per = (nch + QOI_FSCANB_GROUPS - 1) / QOI_FSCANB_GROUPS;
		const uint32_t lo = grp * per < nch ? grp * per : nch, hi = lo + per < nch ? lo + per : nch;
		uint32_t val = 0;
		bool wrote = false;
		for (uint32_t lc = lo; lc < hi; lc += 4) {
			uint32_t v[4];
			uint64_t m[4];
#pragma unroll
			for (int k = 0; k < 4; k++)
				if (lc + k < hi) { m[k] = P.f.a_wr[c0 + lc + k]; v[k] = P.f.a_tab[(size_t)(c0 + lc + k) * 64 + sl]; }
#pragma unroll
			for (int k = 0; k < 4; k++)
				if (lc + k < hi && ((m[k] >> sl) & 1ull)) { val = v[k]; wrote = true; }
		}
		s_val[t] = val;
		s_wr[t] = wrote ? 1u : 0u;
		__syncthreads();
		val = 0; /* index[] starts zeroed (qoi.h:533) */
		for (int q = (int)grp - 1; q >= 0; q--)
			if (s_wr[q * 64 + sl]) { val = s_val[q * 64 + sl]; break; }
		__syncthreads();
		for (uint32_t lc = lo; lc < hi; lc += 4) {
			uint32_t v[4];
			uint64_t m[4];
#pragma unroll
			for (int k = 0; k < 4; k++)
				if (lc + k < hi) { m[k] = P.f.a_wr[c0 + lc + k]; v[k] = P.f.a_tab[(size_t)(c0 + lc + k) * 64 + sl]; }
#pragma unroll
			for (int k = 0; k < 4; k++) {
				if (lc + k < hi) {
					P.f.p_tab[(size_t)(c0 + lc + k) * 64 + sl] = val;
					if ((m[k] >> sl) & 1ull) val = v[k];
				}
			}
		}
	} else {
		const uint32_t q = t - QOI_FSCANB_GROUPS * 64u; /* 0..31 */
		const uint32_t per = (nch + 31u) / 32u;
		const uint32_t lo = q * per < nch ? q * per : nch, hi = lo + per < nch ? lo + per : nch;
		uint64_t cm = 0;
		for (uint32_t lc = lo; lc < hi; lc++) cm = colour_compose(cm, P.f.a_px[c0 + lc]);
		s_map[q] = cm;
		__syncthreads();
		uint32_t px = 0xFF000000u; /* qoi.h:534-537 */
		for (uint32_t k = 0; k < q; k++) px = colour_apply(s_map[k], px);
		__syncthreads();
		for (uint32_t lc = lo; lc < hi; lc++) {
			P.f.p_px[c0 + lc] = px;
			px = colour_apply(P.f.a_px[c0 + lc], px);
		}
	}
}

/* c: every segment's guess against the scanned state, on what the segment depended on; a segment that disagrees gets the
   scanned state as its new guess and is listed for the next round's walk.  `last`: the check after the final round -- no
   new guesses (behind the proven prefix the guesses must stay what the results belong to: the fall-back decides per segment
   whether its last walk stands). */
#define QOI_FVERIFY_SMEM (QOI_FMETA_SMEM + (QOI_FCHUNK * 65 + 8) * 4)
__device__ __forceinline__ void fverify_chunk(const QoiFParams &P, ChunkMeta *M, uint32_t *s_T, uint32_t *s_dep, uint32_t c, uint32_t t, int last) {
	const uint32_t img = chunk_image(P, c);
	if (stream_settled(P, img)) return;
	const uint32_t lc = c - P.f.chunk_base[img];
	const uint32_t live = P.f.n_live[img];
	const uint32_t j0 = lc * QOI_FCHUNK;
	if (j0 >= live) return;
	const uint32_t n = (j0 + QOI_FCHUNK < live ? j0 + QOI_FCHUNK : live) - j0;
	const uint32_t g0 = P.f.seg_base[img] + j0;
	if (t < 2u) s_dep[t] = 0u;
	chunk_meta(P, M, g0, n, stream_proven_end(P, img), t);
	uint32_t T = t < 64u ? P.f.p_tab[(size_t)c * 64 + t] : (t == 64u ? P.f.p_px[c] : 0u);
	uint32_t mis_lo = 0, mis_hi = 0;
	if (t < 64u) {
		for (uint32_t jb = 0; jb < n; jb += QOI_FWALK_BATCH) {
			uint32_t v[QOI_FWALK_BATCH], gv[QOI_FWALK_BATCH];
#pragma unroll
			for (int k = 0; k < QOI_FWALK_BATCH; k++) {
				if (jb + k < n) {
					v[k] = P.f.o_tab[(size_t)(g0 + jb + k) * 64 + t];
					gv[k] = P.f.g_tab[(size_t)(g0 + jb + k) * 64 + t];
				}
			}
#pragma unroll
			for (int k = 0; k < QOI_FWALK_BATCH; k++) {
				if (jb + k < n) {
					const uint32_t j = jb + k;
					s_T[j * 65 + t] = T;
					if (((M->rbw[j] >> t) & 1ull) && gv[k] != T) {
						if (j < 32u) mis_lo |= 1u << j;
						else mis_hi |= 1u << (j - 32u);
					}
					if ((M->gives[j] >> t) & 1ull) T = v[k];
				}
			}
		}
	} else if (t == 64u) {
		for (uint32_t j = 0; j < n; j++) {
			s_T[j * 65 + 64] = T;
			if ((M->gpx[j] ^ T) & M->need[j]) {
				if (j < 32u) mis_lo |= 1u << j;
				else mis_hi |= 1u << (j - 32u);
			}
			T = colour_apply(M->cmap[j], T);
		}
	}
#pragma unroll
	for (int d = 16; d >= 1; d >>= 1) {
		mis_lo |= __shfl_xor_sync(QOI_FULL_MASK, mis_lo, d);
		mis_hi |= __shfl_xor_sync(QOI_FULL_MASK, mis_hi, d);
	}
	if ((t & 31u) == 0u) {
		if (mis_lo) atomicOr(&s_dep[0], mis_lo);
		if (mis_hi) atomicOr(&s_dep[1], mis_hi);
	}
	__syncthreads();
	const uint32_t dep_lo = s_dep[0], dep_hi = s_dep[1];
	if (t < 64u) { /* thread = segment of the chunk (two warps) */
		const bool in = t < n;
		const bool isdep = in && (((t < 32u ? dep_lo : dep_hi) >> (t & 31u)) & 1u);
		/* A segment that was not trusted when the state was scanned had its guess-derived results left out: agreeing with
		   such a state proves nothing yet.  The induction "segment 0 is exact; a segment whose guess agrees with the exact
		   state before it is exact" only runs over a prefix in which every segment was trusted AND agrees. */
		const bool was_trusted = in && M->stale[in ? t : 0u] == 0u;
		const bool bad = in && (isdep || !was_trusted);
		/* the next walk starts from the scanned state: trust what it derives from it until a check says otherwise */
		if (in) P.f.flag[g0 + t] = (uint8_t)((isdep && last) ? QOI_FF_DEP : QOI_FF_TRUST);
		const uint32_t par = P.round & 1u;
		const uint32_t bb = __ballot_sync(QOI_FULL_MASK, bad);
		/* the first unproven segment, and whether it merely was not trusted (then it is exact itself -- everything before it
		   is proven and it agrees -- and the prefix ends behind it): key = 2 * segment + that */
		const uint32_t fl = bb ? (uint32_t)__ffs((int)bb) - 1u : 0u;
		const uint32_t db = __shfl_sync(QOI_FULL_MASK, isdep ? 1u : 0u, fl);
		if (bb && (t & 31u) == 0u) {
			atomicAdd(&P.f.bad_cnt[par * P.n_images + img], (uint32_t)__popc(bb));
			atomicMin((int *)&P.f.first_bad[par * P.n_images + img], (int)(2u * (j0 + (t & 32u) + fl) + (db ? 0u : 1u)));
		}
		const uint32_t rb = __ballot_sync(QOI_FULL_MASK, isdep && !last);
		if (rb) {
			uint32_t at = 0;
			if ((t & 31u) == 0u) at = atomicAdd(&P.f.n_list[P.round + 1u], (uint32_t)__popc(rb));
			at = __shfl_sync(QOI_FULL_MASK, at, 0);
			if (isdep && !last) P.f.list[at + (uint32_t)__popc(rb & ((1u << (t & 31u)) - 1u))] = g0 + t;
		}
	}
	if (!last && t <= 64u) {
		for (uint32_t half = 0; half < 2u; half++) {
			uint32_t mk = half ? dep_hi : dep_lo;
			while (mk) {
				const uint32_t j = half * 32u + (uint32_t)__ffs((int)mk) - 1u;
				mk &= mk - 1u;
				if (t < 64u) P.f.g_tab[(size_t)(g0 + j) * 64 + t] = s_T[j * 65 + t];
				else P.f.g_px[g0 + j] = s_T[j * 65 + 64];
			}
		}
	}
	__syncthreads(); /* shared memory is reused by the next chunk of this CTA */
}
__global__ void __launch_bounds__(QOI_FSCAN_THREADS) fverify_kernel(const QoiFParams P, int last) {
	QOI_DYNAMIC_SMEM(smem); /* QOI_FVERIFY_SMEM */
	if (round_idle(P)) return;
	ChunkMeta *M = (ChunkMeta *)smem;
	uint32_t *s_T = (uint32_t *)(smem + QOI_FMETA_SMEM); /* [segment of the chunk][65]: the scanned state before it */
	uint32_t *s_dep = s_T + QOI_FCHUNK * 65;             /* [2]: segments with a disagreement (64-bit mask) */
	for (uint32_t c = blockIdx.x; c < P.n_chunks; c += gridDim.x) fverify_chunk(P, M, s_T, s_dep, c, threadIdx.x, last);
}

/* d: after the rounds.  Every segment of the proven prefix gets the scanned state in all 64 slots -- the exact state before
   the segment, which fcompose_kernel needs for the slots a restart record's segment had not written -- and the fall-back
   walk its start state.  (A settled stream's scan results are still in place: nothing touched them since.) */
__device__ __forceinline__ uint32_t stream_first_bad(const QoiFParams &P, uint32_t img) {
	const uint32_t par = P.f.bad_cnt[3u * P.n_images + img] & 1u;
	const uint32_t fb = P.f.first_bad[par * P.n_images + img] >> 1, live = P.f.n_live[img]; /* (key, see fverify_kernel) */
	return fb < live ? fb : live;
}
__global__ void __launch_bounds__(QOI_FSCAN_THREADS) ffinal_kernel(const QoiFParams P) {
	QOI_DYNAMIC_SMEM(smem); /* QOI_FMETA_SMEM */
	ChunkMeta *M = (ChunkMeta *)smem;
	const uint32_t c = blockIdx.x, t = threadIdx.x;
	if (c >= P.n_chunks) return;
	const uint32_t img = chunk_image(P, c);
	const uint32_t lc = c - P.f.chunk_base[img];
	const uint32_t live = P.f.n_live[img];
	const uint32_t j0 = lc * QOI_FCHUNK;
	if (j0 >= live) return;
	const uint32_t fb = stream_first_bad(P, img);
	if (j0 > fb) return; /* nothing proven in this chunk, and the fall-back does not start here */
	const uint32_t n = (j0 + QOI_FCHUNK < live ? j0 + QOI_FCHUNK : live) - j0;
	const uint32_t g0 = P.f.seg_base[img] + j0;
	chunk_meta(P, M, g0, n, P.f.seg_base[img] + (fb > 1u ? fb : 1u), t); /* every segment before fb is proven */
	uint32_t T = t < 64u ? P.f.p_tab[(size_t)c * 64 + t] : (t == 64u ? P.f.p_px[c] : 0u);
	if (t < 64u) {
		for (uint32_t jb = 0; jb < n; jb += QOI_FWALK_BATCH) {
			uint32_t v[QOI_FWALK_BATCH];
#pragma unroll
			for (int k = 0; k < QOI_FWALK_BATCH; k++)
				if (jb + k < n) v[k] = P.f.o_tab[(size_t)(g0 + jb + k) * 64 + t];
#pragma unroll
			for (int k = 0; k < QOI_FWALK_BATCH; k++) {
				if (jb + k < n) {
					const uint32_t jj = j0 + jb + k, g = g0 + jb + k;
					if (jj < fb) P.f.g_tab[(size_t)g * 64 + t] = T;
					else if (jj == fb) P.f.x_state[(size_t)img * 65 + t] = T; /* where the fall-back walk starts */
					if ((M->gives[jb + k] >> t) & 1ull) T = v[k];
				}
			}
		}
	} else if (t == 64u) {
		for (uint32_t j = 0; j < n; j++) {
			const uint32_t jj = j0 + j;
			if (jj < fb) P.f.g_px[g0 + j] = T;
			else if (jj == fb) P.f.x_state[(size_t)img * 65 + 64] = T;
			T = colour_apply(M->cmap[j], T);
		}
	}
}

/* ------------------------------------------------------------------ 4. fall-back: one exact walk per unsettled stream */
/* From the first segment the last check could not prove (ffinal_kernel left the exact state before it in x_state:
   everything before it is proven) to the end of the stream, with the exact state in shared memory.  A segment whose last walk started
   from a guess that agrees with the exact state on everything it depended on is taken over as it is (its results advance
   the state; it gets the exact state as its guess, for fcompose_kernel); any other segment is decoded on the spot: lane 0
   runs the reference loop op by op on the segment's bytes, which the warp stages in shared memory, and the restart records
   are dropped with the full exact index[]. */
__global__ void __launch_bounds__(32) ffall_kernel(const QoiFParams P) {
	QOI_DYNAMIC_SMEM(smem);
	uint32_t *tab = (uint32_t *)smem; /* exact index[] */
	uint32_t *stage = tab + 64;       /* one segment of stream words (QOI_FFALL_SMEM) */
	const uint32_t img = blockIdx.x;
	const uint32_t lane = threadIdx.x;
	if (img >= P.n_images) return;
	const uint32_t live = P.f.n_live[img];
	const uint32_t fb = stream_first_bad(P, img);
	if (fb >= live) return; /* everything proven (also: rejected or empty stream) */
	const QoiImgMeta m = P.imgs[img];
	const uint8_t *s = P.streams + m.stream_off;
	const int32_t stop = P.lens[img] - QOI_END_BYTES;
	const uintptr_t s_lo = (uintptr_t)s & ~(uintptr_t)3, s_hi = ((uintptr_t)s + (uintptr_t)P.lens[img] + 3u) & ~(uintptr_t)3;
	const uint32_t base = P.f.seg_base[img];
	tab[lane] = P.f.x_state[(size_t)img * 65 + lane];
	tab[lane + 32] = P.f.x_state[(size_t)img * 65 + lane + 32];
	uint32_t px = P.f.x_state[(size_t)img * 65 + 64];
	__syncwarp();
	for (uint32_t j = fb; j < live; j++) {
		const uint32_t g = base + j;
		/* does the segment's last walk stand?  (its guess against the exact state, on what it depended on) */
		const uint64_t rbw = P.f.rbw[g];
		const uint32_t dep = P.f.dep[g], g_px = P.f.g_px[g];
		const uint32_t g0 = P.f.g_tab[(size_t)g * 64 + lane], g1 = P.f.g_tab[(size_t)g * 64 + lane + 32];
		const uint32_t pmask = ((dep & 1u) ? 0x00FFFFFFu : 0u) | ((dep & 2u) ? 0xFF000000u : 0u);
		bool bad = (((rbw >> lane) & 1ull) && g0 != tab[lane]) || (((rbw >> (lane + 32)) & 1ull) && g1 != tab[lane + 32]);
		bad = bad || (lane == 0 && ((g_px ^ px) & pmask) != 0u);
		/* the exact state before the segment becomes its guess: the records it dropped are completed from it */
		P.f.g_tab[(size_t)g * 64 + lane] = tab[lane];
		P.f.g_tab[(size_t)g * 64 + lane + 32] = tab[lane + 32];
		if (lane == 0) P.f.g_px[g] = px;
		if (__ballot_sync(QOI_FULL_MASK, bad) == 0u) {
			const uint64_t wr = P.f.o_wr[g];
			if ((wr >> lane) & 1ull) tab[lane] = P.f.o_tab[(size_t)g * 64 + lane];
			if ((wr >> (lane + 32)) & 1ull) tab[lane + 32] = P.f.o_tab[(size_t)g * 64 + lane + 32];
			px = P.f.o_px[g];
			__syncwarp();
			continue;
		}
		const int32_t begin = QOI_HEADER_BYTES + (int32_t)j * QOI_FSEG_BYTES;
		const int32_t end = begin + QOI_FSEG_BYTES < stop ? begin + QOI_FSEG_BYTES : stop;
		/* all lanes copy the segment's words (and the few bytes an op may reach past `end`) into shared memory; words outside
		   the stream are never touched */
		const uintptr_t w0 = (uintptr_t)(s + begin) & ~(uintptr_t)3;
		const uintptr_t w1 = ((uintptr_t)(s + end) + 8u + 3u) & ~(uintptr_t)3;
		const uint32_t nw = (uint32_t)((w1 - w0) >> 2);
		for (uint32_t k = lane; k < nw; k += 32) {
			const uintptr_t wa = w0 + 4u * k;
			stage[k] = (wa >= s_lo && wa + 4 <= s_hi) ? __ldg((const uint32_t *)wa) : 0u;
		}
		__syncwarp();
		uint32_t pos = P.f.px_start[g], tb = P.f.tb0[g];
		int32_t p = begin + (int32_t)P.f.entry[g];
		/* lane 0 decodes until the segment is used up or an op starts behind a tile boundary (then the warp writes the record) */
		bool more = true;
		while (more) {
			uint32_t ev = 0; /* 1: restart record due at (pos, p, px) */
			if (lane == 0) {
				const uint8_t *stage_b = (const uint8_t *)stage;
				const int32_t d = (int32_t)((intptr_t)(uintptr_t)s - (intptr_t)w0); /* stream byte q sits at stage_b[q + d] */
				while (p < end && pos < m.npx) {
					if (pos >= tb) { ev = 1; break; }
					const uint32_t tag = stage_b[(p++) + d];
					uint32_t npix = 1;
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
						npix = (tag & 0x3Fu) + 1u;
					}
					tab[qoi_slot(px)] = px;
					pos += npix;
				}
			}
			ev = __shfl_sync(QOI_FULL_MASK, ev, 0);
			p = __shfl_sync(QOI_FULL_MASK, p, 0);
			pos = __shfl_sync(QOI_FULL_MASK, pos, 0);
			px = __shfl_sync(QOI_FULL_MASK, px, 0);
			if (ev) {
				const uint32_t T = m.first_tile + tb / QOI_TILE_PX;
				if (lane == 0) {
					P.st.rs_px[T] = (int32_t)pos;
					P.st.rs_byte[T] = p;
					P.st.rs_prev[T] = px;
					P.f.rec_wr[T] = ~0ull; /* exact in every slot */
					P.f.rs_seg[T] = g;
				}
				if (tb) {
					P.st.ent[(size_t)(T - 1) * 64 + lane] = ((uint64_t)tab[lane] << 32) | 1ull;
					P.st.ent[(size_t)(T - 1) * 64 + lane + 32] = ((uint64_t)tab[lane + 32] << 32) | 1ull;
				}
				tb += QOI_TILE_PX;
			} else {
				more = false;
			}
			__syncwarp();
		}
	}
}

/* ------------------------------------------------------------------ 5. complete the restart records */
/* index[] of a record, in the slots its segment had not written yet when it dropped it, is the exact state before that
   segment (the segment's final guess).  One thread per (tile, slot). */
__global__ void __launch_bounds__(128) fcompose_kernel(const QoiFParams P, uint32_t n_tiles, const uint32_t *tile_img) {
	const uint32_t k = blockIdx.x * 128 + threadIdx.x;
	const uint32_t T = k >> 6, sl = k & 63u;
	if (T >= n_tiles) return;
	const QoiImgMeta m = P.imgs[tile_img[T]];
	if (T == m.first_tile || P.st.rs_px[T] < 0) return; /* tile 0 starts from zeros; no op starts in this tile */
	if ((P.f.rec_wr[T] >> sl) & 1ull) return;
	P.st.ent[(size_t)(T - 1) * 64 + sl] = ((uint64_t)P.f.g_tab[(size_t)P.f.rs_seg[T] * 64 + sl] << 32) | 1ull;
}

/* binds the arrays of a QoiFState onto one allocation laid out by qoi_fstate_layout() (qoi_plan.h) */
static inline void fstate_bind(QoiFState &f, unsigned char *base, const size_t *off) {
	f.seg_base = (const uint32_t *)(base + off[0]);
	f.chunk_base = (const uint32_t *)(base + off[1]);
	f.px_cnt = (uint32_t *)(base + off[2]);
	f.exit_skip = (uint8_t *)(base + off[3]);
	f.entry = (uint8_t *)(base + off[4]);
	f.px_start = (uint32_t *)(base + off[5]);
	f.n_live = (uint32_t *)(base + off[6]);
	f.g_px = (uint32_t *)(base + off[7]);
	f.g_tab = (uint32_t *)(base + off[8]);
	f.o_px = (uint32_t *)(base + off[9]);
	f.o_tab = (uint32_t *)(base + off[10]);
	f.o_wr = (uint64_t *)(base + off[11]);
	f.rbw = (uint64_t *)(base + off[12]);
	f.dep = (uint32_t *)(base + off[13]);
	f.flag = (uint8_t *)(base + off[14]);
	f.list = (uint32_t *)(base + off[15]);
	f.n_list = (uint32_t *)(base + off[16]);
	f.a_tab = (uint32_t *)(base + off[17]);
	f.a_wr = (uint64_t *)(base + off[18]);
	f.a_px = (uint64_t *)(base + off[19]);
	f.p_tab = (uint32_t *)(base + off[20]);
	f.p_px = (uint32_t *)(base + off[21]);
	f.claim = (uint32_t *)(base + off[30]);
	f.active = (uint32_t *)(base + off[31]);
	f.bad_cnt = (uint32_t *)(base + off[22]);
	f.first_bad = (uint32_t *)(base + off[23]);
	f.o_taint = (uint64_t *)(base + off[24]);
	f.last_rel = (uint32_t *)(base + off[25]);
	f.tb0 = (uint32_t *)(base + off[26]);
	f.rec_wr = (uint64_t *)(base + off[27]);
	f.rs_seg = (uint32_t *)(base + off[28]);
	f.x_state = (uint32_t *)(base + off[29]);
	f.last_abs = (uint16_t *)(base + off[32]);
}

}  // namespace qoifdec
