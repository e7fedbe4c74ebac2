/*
 * qoi_dec.cuh -- segment-parallel QOI decoder for sm_100a.
 *
 * Replaces the reference decode loop (qoi.h:540-587).  The op stream of one
 * image is inherently sequential (p, px, run and index[64] all chain), so the
 * unit of parallelism is a *segment*: a position in the stream at which the
 * complete decoder state is known.
 *
 *   foreign stream: no state is known inside the stream; those go through the
 *       speculative decoder of qoi_fdec.cuh.  (use_index = 0 in THIS kernel means one
 *       segment = the whole image from the reference's initial state, qoi.h:533-537;
 *       it is kept as the simplest exact form and is not launched by the C ABI.)
 *   stream written by qoi_encode_kernel (use_index = 1): one segment per
 *       encoder tile.  The chained scan already materialised the state at every
 *       tile boundary: index[] after tile t-1 is ent[t-1], and rs_px /
 *       rs_byte / rs_prev give the first op that starts inside tile t, where it
 *       lies in the stream and the pixel before it.  The index is a side table
 *       in HBM; the stream itself carries nothing extra and stays bit exact.
 *
 * One LANE per segment (not one warp): each lane runs the scalar state machine
 * with its 64-entry table in shared memory, laid out [slot][thread] so that any
 * slot pattern across the lanes is bank-conflict free.  Stream bytes are fetched
 * 32 bytes (one full sector) at a time per lane with two 16-byte loads that are
 * issued one chunk ahead and parked in registers, then dropped into a small
 * per-lane ring in shared memory; the op parser reads a 64-bit window refilled
 * from that ring.  No load touches a byte outside the stream's own words.  All
 * lanes produce one pixel per loop trip, so pixels leave in groups of four with
 * one 16-byte store per lane.
 */
#pragma once
#include "qoi_common.h"
#include "qoi_dev.cuh"

namespace qoidec {
using namespace qoidev;

#define QOI_DEC_THREADS 128 /* default block; a 32-wide variant spreads small batches over the SMs */

__device__ __forceinline__ uint32_t be32(const uint8_t *p) {
	return ((uint32_t)p[0] << 24) | ((uint32_t)p[1] << 16) | ((uint32_t)p[2] << 8) | (uint32_t)p[3];
}

/* per-byte add mod 256 of two packed words (no carries between bytes) */
__device__ __forceinline__ uint32_t add_bytes(uint32_t a, uint32_t b) {
	return ((a & 0x7F7F7F7Fu) + (b & 0x7F7F7F7Fu)) ^ ((a ^ b) & 0x80808080u);
}

#define QOI_DEC_RING_WORDS 16 /* two 32-byte chunks per lane */
#define QOI_DEC_SMEM_WORDS (64 + QOI_DEC_RING_WORDS)

struct Reader {
	uint64_t buf;       /* the next `avail` stream bytes, lowest byte first */
	uint32_t avail;
	int32_t p;          /* stream offset of buf's first byte */
	uint32_t wi;        /* ring word the window is refilled from next */
	const uint4 *gp;    /* next 32-byte chunk to fetch (32-byte aligned address) */
	uintptr_t lo, hi;   /* the stream's own words: [lo, hi), 4-byte aligned */
	uint4 pa, pb;       /* the chunk that is in flight / parked until its ring half is free */
};

/* 32 bytes at gp; words outside [lo, hi) read as zero and are never touched */
__device__ __forceinline__ void load_chunk(Reader &rd) {
	const uintptr_t a = (uintptr_t)rd.gp;
	if (a >= rd.lo && a + 32 <= rd.hi) {
		rd.pa = __ldg(rd.gp);
		rd.pb = __ldg(rd.gp + 1);
	} else {
		uint32_t w[8];
#pragma unroll
		for (int k = 0; k < 8; k++) {
			const uintptr_t wa = a + 4u * k;
			w[k] = (wa >= rd.lo && wa + 4 <= rd.hi) ? __ldg((const uint32_t *)wa) : 0u;
		}
		rd.pa = make_uint4(w[0], w[1], w[2], w[3]);
		rd.pb = make_uint4(w[4], w[5], w[6], w[7]);
	}
	rd.gp += 2;
}

template <int OUT_CH, int BLOCK, bool UNI, bool PRE>
__global__ void __launch_bounds__(BLOCK) qoi_decode_kernel(const QoiDecParams P) {
	QOI_DYNAMIC_SMEM(smem);
	uint32_t *tab = (uint32_t *)smem + threadIdx.x;  /* tab[slot * BLOCK] */
	uint32_t *ring = tab + 64 * BLOCK;               /* ring[word * BLOCK] */
	const uint32_t job_id = blockIdx.x * BLOCK + threadIdx.x;
	if (job_id >= P.n_jobs) return;
	const uint32_t job = job_id; /* the tile (indexed decode) or the image */

	/* ---- which image, which pixels ---- */
	const uint32_t img = P.use_index ? P.tile_img[job] : job;
	const QoiImgMeta m = P.imgs[img];
	const uint8_t *s = P.streams + m.stream_off;
	const int32_t size = P.lens[img];
	const bool first_seg = !P.use_index || job == m.first_tile;

	/* header against the plan (the host entry points reject bad headers before any launch, qoi.h:513-521) */
	bool ok = size >= QOI_HEADER_BYTES + QOI_END_BYTES;
	if (ok) {
		ok = s[0] == 'q' && s[1] == 'o' && s[2] == 'i' && s[3] == 'f' && be32(s + 4) == m.width && be32(s + 8) == m.height &&
		     s[12] == m.channels && s[13] == m.colorspace;
	}
	if (first_seg && P.status) P.status[img] = ok ? 0 : 4; /* QOI_GPU_E_STREAM */
	if (!ok) return;

	uint32_t px_begin = 0, px_end = m.npx;
	int32_t p0 = QOI_HEADER_BYTES;
	uint32_t px = 0xFF000000u; /* {0,0,0,255}, qoi.h:534-537 */
	if (P.use_index) {
		const int32_t b = P.st.rs_px[job];
		if (b < 0) return; /* the image ends before an op starts in this tile */
		px_begin = (uint32_t)b;
		p0 = P.st.rs_byte[job];
		px = P.st.rs_prev[job];
		const uint32_t tiles = (m.npx + QOI_TILE_PX - 1) / QOI_TILE_PX;
		if (job + 1 < m.first_tile + tiles) {
			const int32_t e = P.st.rs_px[job + 1];
			if (e >= 0) px_end = (uint32_t)e;
		}
	}
	/* ---- index[] at the segment start ---- */
	if (first_seg) {
#pragma unroll 8
		for (int k = 0; k < 64; k++) tab[k * BLOCK] = 0;
	} else {
		const uint64_t *src = P.st.ent + (size_t)(job - 1) * 64; /* [63:32] value; 0 in a slot nothing has been stored to yet */
#pragma unroll 8
		for (int k = 0; k < 64; k++) tab[k * BLOCK] = (uint32_t)(src[k] >> 32);
		/* the reference DEcoder also stores the current colour after a RUN op (qoi.h:577); the ENcoder's
		   table does not see the leading run of the start colour.  Only slot 53 can differ, and only
		   until a pixel hashing to 53 is coded. */
		const uint32_t op0 = s[QOI_HEADER_BYTES];
		if ((uint32_t)(src[53] >> 32) == 0u && (op0 & 0xC0u) == 0xC0u && op0 < 0xFEu) tab[53 * BLOCK] = 0xFF000000u; /* no colour of slot 53 is 0 */
	}

	const int32_t stop = size - QOI_END_BYTES; /* qoi.h:539 */
	Reader rd;
	{
		const uintptr_t a = (uintptr_t)(s + p0);
		rd.lo = (uintptr_t)s & ~(uintptr_t)3;
		rd.hi = ((uintptr_t)s + (uintptr_t)size + 3u) & ~(uintptr_t)3;
		rd.gp = (const uint4 *)(a & ~(uintptr_t)31);
		rd.p = p0;
		/* the first two chunks go straight into the ring, the third is put in flight */
#pragma unroll
		for (int half = 0; half < 2; half++) {
			load_chunk(rd);
			uint32_t *r = ring + half * 8 * BLOCK;
			r[0 * BLOCK] = rd.pa.x; r[1 * BLOCK] = rd.pa.y; r[2 * BLOCK] = rd.pa.z; r[3 * BLOCK] = rd.pa.w;
			r[4 * BLOCK] = rd.pb.x; r[5 * BLOCK] = rd.pb.y; r[6 * BLOCK] = rd.pb.z; r[7 * BLOCK] = rd.pb.w;
		}
		load_chunk(rd);
		rd.wi = (uint32_t)((a & 31u) >> 2);
		rd.buf = 0;
		rd.avail = 0;
	}
	/* next 4 stream bytes from the ring; when a ring half has been used up, park the in-flight chunk there and fetch the next */
	/* PRE (experimental, needs UNI; off by default, not yet measured): the ring word is read one call ahead, so that the
	   shared-memory latency is off the parse chain.  The word after the current one is always in the ring: the chunk
	   behind the reader's is parked by the refill() of the trip in which the reader entered its chunk. */
	uint32_t ahead = 0;
	auto next_word = [&]() -> uint32_t {
		const uint32_t w = PRE ? ahead : ring[(rd.wi & (QOI_DEC_RING_WORDS - 1)) * BLOCK];
		rd.wi++;
		if (PRE) ahead = ring[(rd.wi & (QOI_DEC_RING_WORDS - 1)) * BLOCK];
		if (!UNI && (rd.wi & 7u) == 0u) { /* lane-driven refill: the half just read is free */
			uint32_t *r = ring + ((rd.wi - 8u) & (QOI_DEC_RING_WORDS - 1)) * BLOCK;
			r[0 * BLOCK] = rd.pa.x; r[1 * BLOCK] = rd.pa.y; r[2 * BLOCK] = rd.pa.z; r[3 * BLOCK] = rd.pa.w;
			r[4 * BLOCK] = rd.pb.x; r[5 * BLOCK] = rd.pb.y; r[6 * BLOCK] = rd.pb.z; r[7 * BLOCK] = rd.pb.w;
			load_chunk(rd);
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
			uint32_t *r = ring + ((filled & 1u) * 8u) * BLOCK;
			r[0 * BLOCK] = rd.pa.x; r[1 * BLOCK] = rd.pa.y; r[2 * BLOCK] = rd.pa.z; r[3 * BLOCK] = rd.pa.w;
			r[4 * BLOCK] = rd.pb.x; r[5 * BLOCK] = rd.pb.y; r[6 * BLOCK] = rd.pb.z; r[7 * BLOCK] = rd.pb.w;
			filled++;
			load_chunk(rd);
		}
	};
	{ /* prime the window with the word that holds byte p0 */
		if (PRE) ahead = ring[(rd.wi & (QOI_DEC_RING_WORDS - 1)) * BLOCK];
		const uint32_t mis = (uint32_t)((uintptr_t)(s + p0) & 3u);
		rd.buf = (uint64_t)(next_word() >> (8u * mis));
		rd.avail = 4u - mis;
	}
	uint32_t run = 0;

	/* one pixel: qoi.h:541-578.  Lanes of a warp sit in different op types all the time, so the op itself is
	   evaluated branch-free (every candidate result is computed, selects pick one); only the refill branches. */
	auto step = [&]() -> uint32_t {
		const bool parse = (run == 0u) && (rd.p < stop);
		if (parse && rd.avail < 5u) { /* an op needs at most 5 bytes */
			rd.buf |= (uint64_t)next_word() << (8u * rd.avail);
			rd.avail += 4u;
			if (rd.avail < 5u) {
				rd.buf |= (uint64_t)next_word() << (8u * rd.avail);
				rd.avail += 4u;
			}
		}
		const uint32_t tag = (uint32_t)rd.buf & 0xFFu;
		const uint32_t pay = (uint32_t)(rd.buf >> 8); /* the four bytes after the tag */
		const uint32_t top = tag >> 6;
		const bool is_rgb = tag == 0xFEu, is_rgba = tag == 0xFFu; /* qoi.h:547-557 are tested before the 2-bit tags */
		const uint32_t idxv = tab[(tag & 63u) * BLOCK];           /* qoi.h:558-560 */
		/* qoi.h:561-565: each channel += field - 2, wrapping (LUT 0,1,2,3 -> -2,-1,0,+1) */
		const uint32_t sel = ((tag >> 4) & 3u) | (((tag >> 2) & 3u) << 4) | ((tag & 3u) << 8) | 0x4000u;
		const uint32_t dadd = __byte_perm(0x0100FFFEu, 0u, sel);
		/* qoi.h:566-572 */
		const uint32_t vg = (tag & 0x3Fu) + 224u; /* -32 mod 256 */
		const uint32_t rb = vg + 248u;            /* vg - 8 */
		const uint32_t ladd = ((rb + ((pay >> 4) & 0xFu)) & 0xFFu) | ((vg & 0xFFu) << 8) | (((rb + (pay & 0xFu)) & 0xFFu) << 16);
		const uint32_t rel = add_bytes(px, top == 1u ? dadd : ladd);
		uint32_t npx = top == 0u ? idxv : rel;
		npx = top == 3u ? px : npx;                                            /* RUN keeps the colour (qoi.h:573-575) */
		npx = is_rgb ? ((px & 0xFF000000u) | (pay & 0x00FFFFFFu)) : npx;
		npx = is_rgba ? pay : npx;
		const uint32_t len = is_rgba ? 5u : (is_rgb ? 4u : (top == 2u ? 2u : 1u));
		const uint32_t nrun = (top == 3u && tag < 0xFEu) ? (tag & 0x3Fu) : 0u;
		if (parse) {
			px = npx;
			rd.buf >>= 8u * len;
			rd.avail -= len;
			rd.p += (int32_t)len;
			tab[qoi_slot(px) * BLOCK] = px; /* qoi.h:577 */
			run = nrun;
		} else if (run > 0u) {
			run--;
		}
		return px;
	};

	uint8_t *obase = P.out + m.dec_off;
	uint32_t i = px_begin;
	if (OUT_CH == 4) {
		uint32_t *o32 = (uint32_t *)obase;
		const bool vec_ok = ((uintptr_t)obase & 15u) == 0;
		while (i < px_end && (!vec_ok || (i & 3u))) { /* head (or everything, if the destination is not 16-byte aligned) */
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
		const bool vec_ok = ((uintptr_t)obase & 3u) == 0;
		while (i < px_end && (!vec_ok || (i & 3u))) {
			const uint32_t a = step();
			uint8_t *o = obase + (size_t)i * 3;
			o[0] = (uint8_t)a; o[1] = (uint8_t)(a >> 8); o[2] = (uint8_t)(a >> 16);
			i++;
			refill();
		}
		while (i + 4 <= px_end) { /* 4 pixels = 12 bytes = three aligned words */
			const uint32_t a = step(), b = step(), c = step(), d = step();
			uint32_t *o = (uint32_t *)(obase + (size_t)i * 3);
			o[0] = (a & 0x00FFFFFFu) | (b << 24);
			o[1] = ((b >> 8) & 0x0000FFFFu) | (c << 16);
			o[2] = ((c >> 16) & 0x000000FFu) | (d << 8);
			i += 4;
			refill();
		}
		while (i < px_end) {
			const uint32_t a = step();
			uint8_t *o = obase + (size_t)i * 3;
			o[0] = (uint8_t)a; o[1] = (uint8_t)(a >> 8); o[2] = (uint8_t)(a >> 16);
			i++;
			refill();
		}
	}
}

}  // namespace qoidec
