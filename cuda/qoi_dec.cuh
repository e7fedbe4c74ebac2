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
#define QOI_DEC_LUT_BYTES 4096 /* 256 x uint4: what an op's first byte says, see op_lut_entry() */
#define QOI_DEC_SMEM_BYTES(BLOCK) (QOI_DEC_LUT_BYTES + (64 + QOI_DEC_RING_WORDS) * (BLOCK) * 4)

/*
 * Everything an op's first byte determines (qoi.h:547-575), tabulated once per block: with the fields picked apart by
 * shifts and masks the decoders spent 64 ALU instructions per pixel (ALU pipe at 70 % of its peak, profiles/r2_enc_v2);
 * one 16-byte shared-memory load replaces ~30 of them.  What bounds the kernel after that is the L1 data pipe (one pass
 * per distinct line a warp-wide access touches; shared-memory passes count too): 70 % busy whatever the instruction count
 * (profiles/r2_dec_lut, r2_dec_256).  4- and 8-byte tables with rebuilt masks, and a software-pipelined op fetch, were
 * measured and are not here: they move the load between the ALU and that pipe without shortening the kernel; 256-bit
 * global accesses (ld_nc_256 / st_256) halve the passes and took the photo batch from 2.45 to 1.94 ms.
 *   x  r,b delta in 16-bit fields: r | b << 16 (DIFF: field - 2; LUMA: vg - 8, the nibbles of the second byte are added)
 *   y  byte 0: g delta (DIFF: field - 2; LUMA: vg), byte 1: op length in bytes, byte 2: pixels of a RUN op AFTER its first
 *   z  0x000F000F for LUMA (the second byte's nibbles count), else 0
 *   w  bits the new colour takes straight from the four bytes behind the tag: RGB 0x00FFFFFF, RGBA 0xFFFFFFFF, else 0
 * A relative op is then  r,b: (px & 0x00FF00FF) + x + nibbles,  g: ((px >> 8) & 0x00FF00FF) + g delta  (16-bit fields: no
 * carry crosses a channel, the low byte of each field is the sum mod 256, qoi.h:563-571) and one PRMT picks the four low
 * bytes; RUN and INDEX have zero deltas (INDEX takes index[] instead, selected afterwards).
 */
__device__ __forceinline__ uint4 op_lut_entry(uint32_t tag) {
	uint4 e = make_uint4(0u, 0u, 0u, 0u);
	const uint32_t top = tag >> 6;
	uint32_t len = 1u, run = 0u, dg = 0u;
	if (tag == 0xFEu) { len = 4u; e.w = 0x00FFFFFFu; }
	else if (tag == 0xFFu) { len = 5u; e.w = 0xFFFFFFFFu; }
	else if (top == 1u) { /* QOI_OP_DIFF, qoi.h:561-565 */
		const uint32_t dr = (((tag >> 4) & 3u) - 2u) & 0xFFu, db = ((tag & 3u) - 2u) & 0xFFu;
		dg = (((tag >> 2) & 3u) - 2u) & 0xFFu;
		e.x = dr | (db << 16);
	} else if (top == 2u) { /* QOI_OP_LUMA, qoi.h:566-572 */
		const uint32_t vg = (tag & 0x3Fu) - 32u;
		const uint32_t rb = (vg - 8u) & 0xFFu;
		dg = vg & 0xFFu;
		e.x = rb | (rb << 16);
		e.z = 0x000F000Fu;
		len = 2u;
	} else if (top == 3u) { /* QOI_OP_RUN, qoi.h:573-575: run + 1 pixels */
		run = tag & 0x3Fu;
	}
	e.y = dg | (len << 8) | (run << 16);
	return e;
}

/* the new colour of the op `tag` with the four bytes `pay` behind it, index[tag & 63] = idxv, colour before = px */
__device__ __forceinline__ uint32_t op_colour(uint32_t tag, uint32_t pay, uint4 L, uint32_t idxv, uint32_t px) {
	const uint32_t nib = ((((pay & 0xFFu) * 0x00100001u) >> 4) & L.z); /* LUMA: dr_dg at bits 0..3, db_dg at bits 16..19 */
	const uint32_t E = __byte_perm(px, 0u, 0x4240u) + L.x + nib;       /* r, b in 16-bit fields */
	const uint32_t O = __byte_perm(px, 0u, 0x4341u) + __dp4a(L.y, 0x00000001u, 0u); /* g, a */
	uint32_t npx = __byte_perm(E, O, 0x6240u);                          /* low bytes of the four fields */
	npx = (pay & L.w) | (npx & ~L.w);                                   /* RGB keeps alpha, RGBA replaces all (qoi.h:547-557) */
	return tag < 0x40u ? idxv : npx;                                    /* QOI_OP_INDEX, qoi.h:558-560 */
}

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
		ld_nc_256(rd.gp, rd.pa, rd.pb);
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

template <int OUT_CH, int BLOCK>
__global__ void __launch_bounds__(BLOCK) qoi_decode_kernel(const QoiDecParams P) {
	QOI_DYNAMIC_SMEM(smem);
	uint4 *lut = (uint4 *)smem;
	for (uint32_t k = threadIdx.x; k < 256u; k += BLOCK) lut[k] = op_lut_entry(k);
	__syncthreads(); /* before any lane leaves */
	uint32_t *tab = (uint32_t *)(smem + QOI_DEC_LUT_BYTES) + threadIdx.x;  /* tab[slot * BLOCK] */
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
#pragma unroll 4
		for (int q = 0; q < 16; q++) { /* four 8-byte entries per 256-bit load (the record is 512-byte aligned) */
			uint4 a, b;
			ld_nc_256(src + 4 * q, a, b);
			uint32_t *t = tab + (4 * q) * BLOCK;
			t[0 * BLOCK] = a.y; t[1 * BLOCK] = a.w; t[2 * BLOCK] = b.y; t[3 * BLOCK] = b.w;
		}
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
	/* next 4 stream bytes from the ring */
	auto next_word = [&]() -> uint32_t {
		const uint32_t w = ring[(rd.wi & (QOI_DEC_RING_WORDS - 1)) * BLOCK];
		rd.wi++;
		return w;
	};
	/* Warp-uniform refill: called once per loop trip (at most 4 pixels = 20 stream bytes) by all lanes together.
	   Chunk `filled` is parked in registers; it goes into ring half (filled & 1) as soon as the reader has left chunk
	   filled - 2, and the next chunk is put in flight.  Doing this at the same instruction for all lanes matters: the
	   scoreboard tracks the parked registers per warp, so with lane-driven refills every lane's refill waited for the
	   load another lane had just issued (37 % of the decoder's stall samples in round 1). */
	uint32_t filled = 2;
	auto refill = [&]() {
		if ((rd.wi >> 3) + 1u >= filled) {
			uint32_t *r = ring + ((filled & 1u) * 8u) * BLOCK;
			r[0 * BLOCK] = rd.pa.x; r[1 * BLOCK] = rd.pa.y; r[2 * BLOCK] = rd.pa.z; r[3 * BLOCK] = rd.pa.w;
			r[4 * BLOCK] = rd.pb.x; r[5 * BLOCK] = rd.pb.y; r[6 * BLOCK] = rd.pb.z; r[7 * BLOCK] = rd.pb.w;
			filled++;
			load_chunk(rd);
		}
	};
	uint32_t run = 0;
	{ /* prime the window with the word that holds byte p0 */
		const uint32_t mis = (uint32_t)((uintptr_t)(s + p0) & 3u);
		rd.buf = (uint64_t)(next_word() >> (8u * mis));
		rd.avail = 4u - mis;
	}

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
		const uint4 L = lut[tag];
		const uint32_t idxv = tab[(tag & 63u) * BLOCK]; /* qoi.h:558-560 */
		const uint32_t npx = op_colour(tag, pay, L, idxv, px);
		const uint32_t len = __dp4a(L.y, 0x00000100u, 0u);
		const uint32_t nrun = __dp4a(L.y, 0x00010000u, 0u);
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
		if (vec_ok && i + 4 <= px_end && (((uintptr_t)(o32 + i)) & 31u)) { /* up to a 32-byte boundary */
			const uint32_t a = step(), b = step(), c = step(), d = step();
			*(uint4 *)(o32 + i) = make_uint4(a, b, c, d);
			i += 4;
			refill();
		}
		while (i + 8 <= px_end) { /* eight pixels = one 32-byte sector per lane and store */
			uint4 lo, hi;
			lo.x = step(); lo.y = step(); lo.z = step(); lo.w = step();
			refill(); /* once per four pixels (<= 20 stream bytes): the ring holds two chunks */
			hi.x = step(); hi.y = step(); hi.z = step(); hi.w = step();
			st_256(o32 + i, lo, hi);
			i += 8;
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
		const bool wide_ok = ((uintptr_t)obase & 7u) == 0;
		if (wide_ok && (i & 4u) && i + 4 <= px_end) { /* up to a multiple of 8 pixels: 24-byte groups are 8-byte aligned */
			const uint32_t a = step(), b = step(), c = step(), d = step();
			uint32_t *o = (uint32_t *)(obase + (size_t)i * 3);
			o[0] = (a & 0x00FFFFFFu) | (b << 24);
			o[1] = ((b >> 8) & 0x0000FFFFu) | (c << 16);
			o[2] = ((c >> 16) & 0x000000FFu) | (d << 8);
			i += 4;
			refill();
		}
		while (wide_ok && i + 8 <= px_end) { /* 8 pixels = 24 bytes = three 8-byte stores (a third of the passes of six word stores) */
			const uint32_t a = step(), b = step(), c = step(), d = step();
			refill();
			const uint32_t e = step(), f = step(), g = step(), h = step();
			uint2 *o = (uint2 *)(obase + (size_t)i * 3);
			o[0] = make_uint2((a & 0x00FFFFFFu) | (b << 24), ((b >> 8) & 0x0000FFFFu) | (c << 16));
			o[1] = make_uint2(((c >> 16) & 0x000000FFu) | (d << 8), (e & 0x00FFFFFFu) | (f << 24));
			o[2] = make_uint2(((f >> 8) & 0x0000FFFFu) | (g << 16), ((g >> 16) & 0x000000FFu) | (h << 8));
			i += 8;
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
