/*
 * qoi_enc4.cuh -- EXPERIMENTAL "striped" variant of the chained-scan encoder (QOI_GPU_ENC=striped | striped-match).
 * OFF BY DEFAULT.  Checked bit-exact on the CPU emulator (tests/test_emu.py) and, once, on a B200 (tools/striped_smoke.c:
 * streams byte-identical to the shipped encoder, 8-12 % slower -- profiles/r1_final/striped_smoke.txt; not profiled yet).
 * The shipped encoder is qoi_enc.cuh; this file only adds kernels.
 *
 * Same CTA skeleton, look-back records, loader and scout protocol as qoi_enc.cuh (reference loop: qoi.h:406-478);
 * what changes is how the 8 compute warps walk a tile.  There a THREAD owns 16 consecutive pixels and finds the
 * previous pixel of its hash slot through a slot x position bit matrix (divergent walks, byte-wise stores, ~9 warp
 * instructions per pixel, profiles/r1_final/encode_lines.txt).  Here a WARP owns 512 consecutive pixels and walks them
 * as 16 rows of 32, lane = pixel, so every step of qoi.h's state machine is a warp-wide operation:
 *
 *   px == px_prev (qoi.h:415)      one compare against the left neighbour (shared memory), ballot = non-run mask
 *   index[hash] == px (qoi.h:432)  lanes with my hash slot from six ballots of the hash bits (or match.any); the
 *                                  nearest lower one is my predecessor (shuffle); none in the row -> the warp's own
 *                                  64-entry table of the rows before; none there -> deferred to phase R (<= 64 per warp)
 *   run counter (qoi.h:416-421)    distance to the last non-run pixel: clz of the ballot, else the warp's carry
 *   op selection (qoi.h:438-474)   branch-free classify on the packed pixel
 *   output cursor p                packed shuffle scan over four rows at a time, warp totals, tile look-back (scout)
 *
 * The bit matrix is gone (its shared memory holds the staging buffer, the per-warp tables and the patch words).
 */
#pragma once
#include "qoi_enc.cuh"

namespace qoienc4 {
using namespace qoidev;
using namespace qoienc;

/* what lives in the former bit-matrix area of Layout<CH, EXPAND> */
template <int CH, bool EXPAND>
struct Sub {
	using L = Layout<CH, EXPAND>;
	static constexpr int OFF_STAGE = L::OFF_CELLS;
	static constexpr int OFF_LENT = OFF_STAGE + ((L::STAGE_BYTES + 15) & ~15); /* uint2 [8][64]: last non-run pixel per slot in the warp's chunk */
	static constexpr int OFF_FPOS = OFF_LENT + QOI_ENC_WARPS * 64 * 8;          /* u16 [8][64]: first pixel of a slot in the warp's chunk */
	static constexpr int OFF_PATCH = OFF_FPOS + QOI_ENC_WARPS * 64 * 2;         /* u32 [8][16]: pixels that phase R turned into INDEX hits */
	static constexpr int END = OFF_PATCH + QOI_ENC_WARPS * 16 * 4;
	static_assert(END <= L::OFF_CELLS + L::CELLS_BYTES, "the striped variant must fit into the bit-matrix area");
	static_assert(OFF_LENT % 8 == 0, "alignment");
};

/* qoi.h:438-474 for a non-run pixel that is not an INDEX hit; same result as qoienc::classify, no branches */
__device__ __forceinline__ uint32_t classify_bf(uint32_t v, uint32_t p) {
	const uint32_t d = __vsub4(v, p);           /* per-channel difference mod 256 (qoi.h:439-441) */
	const uint32_t t = __vadd4(d, 0x00020202u); /* dr+2, dg+2, db+2, da */
	const bool is_diff = (t & 0xFFFCFCFCu) == 0u; /* qoi.h:446-451 */
	const uint32_t diff_cd = 1u | ((0x40u | ((t & 3u) << 4) | (((t >> 8) & 3u) << 2) | ((t >> 16) & 3u)) << 8);
	const uint32_t dg = (d >> 8) & 0xFFu;
	const uint32_t e = __vsub4(d, dg * 0x00010101u); /* dr-dg, 0, db-dg, da-dg (qoi.h:443-444) */
	const uint32_t t2 = __vadd4(e, 0x00080008u);
	const uint32_t g32 = (dg + 32u) & 0xFFu;
	const bool a_eq = (d >> 24) == 0u;
	const bool is_luma = a_eq && (t2 & 0x00F000F0u) == 0u && (g32 & 0xC0u) == 0u; /* qoi.h:453-459 */
	const uint32_t luma_cd = 2u | ((0x80u | g32) << 8) | ((((t2 & 0xFu) << 4) | ((t2 >> 16) & 0xFu)) << 16);
	const uint32_t abs_cd = (a_eq ? (4u | (0xFEu << 8)) : (5u | (0xFFu << 8))) | ((v & 0xFFu) << 16); /* qoi.h:461-473 */
	return is_diff ? diff_cd : (is_luma ? luma_cd : abs_cd);
}

/* op of a run pixel (qoi.h:415-421): k = its position in the run (1-based).  The run byte 0xC0 | (k-1)%62 is kept in
   bits 8..15 even when nothing is emitted, so that "an op starts here" ((k-1)%62 == 0) can be read back later */
__device__ __forceinline__ uint32_t run_code(uint32_t k, bool ends) {
	const uint32_t r = (k - 1u) % 62u;
	return ((r == 61u || ends) ? 1u : 0u) | ((0xC0u | r) << 8);
}

/* plane = ballot(x != 0), mine = x != 0 ? ~0 : 0, off one predicate (the compiler's own form recomputes the bit three times) */
__device__ __forceinline__ void ballot_and_mask(uint32_t x, uint32_t &plane, uint32_t &mine) {
#ifdef QOI_EMU
	plane = __ballot_sync(QOI_FULL_MASK, x != 0u);
	mine = x != 0u ? 0xFFFFFFFFu : 0u;
#else
	asm volatile(
	    "{\n\t.reg .pred p;\n\t"
	    "setp.ne.u32 p, %2, 0;\n\t"
	    "vote.sync.ballot.b32 %0, p, 0xffffffff;\n\t"
	    "selp.u32 %1, 0xffffffff, 0, p;\n\t}"
	    : "=r"(plane), "=r"(mine)
	    : "r"(x));
#endif
}


/* ---------------------------------------------------------------- compact layout (slab mode, QOI_GPU_ENC=slab4 | slab4-pipe) */
/* bytes of the slab-mode work area: 8 slabs, per-warp tables, first-occurrence positions, hit and non-run words */
template <int CH>
struct SlabArea {
	static constexpr int SLAB_BYTES = QOI_WARP_PX * (CH + 1) + 16;
	static constexpr int BYTES = QOI_ENC_WARPS * SLAB_BYTES + QOI_ENC_WARPS * 64 * 8 + QOI_ENC_WARPS * 64 * 2 + 2 * QOI_ENC_WARPS * QOI_WARP_STEPS * 4;
};
/*
 * The kernel is bound by dependent latency (profiles/r1_final/striped_smoke.txt), so the obvious lever is more resident
 * warps.  Slab mode needs 48 registers with 8 bytes of spill and, with ONE pixel buffer instead of two and the work area
 * cut to what slab mode uses, 45 KB of shared memory (RGBA zero-copy): 4 CTAs (40 warps) per SM instead of 3 (30).  The
 * price: the next tile's bulk copy can only start when the compute warps are done with the current pixels, so its latency
 * is covered by the other CTAs of the SM instead of by this CTA's own work.
 */
template <int CH, bool EXPAND>
struct LayoutC {
	static constexpr int NBUF = 1;
	static constexpr int RAW_STRIDE = 16 + 16 + QOI_TILE_PX * CH + 16;
	static constexpr int OFF_RAW = 0;
	static constexpr int OFF_PXBUF = NBUF * RAW_STRIDE;
	static constexpr int PXBUF_BYTES = EXPAND ? (16 + QOI_TILE_PX * 4 + 16) : 0;
	static constexpr int OFF_CELLS = OFF_PXBUF + PXBUF_BYTES;
	static constexpr int CELLS_BYTES = (SlabArea<CH>::BYTES + 15) & ~15;
	static constexpr int OFF_TINVAL = OFF_CELLS + CELLS_BYTES;        /* u32 [64] */
	static constexpr int OFF_HITW = OFF_TINVAL + 64 * 4;              /* u32 [16]: the scout's per-warp output offsets, two copies */
	static constexpr int OFF_TINFLAG = OFF_HITW + 16 * 4;             /* u8 [64] */
	static constexpr int OFF_MBAR = OFF_TINFLAG + 64;                 /* u64 [2] */
	static constexpr int OFF_INFO = OFF_MBAR + 16;                    /* TileInfo [2] */
	static constexpr int OFF_NBR = OFF_INFO + 2 * (int)sizeof(TileInfo);
	static constexpr int OFF_WTOT = OFF_NBR + 16;
	static constexpr int OFF_WLNR = OFF_WTOT + 32;
	static constexpr int OFF_WLNRIN = OFF_WLNR + 32;
	static constexpr int OFF_MISC = OFF_WLNRIN + 32;
	static constexpr int TOTAL = OFF_MISC + 32;
	static_assert(OFF_CELLS % 16 == 0 && OFF_TINVAL % 16 == 0 && OFF_MBAR % 8 == 0 && OFF_INFO % 16 == 0, "alignment");
};
template <int CH, bool EXPAND, bool COMPACT>
struct PickLayout { using type = Layout<CH, EXPAND>; };
template <int CH, bool EXPAND>
struct PickLayout<CH, EXPAND, true> { using type = LayoutC<CH, EXPAND>; };

/* qoienc::fetch_issue for any layout type L (same code; the shipped function is tied to Layout<CH, EXPAND>).
   step 2: start the bulk copy into raw buffer `rb`, fill fragments / neighbours, publish the info record `ib` */
template <class L, int CH, bool EXPAND>
__device__ __forceinline__ void fetch_issue_l(const QoiEncParams &P, unsigned char *smem, const Fetch &F, int rb, int ib) {
	const uint32_t lane = lane_id();
	TileInfo *info = (TileInfo *)(smem + L::OFF_INFO) + ib;
	uint64_t *mbar = (uint64_t *)(smem + L::OFF_MBAR) + rb;
	if (!F.any) {
		if (lane == 0) info->tile = -1;
		return;
	}
	const QoiImgMeta &m = F.m;
	const uint32_t t_in = F.tile - m.first_tile;
	const uint32_t px0 = t_in * QOI_TILE_PX;
	const uint32_t n_valid = (m.npx - px0) < (uint32_t)QOI_TILE_PX ? (m.npx - px0) : (uint32_t)QOI_TILE_PX;
	const uint8_t *gsrc = P.pixels + m.pix_off + (uint64_t)px0 * CH;
	const uint32_t nbytes = n_valid * CH;
	const uint32_t sh = EXPAND ? (uint32_t)((uintptr_t)gsrc & 15u) : 0u;
	unsigned char *data = smem + L::OFF_RAW + rb * L::RAW_STRIDE + 16; /* data[sh + k] = gsrc[k] */
	const uint32_t lo_off = (sh + 15u) & ~15u;
	const uint32_t hi_off = (sh + nbytes) & ~15u;
	const uint32_t tma_bytes = hi_off > lo_off ? hi_off - lo_off : 0u;
	if (lane == 0) {
		mbar_arrive_expect_tx(mbar, tma_bytes);
		if (tma_bytes) tma_load_1d_hint(data + lo_off, gsrc + (lo_off - sh), tma_bytes, mbar, l2_evict_first_policy());
	}
	/* the <16-byte fragments on either side of the 16-byte aligned interior */
	uint32_t a_lo = sh, a_hi = tma_bytes ? lo_off : sh + nbytes;
	uint32_t b_lo = tma_bytes ? hi_off : 0u, b_hi = tma_bytes ? sh + nbytes : 0u;
	uint32_t o = a_lo + lane;
	if (o < a_hi) data[o] = gsrc[o - sh];
	o = b_lo + lane;
	if (o < b_hi) data[o] = gsrc[o - sh];
	/* neighbours: the pixel before the tile (start colour {0,0,0,255} at the image start, qoi.h:396-399)
	   and the pixel after it (needed to know whether a run ends on the tile's last pixel) */
	uint32_t nb = 0;
	if (lane == 0) {
		if (t_in == 0) nb = 0xFF000000u;
		else {
			const uint8_t *q = gsrc - CH;
			nb = (uint32_t)q[0] | ((uint32_t)q[1] << 8) | ((uint32_t)q[2] << 16) | (CH == 4 ? ((uint32_t)q[CH - 1] << 24) : 0xFF000000u);
		}
	} else if (lane == 1) {
		if (px0 + n_valid < m.npx) {
			const uint8_t *q = gsrc + nbytes;
			nb = (uint32_t)q[0] | ((uint32_t)q[1] << 8) | ((uint32_t)q[2] << 16) | (CH == 4 ? ((uint32_t)q[CH - 1] << 24) : 0xFF000000u);
		}
	}
	if (lane < 2) {
		if (EXPAND) ((uint32_t *)(smem + L::OFF_NBR))[ib * 2 + lane] = nb;
		else ((uint32_t *)data)[lane == 0 ? -1 : QOI_TILE_PX] = nb;
	}
	if (lane == 0) {
		info->tile = (int32_t)F.tile;
		info->img = F.img;
		info->t_in_img = t_in;
		info->n_valid = n_valid;
		info->px0 = px0;
		info->npx = m.npx;
		info->first_tile = m.first_tile;
		info->sh = sh;
		info->stream_off = m.stream_off;
		info->width = m.width;
		info->height = m.height;
		info->chan_cs = (uint32_t)m.channels | ((uint32_t)m.colorspace << 8);
	}
}

/* ---------------------------------------------------------------- "slab" mode (QOI_GPU_ENC=slab | slab-match) */
/*
 * What the first GPU run of the striped kernel said (profiles/r1_final/striped_smoke.txt): 40 % fewer instructions did not
 * buy time, so the limiter is dependent latency and (probably) instruction fetch -- the striped main pass is unrolled 16 x
 * because it keeps one op code per row in registers, ~7 700 SASS instructions of which every warp walks most once per tile.
 * Slab mode keeps the striped idea and removes both: rolled loops, nothing kept per row in registers.
 *
 *   index pass   (before T_READY) the serial part only: non-run mask, same-slot predecessor, per-warp table, deferred
 *                first occurrences; leaves one non-run word and one hit word per row in shared memory
 *   phase R      sets the hit bits of the deferred pixels in those words
 *   emit pass    (after T_READY) rows are independent now: classify, sizes, packed scan over four rows, and the op bytes
 *                go straight into the warp's own SLAB of the staging area (512 px x 5 bytes) at a running offset
 *   copy-out     after B_READY every warp copies its slab to tile_base + (bytes of the warps before it) with a
 *                byte-shifted 16-byte copy -- no second pass over the ops, no staging hand-off, no bulk store
 */
template <class L, int CH>
struct SubSlab {
	static constexpr int SLAB_BYTES = SlabArea<CH>::SLAB_BYTES; /* worst case of a chunk + room for the shifted copy's last word */
	static constexpr int OFF_SLAB = L::OFF_CELLS;
	static constexpr int OFF_LENT = OFF_SLAB + QOI_ENC_WARPS * SLAB_BYTES;
	static constexpr int OFF_FPOS = OFF_LENT + QOI_ENC_WARPS * 64 * 8;
	static constexpr int OFF_HITW = OFF_FPOS + QOI_ENC_WARPS * 64 * 2;                   /* u32 [8][16] */
	static constexpr int OFF_NRW = OFF_HITW + QOI_ENC_WARPS * QOI_WARP_STEPS * 4;        /* u32 [8][16] */
	static constexpr int END = OFF_NRW + QOI_ENC_WARPS * QOI_WARP_STEPS * 4;
	static_assert(END <= L::OFF_CELLS + L::CELLS_BYTES, "slab mode must fit into the bit-matrix area");
	static_assert(SLAB_BYTES % 16 == 0 && OFF_LENT % 8 == 0, "alignment");
};

/* what a warp keeps of a tile between its emit pass and its copy-out (PIPE: across the next tile's index pass) */
struct SlabCarry {
	bool valid;
	uint32_t t;          /* tile */
	uint8_t *gstream;    /* its image's stream */
	uint32_t bytes;      /* bytes in my slab */
	int32_t rs_px;       /* restart pixel (image coordinates), INT_MAX: none in this tile (warp 0 only) */
	uint32_t rs_off;     /* its byte offset inside the tile, 0xFFFFFFFF in lanes that do not own it */
	uint32_t rs_prev;    /* the pixel before it */
	uint32_t parity;     /* which copy of the scout's per-tile outputs belongs to the tile (PIPE) */
};

/* misc[] holds two copies of {tile base, -, tile total} so that the scout can finish tile k+1 while tile k is still copied out */
__device__ __forceinline__ uint32_t misc_slot(uint32_t parity, uint32_t what) { return parity * 4u + what; }

/* B_READY, restart record, my slab -> tile_base + bytes of the warps before me (byte-shifted 16-byte copy) */
template <int CH, bool EXPAND, bool PIPE, bool COMPACT>
__device__ __forceinline__ void slab_finish(const QoiEncParams &P, unsigned char *smem, const SlabCarry &c) {
	using L = typename PickLayout<CH, EXPAND, COMPACT>::type;
	using S = SubSlab<L, CH>;
	const uint32_t tid = threadIdx.x, lane = tid & 31u, warp = tid >> 5;
	const unsigned char *slab = smem + S::OFF_SLAB + warp * S::SLAB_BYTES;
	const uint32_t *misc = (const uint32_t *)(smem + L::OFF_MISC);
	const uint32_t *wtot = (const uint32_t *)(smem + L::OFF_WTOT);
	const uint32_t *wpre = (const uint32_t *)(smem + L::OFF_HITW); /* PIPE: u32 [2][8] exclusive prefix of wtot, written by the scout */
	bar_sync(BAR_B_READY, QOI_N_CS);
	const uint32_t tile_base = misc[misc_slot(PIPE ? c.parity : 0u, MISC_TILE_BASE)];
	uint32_t woff = 0;
	if (PIPE) woff = wpre[c.parity * QOI_ENC_WARPS + warp];
	else for (uint32_t w = 0; w < warp; w++) woff += wtot[w];
	if (warp == 0 || warp == QOI_ENC_WARPS / 2) { /* restart records: tile start (warp 0), second half (warp 4) */
		int32_t *o_px = warp ? P.st.mid_px : P.st.rs_px, *o_byte = warp ? P.st.mid_byte : P.st.rs_byte;
		uint32_t *o_prev = warp ? P.st.mid_prev : P.st.rs_prev;
		if (c.rs_off != 0xFFFFFFFFu) { /* the lane that owns the restart pixel records where its op starts */
			o_px[c.t] = c.rs_px;
			o_byte[c.t] = (int32_t)(tile_base + woff + c.rs_off);
			o_prev[c.t] = c.rs_prev;
		}
		if (lane == 0 && c.rs_px == INT_MAX) o_px[c.t] = -1;
	}
	uint8_t *gdst = c.gstream + tile_base + woff;
	const uint32_t total = c.bytes;
	uint32_t head = (16u - (uint32_t)((uintptr_t)gdst & 15u)) & 15u; /* bytes up to the first aligned 16 */
	if (head > total) head = total;
	const uint32_t nvec = (total - head) >> 4;
	if (lane < head) gdst[lane] = slab[lane];
	const uint32_t *s32 = (const uint32_t *)slab + (head >> 2);
	const uint32_t sh = (head & 3u) * 8u;
	for (uint32_t k = lane; k < nvec; k += 32u) {
		const uint32_t *w = s32 + 4u * k;
		const uint32_t w0 = w[0], w1 = w[1], w2 = w[2], w3 = w[3], w4 = sh ? w[4] : 0u;
		uint4 q;
		q.x = __funnelshift_r(w0, w1, sh);
		q.y = __funnelshift_r(w1, w2, sh);
		q.z = __funnelshift_r(w2, w3, sh);
		q.w = __funnelshift_r(w3, w4, sh);
		*(uint4 *)(gdst + head + 16u * k) = q;
	}
	const uint32_t tail0 = head + 16u * nvec;
	if (tail0 + lane < total) gdst[tail0 + lane] = slab[tail0 + lane]; /* fewer than 16 bytes */
	__syncwarp();
}

/* PIPE: the copy-out of a tile (and the wait for its output offset) moves behind the NEXT tile's index pass:
   per iteration  index(k) -> copy-out(k-1) -> emit(k); the pixel buffer is released to the loader after emit(k) already */
template <int CH, bool EXPAND, bool MATCH, bool PIPE, bool COMPACT>
__device__ __forceinline__ void slab_compute(const QoiEncParams &P, unsigned char *smem, const TileInfo &ti, const uint32_t *pxv,
                                             uint8_t *gstream, uint32_t parity, SlabCarry &carry) {
	using L = typename PickLayout<CH, EXPAND, COMPACT>::type;
	using S = SubSlab<L, CH>;
	const uint32_t tid = threadIdx.x, lane = tid & 31u, warp = tid >> 5;
	const uint32_t t = (uint32_t)ti.tile;
	unsigned char *slab = smem + S::OFF_SLAB + warp * S::SLAB_BYTES;
	uint2 *Lent_all = (uint2 *)(smem + S::OFF_LENT);
	uint2 *Lent = Lent_all + warp * 64;
	uint16_t *Fpos = (uint16_t *)(smem + S::OFF_FPOS) + warp * 64;
	uint32_t *hitw = (uint32_t *)(smem + S::OFF_HITW) + warp * QOI_WARP_STEPS;
	uint32_t *nrw = (uint32_t *)(smem + S::OFF_NRW) + warp * QOI_WARP_STEPS;
	const uint32_t *Tin_val = (const uint32_t *)(smem + L::OFF_TINVAL);
	const uint8_t *Tin_flag = (const uint8_t *)(smem + L::OFF_TINFLAG);
	uint32_t *wtot = (uint32_t *)(smem + L::OFF_WTOT);
	int32_t *wlnr = (int32_t *)(smem + L::OFF_WLNR);
	const int32_t *wlnr_in = (const int32_t *)(smem + L::OFF_WLNRIN);
	const uint32_t *misc = (const uint32_t *)(smem + L::OFF_MISC);

	Lent[lane] = make_uint2(0u, 0u);
	Lent[lane + 32] = make_uint2(0u, 0u);
	Fpos[lane] = 0xFFFFu;
	Fpos[lane + 32] = 0xFFFFu;
	if (lane < QOI_WARP_STEPS) {
		hitw[lane] = 0u;
		nrw[lane] = 0u;
	}
	__syncwarp();

	const uint32_t nv = ti.n_valid; /* <= 4096; only an image's last tile can be partial */
	const bool last_tile = ti.px0 + ti.n_valid == ti.npx;
	const uint32_t lt = lanemask_lt(), gt = lanemask_gt();
	const uint32_t wbase = warp * (uint32_t)QOI_WARP_PX;

	/* ---- index pass: the only part in which a row depends on the row before it (through the warp's table) ---- */
	{
		int32_t last_nr = -1; /* tile-local index of the last non-run pixel of this chunk so far (uniform) */
		uint32_t v = pxv[wbase + lane], p = pxv[(int)(wbase + lane) - 1]; /* row 0; the next row's pixels are loaded a row ahead */
#pragma unroll 1
		for (int s = 0; s < QOI_WARP_STEPS; s++) {
			const uint32_t rb0 = wbase + (uint32_t)s * 32u;
			if (rb0 >= nv) break; /* uniform */
			const uint32_t il = rb0 + lane;
			const uint32_t iln = (s + 1 < QOI_WARP_STEPS) ? il + 32u : il; /* stays inside pxv[-1 .. 4096] */
			const uint32_t v_next = pxv[iln], p_next = pxv[(int)iln - 1];
			const bool nonrun = il < nv && v != p; /* qoi.h:415 */
			const uint32_t nr = __ballot_sync(QOI_FULL_MASK, nonrun);
			const uint32_t h = qoi_slot(v);
			uint32_t m; /* the non-run lanes of this row with my slot */
			if (MATCH) {
				m = __match_any_sync(QOI_FULL_MASK, nonrun ? h : 64u + lane);
			} else {
				uint32_t differ = 0;
#pragma unroll
				for (int b = 0; b < 6; b++) {
					uint32_t plane, mine;
					ballot_and_mask(h & (1u << b), plane, mine);
					differ |= plane ^ mine;
				}
				m = nr & ~differ;
			}
			const uint32_t lower = m & lt;
			const int src = lower ? (31 - __clz((int)lower)) : (int)lane;
			const uint32_t pv = __shfl_sync(QOI_FULL_MASK, v, src);
			const uint2 ent = Lent[h];
			bool hit = false;
			if (nonrun) {
				if (lower) hit = (pv == v);          /* predecessor with my slot inside this row */
				else if (ent.y) hit = (ent.x == v);  /* ... in an earlier row of this chunk */
				else Fpos[h] = (uint16_t)il;         /* first pixel of the slot in the chunk: phase R decides */
			}
			const uint32_t hb = __ballot_sync(QOI_FULL_MASK, hit);
			__syncwarp();
			if (nonrun && (m & gt) == 0u) Lent[h] = make_uint2(v, 1u); /* last pixel of its slot in this row */
			if (lane == 0) {
				nrw[s] = nr;
				hitw[s] = hb;
			}
			if (nr) last_nr = (int32_t)(rb0 + 31u - (uint32_t)__clz((int)nr));
			__syncwarp();
			v = v_next;
			p = p_next;
		}
		if (lane == 0) wlnr[warp] = last_nr;
	}
	__syncwarp();
	bar_arrive(BAR_I_DONE, QOI_N_CS);

	if (PIPE && carry.valid) slab_finish<CH, EXPAND, true, COMPACT>(P, smem, carry); /* the previous tile leaves while the scout looks back for this one */

	bar_sync(BAR_T_READY, QOI_N_CS); /* every warp's table is final; T_in and the run carry are in shared memory */
	/* ---- phase R: my chunk's deferred pixels against the earlier warps' tables, then the incoming table ---- */
#pragma unroll
	for (int q = 0; q < 2; q++) {
		const uint32_t sl = lane + 32u * q;
		const uint32_t pos = Fpos[sl];
		if (pos != 0xFFFFu) {
			const uint32_t pvv = pxv[pos];
			bool found = false;
			uint32_t val = 0;
			for (int w2 = (int)warp - 1; w2 >= 0; w2--) {
				const uint2 e = Lent_all[w2 * 64 + sl];
				if (e.y) {
					val = e.x;
					found = true;
					break;
				}
			}
			if (!found && Tin_flag[sl]) {
				val = Tin_val[sl];
				found = true;
			}
			/* an untouched index[] slot holds 0 (qoi.h:393): only the all-zero pixel can hit it */
			const bool hit = found ? (val == pvv) : (pvv == 0u);
			if (hit) atomicOr(&hitw[(pos >> 5) & (QOI_WARP_STEPS - 1)], 1u << (pos & 31u));
		}
	}
	__syncwarp();
	/* last non-run pixel before the chunk (image coordinates; -1: none since the image start) */
	int32_t lnr_img = wlnr_in[0];
	for (int w2 = (int)warp - 1; w2 >= 0; w2--) {
		if (wlnr[w2] >= 0) {
			lnr_img = (int32_t)ti.px0 + wlnr[w2];
			break;
		}
	}
	/* table at the tile's mid point = what qoi.h's index[] holds before the first pixel of warp 4's chunk: the nearest of the
	   first four warps' tables, else the incoming table (the same walk phase R does); pixels between the mid point and the
	   restart op are run pixels and write nothing.  Read by the decoder when it runs two lanes per tile. */
	if (warp == QOI_ENC_WARPS / 2) {
#pragma unroll
		for (int q = 0; q < 2; q++) {
			const uint32_t sl = lane + 32u * q;
			bool found = false;
			uint32_t val = 0;
			for (int w2 = (int)warp - 1; w2 >= 0; w2--) {
				const uint2 e = Lent_all[w2 * 64 + sl];
				if (e.y) {
					val = e.x;
					found = true;
					break;
				}
			}
			if (!found && Tin_flag[sl]) {
				val = Tin_val[sl];
				found = true;
			}
			P.st.mid_ent[(size_t)t * 64 + sl] = ((uint64_t)val << 32) | (found ? 1ull : 0ull);
		}
	}
	/* PIPE: nobody waits for B_READY before the next tile's index pass re-initialises the tables, so everybody must be
	   done reading the other warps' tables (phase R) and wlnr[] before anybody goes on */
	if (PIPE) bar_sync(BAR_COMPUTE, QOI_ENC_COMPUTE);

	/* ---- emit pass: four independent rows at a time; op bytes go into the warp's slab at a running offset ---- */
	uint32_t off = 0;                 /* bytes of this chunk so far (uniform) */
	/* decoder restart points: the first pixel at which an op starts in the tile (warp 0) and in its second half (warp 4) */
	const bool rs_owner = warp == 0 || warp == QOI_ENC_WARPS / 2;
	int32_t cand = INT_MAX;
	uint32_t rs_off = 0xFFFFFFFFu;    /* its byte offset, in the lane that owns the pixel */
#pragma unroll 1
	for (int g = 0; g < QOI_WARP_STEPS / 4; g++) {
		if (wbase + (uint32_t)g * 128u >= nv) break; /* uniform */
		uint32_t cd[4];
		uint32_t starts = 0; /* warp 0, group 0: bit k = an op starts at my pixel of row k */
#pragma unroll
		for (int k = 0; k < 4; k++) {
			const int s = 4 * g + k;
			const uint32_t rb0 = wbase + (uint32_t)s * 32u;
			const uint32_t il = rb0 + lane;
			const uint32_t nr = nrw[s], hb = hitw[s];
			uint32_t c = 0;
			if (rb0 < nv) { /* uniform */
				const uint32_t v = pxv[il], p = pxv[(int)il - 1];
				const bool valid = il < nv;
				const bool nonrun = (nr >> lane) & 1u;
				if (nonrun) {
					c = ((hb >> lane) & 1u) ? (1u | (qoi_slot(v) << 8)) : classify_bf(v, p); /* qoi.h:432-434 / 438-474 */
				} else if (valid) { /* qoi.h:415-421 */
					const uint32_t below = nr & lt;
					const int32_t lnr = below ? (int32_t)(ti.px0 + rb0 + 31u - (uint32_t)__clz((int)below)) : lnr_img;
					const uint32_t gi = ti.px0 + il;
					c = run_code(gi - (uint32_t)lnr, (last_tile && gi == ti.npx - 1u) || pxv[il + 1] != v);
				}
				if (rs_owner && g == 0 && k < 2 && valid && (nonrun || ((c >> 8) & 0x3Fu) == 0u)) starts |= 1u << k;
				if (nr) lnr_img = (int32_t)(ti.px0 + rb0 + 31u - (uint32_t)__clz((int)nr));
			}
			cd[k] = c;
		}
		/* packed inclusive scan of the four rows' sizes (a row holds <= 160 bytes) */
		const uint32_t pk = (cd[0] & 7u) | ((cd[1] & 7u) << 8) | ((cd[2] & 7u) << 16) | ((cd[3] & 7u) << 24);
		uint32_t incl = pk;
#pragma unroll
		for (int dlt = 1; dlt < 32; dlt <<= 1) {
			const uint32_t o = __shfl_up_sync(QOI_FULL_MASK, incl, dlt);
			if (lane >= (uint32_t)dlt) incl += o;
		}
		const uint32_t ex = incl - pk, tot = __shfl_sync(QOI_FULL_MASK, incl, 31);
		if (rs_owner && g == 0) {
#pragma unroll
			for (int k = 0; k < 2; k++) {
				const uint32_t b = __ballot_sync(QOI_FULL_MASK, (starts >> k) & 1u);
				if (b && cand == INT_MAX) {
					cand = (int32_t)(wbase + (uint32_t)k * 32u + (uint32_t)__ffs((int)b) - 1u);
					if ((int32_t)(wbase + (uint32_t)k * 32u + lane) == cand) rs_off = off + (k ? (tot & 0xFFu) : 0u) + ((ex >> (8 * k)) & 0xFFu);
				}
			}
		}
#pragma unroll
		for (int k = 0; k < 4; k++) {
			const uint32_t c = cd[k], sz = c & 7u;
			unsigned char *o = slab + off + ((ex >> (8 * k)) & 0xFFu);
			if (sz) {
				o[0] = (uint8_t)(c >> 8);
				if (sz >= 2) {
					o[1] = (uint8_t)(c >> 16);
					if (sz >= 4) {
						const uint32_t v = pxv[wbase + (uint32_t)(4 * g + k) * 32u + lane];
						o[2] = (uint8_t)(v >> 8);
						o[3] = (uint8_t)(v >> 16);
						if (sz == 5) o[4] = (uint8_t)(v >> 24);
					}
				}
			}
			off += (tot >> (8 * k)) & 0xFFu;
		}
	}
	if (lane == 0) wtot[warp] = off;
	carry.valid = true;
	carry.t = t;
	carry.gstream = gstream;
	carry.bytes = off;
	carry.rs_px = (rs_owner && cand != INT_MAX) ? (int32_t)ti.px0 + cand : INT_MAX;
	carry.rs_off = rs_off;
	carry.rs_prev = (rs_owner && cand != INT_MAX) ? pxv[cand - 1] : 0u;
	carry.parity = parity;
	__syncwarp();
	bar_arrive(BAR_C_DONE, QOI_N_CS);
	if (PIPE) {
		bar_arrive(BAR_W_DONE, QOI_N_CS); /* the pixel buffer is free (nothing reads pxv any more): the loader may refill it */
	} else {
		slab_finish<CH, EXPAND, false, COMPACT>(P, smem, carry);
		bar_arrive(BAR_W_DONE, QOI_N_CS);
	}
}

/* MATCH: same-slot lanes from match.any instead of six ballots (A/B switch).
   EARLYW: phase W does not wait for the tile's output offset -- the op bytes are staged from offset 0 as soon as the warp
   totals are known, the wait for B_READY comes afterwards (the look-back latency hides behind the emission) and the compute
   warps copy the staging buffer out themselves with a byte-shifted 16-byte copy, instead of the loader's bulk store (which
   needs source and destination congruent mod 16).  The STAGE_FREE hand-off is not needed then. */
/* SLAB: the compute warps run slab_compute() above (no staging hand-off either, like EARLYW) */
/* PIPE (slab mode only): see slab_compute -- the scout then also writes header, end marker and length (the loader no longer
   learns when a tile's bytes are out) and publishes the per-warp output offsets in two alternating copies */
/* COMPACT (slab mode only): LayoutC above -- one pixel buffer, 4 CTAs per SM */
template <int CH, bool EXPAND, bool MATCH, bool EARLYW = false, bool SLAB = false, bool PIPE = false, bool COMPACT = false>
__global__ void __launch_bounds__(QOI_ENC_THREADS, COMPACT ? 4 : 3) qoi_encode_striped_kernel(const QoiEncParams P) {
	static_assert(!COMPACT || SLAB, "the compact layout only holds what slab mode needs");
	constexpr bool NOSTAGE = EARLYW || SLAB; /* the loader never touches the staging area */
	using L = typename PickLayout<CH, EXPAND, COMPACT>::type;
	using S = Sub<CH, EXPAND>; /* striped (non-slab) work area; only meaningful with Layout<CH, EXPAND> */
	QOI_DYNAMIC_SMEM(smem);
	const uint32_t tid = threadIdx.x, lane = tid & 31u, warp = tid >> 5;
	const bool scout = warp == QOI_ENC_WARPS;
	const bool loader = warp == QOI_ENC_WARPS + 1;
	uint64_t *mbar = (uint64_t *)(smem + L::OFF_MBAR);
	TileInfo *infos = (TileInfo *)(smem + L::OFF_INFO);
	unsigned char *stage = smem + S::OFF_STAGE;
	uint2 *Lent_all = (uint2 *)(smem + (SLAB ? SubSlab<L, CH>::OFF_LENT : S::OFF_LENT));
	uint16_t *Fpos_all = (uint16_t *)(smem + S::OFF_FPOS);
	uint32_t *patch_all = (uint32_t *)(smem + S::OFF_PATCH);
	uint32_t *Tin_val = (uint32_t *)(smem + L::OFF_TINVAL);
	uint8_t *Tin_flag = (uint8_t *)(smem + L::OFF_TINFLAG);
	uint32_t *wtot = (uint32_t *)(smem + L::OFF_WTOT);
	int32_t *wlnr = (int32_t *)(smem + L::OFF_WLNR);
	int32_t *wlnr_in = (int32_t *)(smem + L::OFF_WLNRIN);
	uint32_t *misc = (uint32_t *)(smem + L::OFF_MISC);

	if (tid == 0) {
		mbar_init(&mbar[0], 1);
		mbar_init(&mbar[1], 1);
		mbar_fence_init();
	}
	__syncthreads();
	if (loader) {
		Fetch F;
		fetch_ticket(P, F);
		fetch_issue_l<L, CH, EXPAND>(P, smem, F, 0, 0);
		__syncwarp();
		bar_arrive(BAR_INFO_READY, QOI_N_ALL);
		if (!NOSTAGE) bar_arrive(BAR_STAGE_FREE, QOI_N_CS);
	}

	int cur = 0;
	uint32_t parity0 = 0, parity1 = 0;
	const uint32_t epoch = P.epoch;
	SlabCarry carry; /* slab mode: a compute warp's tile that is finished but not copied out yet */
	carry.valid = false;

	while (true) {
		if (!loader) bar_sync(BAR_INFO_READY, QOI_N_ALL); /* infos[cur] is published */
		const TileInfo ti = infos[cur];
		if (ti.tile < 0) break;
		const uint32_t t = (uint32_t)ti.tile;
		const int rb = (EXPAND || COMPACT) ? 0 : cur;
		unsigned char *data = smem + L::OFF_RAW + rb * L::RAW_STRIDE + 16;
		uint32_t *pxv = EXPAND ? (uint32_t *)(smem + L::OFF_PXBUF + 16) : (uint32_t *)data; /* pxv[-1 .. 4096] */
		uint8_t *gstream = P.streams + ti.stream_off;

		if (loader) {
			/* =============================== loader warp (as in qoi_enc.cuh) =============================== */
			Fetch F;
			fetch_ticket(P, F);
			if (EXPAND) bar_sync(BAR_E_DONE, QOI_N_CS);
			if (COMPACT && !EXPAND) bar_sync(BAR_W_DONE, QOI_N_CS); /* one pixel buffer: the compute warps must be done with it */
			fetch_issue_l<L, CH, EXPAND>(P, smem, F, (EXPAND || COMPACT) ? 0 : (cur ^ 1), cur ^ 1);

			if (!(COMPACT && !EXPAND)) bar_sync(BAR_W_DONE, QOI_N_CS); /* staging is complete (COMPACT: waited for above) */
			__syncwarp();
			bar_arrive(BAR_INFO_READY, QOI_N_ALL); /* only now: see the note in qoi_enc.cuh */
			{
				const uint32_t excl = misc[MISC_TILE_BASE], total = misc[MISC_TILE_TOTAL];
				uint8_t *gdst = gstream + excl;
				const uint32_t padb = (uint32_t)((uintptr_t)gdst & 15u);
				const uint32_t lo_off = (padb + 15u) & ~15u;
				const uint32_t end = padb + total;
				const uint32_t hi_off = end & ~15u;
				const bool bulk = !NOSTAGE && hi_off > lo_off;
				if (lane == 0 && bulk) {
					fence_proxy_async_smem();
					tma_store_1d_hint(gdst + (lo_off - padb), stage + lo_off, hi_off - lo_off, l2_evict_first_policy());
					tma_store_commit();
				}
				const uint32_t a_hi = bulk ? lo_off : end;
				const uint32_t b_lo = bulk ? hi_off : end;
				if (!NOSTAGE) { /* EARLYW / SLAB: the compute warps have copied everything out before W_DONE */
					for (uint32_t o = padb + lane; o < a_hi; o += 32) gdst[o - padb] = stage[o];
					for (uint32_t o = b_lo + lane; o < end; o += 32) gdst[o - padb] = stage[o];
				}
				if (!PIPE && t == ti.first_tile && lane < QOI_HEADER_BYTES) { /* header, qoi.h:384-388 (PIPE: the scout writes it) */
					const uint32_t k = lane;
					uint8_t b;
					if (k < 4) b = (uint8_t)("qoif"[k]);
					else if (k < 8) b = (uint8_t)(ti.width >> (8 * (7 - k)));
					else if (k < 12) b = (uint8_t)(ti.height >> (8 * (11 - k)));
					else if (k == 12) b = (uint8_t)(ti.chan_cs & 0xFFu);
					else b = (uint8_t)(ti.chan_cs >> 8);
					gstream[k] = b;
				}
				if (!PIPE && ti.px0 + ti.n_valid == ti.npx && lane >= 16 && lane < 16 + QOI_END_BYTES + 1) { /* end marker, length: qoi.h:480-484 */
					const uint32_t k = lane - 16;
					const uint32_t fin = excl + total;
					if (k < QOI_END_BYTES) gstream[fin + k] = (k == QOI_END_BYTES - 1) ? 1 : 0;
					else P.lens[ti.img] = (int32_t)(fin + QOI_END_BYTES);
				}
				if (!NOSTAGE && lane == 0) tma_store_wait_read(); /* the bulk store has finished reading the staging buffer */
			}
			__syncwarp();
			if (!NOSTAGE) bar_arrive(BAR_STAGE_FREE, QOI_N_CS);
		} else if (scout) {
			/* =============================== scout warp =============================== */
			bar_sync(BAR_I_DONE, QOI_N_CS); /* the per-warp tables and wlnr[] of this tile are complete */
			/* ---- phase T: tile aggregate = the last warp that wrote each slot (lane handles slots lane, lane+32) ---- */
			uint32_t own_val[2];
			bool own_ok[2];
#pragma unroll
			for (int q = 0; q < 2; q++) {
				const uint32_t sl = lane + 32u * q;
				own_ok[q] = false;
				own_val[q] = 0;
				for (int w = QOI_ENC_WARPS - 1; w >= 0; w--) {
					const uint2 e = Lent_all[w * 64 + sl];
					if (e.y) {
						own_val[q] = e.x;
						own_ok[q] = true;
						break;
					}
				}
			}
			int32_t own_lnr = QOI_LNR_NONE;
			for (int w = QOI_ENC_WARPS - 1; w >= 0; w--) {
				if (wlnr[w] >= 0) {
					own_lnr = (int32_t)ti.px0 + wlnr[w];
					break;
				}
			}
			const uint32_t own_lo = __ballot_sync(QOI_FULL_MASK, own_ok[0]);
			const uint32_t own_hi = __ballot_sync(QOI_FULL_MASK, own_ok[1]);
			LookBack lb;
			lb.in_ok[0] = lb.in_ok[1] = false;
			lb.in_val[0] = lb.in_val[1] = 0;
			lb.lnr_in = -1;
			if (t != ti.first_tile) {
				st_relaxed_u64(&P.st.agg_ent[(size_t)t * 64 + lane], pack_ent(epoch, own_val[0], own_ok[0]));
				st_relaxed_u64(&P.st.agg_ent[(size_t)t * 64 + lane + 32], pack_ent(epoch, own_val[1], own_ok[1]));
				if (lane < 3) {
					const uint32_t payload = lane == 0 ? own_lo : (lane == 1 ? own_hi : (uint32_t)own_lnr);
					st_relaxed_u64(&P.st.agg_sum[(size_t)t * QOI_SUM_WORDS + lane], pack_sum(epoch, payload));
				}
				table_lookback(P, t, ti.first_tile, epoch, lb);
			}
			Tin_val[lane] = lb.in_val[0];
			Tin_val[lane + 32] = lb.in_val[1];
			Tin_flag[lane] = lb.in_ok[0] ? 1 : 0;
			Tin_flag[lane + 32] = lb.in_ok[1] ? 1 : 0;
			const uint32_t inc_lo = __ballot_sync(QOI_FULL_MASK, own_ok[0] || lb.in_ok[0]);
			const uint32_t inc_hi = __ballot_sync(QOI_FULL_MASK, own_ok[1] || lb.in_ok[1]);
			st_relaxed_u64(&P.st.inc_ent[(size_t)t * 64 + lane], pack_ent(epoch, own_ok[0] ? own_val[0] : lb.in_val[0], own_ok[0] || lb.in_ok[0]));
			st_relaxed_u64(&P.st.inc_ent[(size_t)t * 64 + lane + 32], pack_ent(epoch, own_ok[1] ? own_val[1] : lb.in_val[1], own_ok[1] || lb.in_ok[1]));
			if (lane < 3) {
				const int32_t inc_lnr = (own_lnr != QOI_LNR_NONE) ? own_lnr : lb.lnr_in;
				const uint32_t payload = lane == 0 ? inc_lo : (lane == 1 ? inc_hi : (uint32_t)inc_lnr);
				st_relaxed_u64(&P.st.inc_sum[(size_t)t * QOI_SUM_WORDS + lane], pack_sum(epoch, payload));
			}
			if (lane == 0) wlnr_in[0] = lb.lnr_in; /* last non-run pixel index before the tile (image coordinates) */
			__syncwarp();
			bar_arrive(BAR_T_READY, QOI_N_CS);

			bar_sync(BAR_C_DONE, QOI_N_CS); /* byte counts of the compute warps are in wtot[] */
			/* ---- phase B: output offset look-back, 32 predecessors per probe (as in qoi_enc.cuh) ---- */
			uint32_t total = 0;
#pragma unroll
			for (int w = 0; w < QOI_ENC_WARPS; w++) total += wtot[w];
			uint32_t excl = QOI_HEADER_BYTES;
			if (t != ti.first_tile) {
				if (lane == 0) st_relaxed_u64(&P.st.byte_state[t], pack_bytes(epoch, QOI_ST_AGG, total));
				uint32_t sum = 0;
				int64_t base = (int64_t)t - 1;
				while (true) {
					const int64_t tp = base - lane;
					uint32_t state = QOI_ST_AGG, val = 0;
					if (tp >= (int64_t)ti.first_tile) {
						while (true) {
							const uint64_t st = ld_relaxed_u64(&P.st.byte_state[tp]);
							if ((uint32_t)(st >> 34) == epoch && ((st >> 32) & 3u) != QOI_ST_NONE) {
								state = (uint32_t)(st >> 32) & 3u;
								val = (uint32_t)st;
								break;
							}
							spin_pause();
						}
					}
					const uint32_t incm = __ballot_sync(QOI_FULL_MASK, state == QOI_ST_INC);
					if (incm) {
						const uint32_t first = (uint32_t)__ffs((int)incm) - 1u;
						sum += __reduce_add_sync(QOI_FULL_MASK, lane <= first ? val : 0u);
						break;
					}
					sum += __reduce_add_sync(QOI_FULL_MASK, val);
					base -= 32;
				}
				excl = sum;
			}
			if (lane == 0) {
				st_relaxed_u64(&P.st.byte_state[t], pack_bytes(epoch, QOI_ST_INC, excl + total));
				misc[misc_slot(PIPE ? (uint32_t)cur : 0u, MISC_TILE_BASE)] = excl;
				misc[misc_slot(PIPE ? (uint32_t)cur : 0u, MISC_TILE_TOTAL)] = total;
			}
			if (PIPE) {
				if (lane < QOI_ENC_WARPS) { /* where each warp's slab goes inside the tile */
					uint32_t pre = 0;
					for (uint32_t w = 0; w < lane; w++) pre += wtot[w];
					((uint32_t *)(smem + L::OFF_HITW))[(uint32_t)cur * QOI_ENC_WARPS + lane] = pre;
				}
				if (t == ti.first_tile && lane < QOI_HEADER_BYTES) { /* header, qoi.h:384-388 */
					const uint32_t k = lane;
					uint8_t b;
					if (k < 4) b = (uint8_t)("qoif"[k]);
					else if (k < 8) b = (uint8_t)(ti.width >> (8 * (7 - k)));
					else if (k < 12) b = (uint8_t)(ti.height >> (8 * (11 - k)));
					else if (k == 12) b = (uint8_t)(ti.chan_cs & 0xFFu);
					else b = (uint8_t)(ti.chan_cs >> 8);
					gstream[k] = b;
				}
				if (ti.px0 + ti.n_valid == ti.npx && lane >= 16 && lane < 16 + QOI_END_BYTES + 1) { /* end marker, length: qoi.h:480-484 */
					const uint32_t k = lane - 16;
					const uint32_t fin = excl + total;
					if (k < QOI_END_BYTES) gstream[fin + k] = (k == QOI_END_BYTES - 1) ? 1 : 0;
					else P.lens[ti.img] = (int32_t)(fin + QOI_END_BYTES);
				}
			}
			__syncwarp();
			bar_arrive(BAR_B_READY, QOI_N_CS);
		} else {
			/* =============================== compute warps =============================== */
			if (rb == 0) { mbar_wait(&mbar[0], parity0); parity0 ^= 1u; }
			else { mbar_wait(&mbar[1], parity1); parity1 ^= 1u; }

			uint2 *Lent = Lent_all + warp * 64;
			uint16_t *Fpos = Fpos_all + warp * 64;
			uint32_t *patchw = patch_all + warp * 16;
			if (!SLAB) Lent[lane] = make_uint2(0u, 0u);
			if (!SLAB) Lent[lane + 32] = make_uint2(0u, 0u);
			if (!SLAB) Fpos[lane] = 0xFFFFu;
			if (!SLAB) Fpos[lane + 32] = 0xFFFFu;
			if (!SLAB && lane < 16) patchw[lane] = 0u;
			if (EXPAND) {
				/* ---- phase E: 16 pixels per thread, raw bytes at data[sh + CH*i] -> RGBA words (as in qoi_enc.cuh) ---- */
				const uint32_t b0 = ti.sh + tid * (16u * CH);
				const uint32_t *wsrc = (const uint32_t *)data + (b0 >> 2);
				const uint32_t shf = (b0 & 3u) * 8u;
				uint32_t out[16];
				if (CH == 4) {
					uint32_t w[17];
#pragma unroll
					for (int j = 0; j < 17; j++) w[j] = wsrc[j];
#pragma unroll
					for (int j = 0; j < 16; j++) out[j] = __funnelshift_r(w[j], w[j + 1], shf);
				} else {
					uint32_t w[13];
#pragma unroll
					for (int j = 0; j < 13; j++) w[j] = wsrc[j];
					uint32_t s[12];
#pragma unroll
					for (int j = 0; j < 12; j++) s[j] = __funnelshift_r(w[j], w[j + 1], shf);
#pragma unroll
					for (int g = 0; g < 4; g++) {
						const uint32_t s0 = s[3 * g], s1 = s[3 * g + 1], s2 = s[3 * g + 2];
						out[4 * g + 0] = (s0 & 0x00FFFFFFu) | 0xFF000000u;
						out[4 * g + 1] = (s0 >> 24) | ((s1 & 0x0000FFFFu) << 8) | 0xFF000000u;
						out[4 * g + 2] = (s1 >> 16) | ((s2 & 0x000000FFu) << 16) | 0xFF000000u;
						out[4 * g + 3] = (s2 >> 8) | 0xFF000000u;
					}
				}
				uint4 *dst = (uint4 *)(pxv + tid * 16u);
				dst[0] = make_uint4(out[0], out[1], out[2], out[3]);
				dst[1] = make_uint4(out[4], out[5], out[6], out[7]);
				dst[2] = make_uint4(out[8], out[9], out[10], out[11]);
				dst[3] = make_uint4(out[12], out[13], out[14], out[15]);
				if (tid == 0) {
					const uint32_t *nbr = (const uint32_t *)(smem + L::OFF_NBR) + cur * 2;
					pxv[-1] = nbr[0];
					pxv[QOI_TILE_PX] = nbr[1];
				}
				bar_sync(BAR_COMPUTE, QOI_ENC_COMPUTE);
				bar_arrive(BAR_E_DONE, QOI_N_CS); /* the raw buffer may be refilled */
			}
			__syncwarp();

			if (SLAB) {
				slab_compute<CH, EXPAND, MATCH, PIPE, COMPACT>(P, smem, ti, pxv, gstream, (uint32_t)cur, carry);
			} else {
			const uint32_t nv = ti.n_valid; /* <= 4096; only an image's last tile can be partial */
			const bool last_tile = ti.px0 + ti.n_valid == ti.npx;
			const uint32_t lt = lanemask_lt(), gt = lanemask_gt();
			const uint32_t wbase = warp * (uint32_t)QOI_WARP_PX;

			/* ---- pass 0: 16 rows of 32 pixels, lane = pixel.  code[s] = size | b0 << 8 | b1 << 16 of my pixel in row s ---- */
			uint32_t code[QOI_WARP_STEPS];
			uint32_t nrmask = 0; /* bit s: my pixel of row s is a non-run pixel */
			uint32_t pend = 0;   /* bit s: my pixel of row s is a run pixel whose run started before this warp's chunk */
			int32_t last_nr = -1; /* tile-local index of the last non-run pixel of this chunk so far (uniform) */
#pragma unroll
			for (int s = 0; s < QOI_WARP_STEPS; s++) {
				const uint32_t rb0 = wbase + (uint32_t)s * 32u;
				uint32_t cd = 0;
				if (rb0 < nv) { /* uniform */
					const uint32_t il = rb0 + lane;
					const uint32_t v = pxv[il], p = pxv[(int)il - 1], nx = pxv[il + 1];
					const bool valid = il < nv;
					const bool nonrun = valid && v != p; /* qoi.h:415 */
					const uint32_t nr = __ballot_sync(QOI_FULL_MASK, nonrun);
					const uint32_t h = qoi_slot(v);
					uint32_t m; /* the non-run lanes of this row with my slot */
					if (MATCH) {
						m = __match_any_sync(QOI_FULL_MASK, nonrun ? h : 64u + lane);
					} else {
						uint32_t differ = 0; /* lanes whose slot differs from mine in some bit */
#pragma unroll
						for (int b = 0; b < 6; b++) {
							uint32_t plane, mine;
							ballot_and_mask(h & (1u << b), plane, mine);
							differ |= plane ^ mine;
						}
						m = nr & ~differ;
					}
					const uint32_t lower = m & lt;
					const int src = lower ? (31 - __clz((int)lower)) : (int)lane;
					const uint32_t pv = __shfl_sync(QOI_FULL_MASK, v, src);
					const uint2 ent = Lent[h];
					bool hit = false;
					if (nonrun) {
						if (lower) hit = (pv == v);          /* predecessor with my slot inside this row */
						else if (ent.y) hit = (ent.x == v);  /* ... in an earlier row of this chunk */
						else Fpos[h] = (uint16_t)il;         /* first pixel of the slot in the chunk: phase R decides */
					}
					__syncwarp();
					if (nonrun && (m & gt) == 0u) Lent[h] = make_uint2(v, 1u); /* last pixel of its slot in this row */
					if (nonrun) {
						cd = hit ? (1u | (h << 8)) : classify_bf(v, p); /* qoi.h:432-434 / 438-474 */
						nrmask |= 1u << s;
					} else if (valid) {
						const uint32_t below = nr & lt;
						const int32_t lnr_l = below ? (int32_t)(rb0 + 31u - (uint32_t)__clz((int)below)) : last_nr;
						if (lnr_l >= 0) cd = run_code(il - (uint32_t)lnr_l, (last_tile && ti.px0 + il == ti.npx - 1u) || nx != v);
						else pend |= 1u << s;
					}
					if (nr) last_nr = (int32_t)(rb0 + 31u - (uint32_t)__clz((int)nr));
					__syncwarp();
				}
				code[s] = cd;
			}
			if (lane == 0) wlnr[warp] = last_nr;
			__syncwarp();
			bar_arrive(BAR_I_DONE, QOI_N_CS);

			bar_sync(BAR_T_READY, QOI_N_CS); /* every warp's table is final; T_in and the run carry are in shared memory */
			/* ---- phase R: my chunk's deferred pixels against the earlier warps' tables, then the incoming table ---- */
#pragma unroll
			for (int q = 0; q < 2; q++) {
				const uint32_t sl = lane + 32u * q;
				const uint32_t pos = Fpos[sl];
				if (pos != 0xFFFFu) {
					const uint32_t pvv = pxv[pos];
					bool found = false;
					uint32_t val = 0;
					for (int w2 = (int)warp - 1; w2 >= 0; w2--) {
						const uint2 e = Lent_all[w2 * 64 + sl];
						if (e.y) {
							val = e.x;
							found = true;
							break;
						}
					}
					if (!found && Tin_flag[sl]) {
						val = Tin_val[sl];
						found = true;
					}
					/* an untouched index[] slot holds 0 (qoi.h:393): only the all-zero pixel can hit it */
					const bool hit = found ? (val == pvv) : (pvv == 0u);
					if (hit) atomicOr(&patchw[(pos >> 5) & 15u], 1u << (pos & 31u));
				}
			}
			__syncwarp();
			/* ---- patch the pixels that turned into INDEX ops; finish the runs that were carried into the chunk ---- */
			int32_t lnr_carry = wlnr_in[0]; /* image coordinates; -1: no non-run pixel since the image start */
			for (int w2 = (int)warp - 1; w2 >= 0; w2--) {
				if (wlnr[w2] >= 0) {
					lnr_carry = (int32_t)ti.px0 + wlnr[w2];
					break;
				}
			}
			const bool any_pend = __ballot_sync(QOI_FULL_MASK, pend != 0u) != 0u;
#pragma unroll
			for (int s = 0; s < QOI_WARP_STEPS; s++) {
				const uint32_t il = wbase + (uint32_t)s * 32u + lane;
				if ((patchw[s] >> lane) & 1u) code[s] = 1u | (qoi_slot(pxv[il]) << 8);
				if (any_pend && ((pend >> s) & 1u)) {
					const uint32_t gi = ti.px0 + il;
					code[s] = run_code(gi - (uint32_t)lnr_carry, (last_tile && gi == ti.npx - 1u) || pxv[il + 1] != pxv[il]);
				}
			}
			/* decoder restart point of the tile: the first pixel at which an op starts (within the first 62 pixels) */
			int32_t cand = INT_MAX;
			if (warp == 0) {
#pragma unroll
				for (int s = 0; s < 2; s++) {
					const uint32_t il = (uint32_t)s * 32u + lane;
					const bool valid = il < nv;
					const bool starts = valid && (((nrmask >> s) & 1u) || ((code[s] >> 8) & 0x3Fu) == 0u);
					const uint32_t b = __ballot_sync(QOI_FULL_MASK, starts);
					if (b && cand == INT_MAX) cand = (int32_t)((uint32_t)s * 32u + (uint32_t)__ffs((int)b) - 1u);
				}
			}
			/* ---- byte offsets: packed inclusive scan over the sizes of four rows at a time (a row holds <= 160 bytes) ---- */
			uint32_t ex4[4], tot4[4];
			uint32_t wsum = 0;
#pragma unroll
			for (int g = 0; g < 4; g++) {
				const uint32_t pk = (code[4 * g] & 7u) | ((code[4 * g + 1] & 7u) << 8) | ((code[4 * g + 2] & 7u) << 16) | ((code[4 * g + 3] & 7u) << 24);
				uint32_t incl = pk;
#pragma unroll
				for (int dlt = 1; dlt < 32; dlt <<= 1) {
					const uint32_t o = __shfl_up_sync(QOI_FULL_MASK, incl, dlt);
					if (lane >= (uint32_t)dlt) incl += o;
				}
				ex4[g] = incl - pk;
				tot4[g] = __shfl_sync(QOI_FULL_MASK, incl, 31);
				wsum = __dp4a(tot4[g], 0x01010101u, wsum);
			}
			if (lane == 0) wtot[warp] = wsum;
			__syncwarp();
			bar_arrive(BAR_C_DONE, QOI_N_CS);

			uint32_t tile_base = 0, padb = 0;
			if (EARLYW) {
				bar_sync(BAR_COMPUTE, QOI_ENC_COMPUTE); /* wtot[] of every warp is visible; every warp is past its last use of the staging buffer */
			} else {
				bar_sync(BAR_B_READY, QOI_N_CS);
				bar_sync(BAR_STAGE_FREE, QOI_N_CS); /* the previous tile's bulk store has drained the staging buffer */
				tile_base = misc[MISC_TILE_BASE];
				padb = (uint32_t)((uintptr_t)(gstream + tile_base) & 15u); /* staging congruent to the destination */
			}
			/* ---- phase W: op bytes -> staging ---- */
			uint32_t rowbase = 0;
			for (uint32_t w = 0; w < warp; w++) rowbase += wtot[w];
			uint32_t rs_off = 0xFFFFFFFFu; /* tile-relative byte offset of the restart op, in the lane that owns the restart pixel */
#pragma unroll
			for (int s = 0; s < QOI_WARP_STEPS; s++) {
				const uint32_t il = wbase + (uint32_t)s * 32u + lane;
				const uint32_t off = rowbase + ((ex4[s >> 2] >> (8 * (s & 3))) & 0xFFu);
				const uint32_t cd = code[s];
				const uint32_t sz = cd & 7u;
				unsigned char *o = stage + padb + off;
				if (s < 2 && warp == 0 && (int32_t)il == cand) rs_off = off;
				if (sz) {
					o[0] = (uint8_t)(cd >> 8);
					if (sz >= 2) {
						o[1] = (uint8_t)(cd >> 16);
						if (sz >= 4) {
							const uint32_t v = pxv[il];
							o[2] = (uint8_t)(v >> 8);
							o[3] = (uint8_t)(v >> 16);
							if (sz == 5) o[4] = (uint8_t)(v >> 24);
						}
					}
				}
				rowbase += (tot4[s >> 2] >> (8 * (s & 3))) & 0xFFu;
			}
			if (EARLYW) {
				bar_sync(BAR_B_READY, QOI_N_CS);
				tile_base = misc[MISC_TILE_BASE];
			}
			if (rs_off != 0xFFFFFFFFu) { /* the owner of the restart pixel records where its op starts */
				P.st.rs_px[t] = (int32_t)ti.px0 + cand;
				P.st.rs_byte[t] = (int32_t)(tile_base + rs_off);
				P.st.rs_prev[t] = pxv[cand - 1];
			}
			if (EARLYW) {
				bar_sync(BAR_COMPUTE, QOI_ENC_COMPUTE); /* the staging buffer is complete */
				const uint32_t total = misc[MISC_TILE_TOTAL];
				uint8_t *gdst = gstream + tile_base;
				uint32_t head = (16u - (uint32_t)((uintptr_t)gdst & 15u)) & 15u; /* bytes up to the first aligned 16 */
				if (head > total) head = total;
				const uint32_t nvec = (total - head) >> 4;
				if (tid < head) gdst[tid] = stage[tid];
				const uint32_t *s32 = (const uint32_t *)stage + (head >> 2);
				const uint32_t sh = (head & 3u) * 8u;
				for (uint32_t k = tid; k < nvec; k += QOI_ENC_COMPUTE) {
					const uint32_t *w = s32 + 4u * k;
					const uint32_t w0 = w[0], w1 = w[1], w2 = w[2], w3 = w[3], w4 = sh ? w[4] : 0u;
					uint4 q;
					q.x = __funnelshift_r(w0, w1, sh);
					q.y = __funnelshift_r(w1, w2, sh);
					q.z = __funnelshift_r(w2, w3, sh);
					q.w = __funnelshift_r(w3, w4, sh);
					*(uint4 *)(gdst + head + 16u * k) = q;
				}
				const uint32_t tail0 = head + 16u * nvec;
				if (tail0 + tid < total) gdst[tail0 + tid] = stage[tail0 + tid]; /* fewer than 16 bytes */
			}
			if (tid == 0 && cand == INT_MAX) P.st.rs_px[t] = -1;
			__syncwarp();
			bar_arrive(BAR_W_DONE, QOI_N_CS);
			} /* !SLAB */
		}
		cur ^= 1;
	}

	if (SLAB && PIPE && !loader && !scout && carry.valid) slab_finish<CH, EXPAND, true, COMPACT>(P, smem, carry); /* the last tile's copy-out */
	/* the last CTA to finish rearms the ticket for the next launch */
	__syncthreads();
	if (loader && lane == 0) { /* this thread issued the bulk stores */
		tma_store_wait_all();
		const uint32_t done = atomicAdd(&P.st.ticket[1], 1u);
		if (done == gridDim.x - 1u) {
			P.st.ticket[0] = 0;
			P.st.ticket[1] = 0;
		}
	}
}

}  // namespace qoienc4
