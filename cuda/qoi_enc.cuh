/*
 * qoi_enc.cuh -- single-pass chained-scan QOI encoder for sm_100a.
 *
 * Replaces the sequential loop of the reference encoder (qoi.h:406-478) for a
 * whole batch of images in one launch.  The per-pixel state machine is carried
 * by px_prev, a run counter, index[64] and the output cursor p; here each of
 * these becomes a scan over 4096-pixel tiles (order-free form):
 *
 *   run position      k(i) = i - (last non-run pixel before i)              max-scan
 *   INDEX hit         value of the previous NON-RUN pixel with the same
 *                     hash slot equals this pixel                           64-wide last-writer scan
 *   output offset     exclusive prefix sum of the bytes each pixel emits    +-scan
 *
 * Persistent CTAs take tiles from an atomic ticket (so every predecessor of a
 * running tile is itself running or finished) and publish / look back through
 * per-tile records in HBM (decoupled look-back, fence-free: every 64-bit record
 * word carries its launch epoch).  Pixels are read ONCE: a 1-D TMA bulk copy
 * (cp.async.bulk, SASS UBLKCP) brings the next tile into shared memory while the
 * current one is processed; op bytes are assembled in a shared staging buffer
 * whose alignment is congruent to the destination and leave with one TMA bulk
 * store plus <16-byte edges.
 *
 * Inside a tile the "previous pixel with my hash slot" is found through a
 * slot x position BIT MATRIX in shared memory (64 rows of 4096 bits): every
 * non-run pixel sets its bit, then looks backwards along its row for the nearest
 * set bit.  The gaps along one row sum to at most the tile length, so a tile
 * costs O(pixels + 64 * words) word reads however the slots are distributed.
 *
 * CTA = 8 compute warps (16 consecutive px per thread) + scout warp + loader warp.
 *
 *   compute warps                              scout warp               loader warp
 *   -------------                              ----------               -----------
 *   wait tile bytes (mbarrier)                                          ticket, TMA for the NEXT tile
 *   E  (EXPAND) raw bytes -> RGBA words
 *   A  hash + non-run flag per pixel, set bits
 *      in the slot x position matrix
 *        ------- I_DONE ------->               T  tile aggregate (last
 *   C0 per pixel: nearest earlier bit in its      writer per slot) published;
 *      row -> INDEX hit or not; RUN / DIFF /      32-tile window look-back for
 *      LUMA / RGB / RGBA exactly like              the incoming table and the last
 *      qoi.h:415-474.  A slot's first pixel        non-run index; inclusive record
 *      in the tile is deferred (<= 64)             published
 *        <------ T_READY -------
 *   R  deferred pixels against T_in
 *   C1 patch pixels that became hits; threads
 *      inside a run carried into the tile redo
 *      their 16 pixels; block scan of sizes
 *        ------- C_DONE ------->               B  output-offset look-back
 *        <------ B_READY -------                  (32 tiles per probe)
 *   W  op bytes -> staging; restart entry
 *        ------- W_DONE ----------------------------------------------> staging -> global (TMA store
 *        <------ STAGE_FREE -------------------------------------------  + edges), header, end marker
 */
#pragma once
#include <limits.h>

#include "qoi_common.h"
#include "qoi_dev.cuh"

namespace qoienc {
using namespace qoidev;

struct TileInfo {
	int32_t tile; /* -1: no more work */
	uint32_t img;
	uint32_t t_in_img;
	uint32_t n_valid;   /* pixels of this tile that exist */
	uint32_t px0;       /* image pixel index of the tile's first pixel */
	uint32_t npx;       /* pixels in the image */
	uint32_t first_tile;
	uint32_t sh;        /* (global source address) & 15, 0 on the zero-copy path */
	uint64_t stream_off;
	uint32_t width, height;
	uint32_t chan_cs;   /* channels | colorspace << 8 */
	uint32_t pad[3];
};

enum {
	BAR_COMPUTE = 1,    /* the 256 compute threads */
	BAR_I_DONE = 2,     /* compute -> scout */
	BAR_T_READY = 3,    /* scout -> compute */
	BAR_C_DONE = 4,     /* compute -> scout */
	BAR_B_READY = 5,    /* scout -> compute */
	BAR_W_DONE = 6,     /* compute -> loader */
	BAR_STAGE_FREE = 7, /* loader -> compute: the bulk store has drained the staging buffer (== bit matrix memory) */
	BAR_E_DONE = 8,     /* compute -> loader (EXPAND): the raw buffer has been consumed */
	BAR_INFO_READY = 9  /* loader -> compute + scout: the next tile's info record is in shared memory */
};
#define QOI_N_CS (QOI_ENC_COMPUTE + 32)      /* compute + one helper warp */
#define QOI_N_ALL QOI_ENC_THREADS

#define QOI_CELL_ROW 520 /* bytes per slot row: 4096 bits + 8 bytes of padding (spreads rows over the banks) */

template <int CH, bool EXPAND>
struct Layout {
	static constexpr int NBUF = EXPAND ? 1 : 2; /* EXPAND frees the raw buffer right after phase E */
	static constexpr int RAW_STRIDE = 16 + 16 + QOI_TILE_PX * CH + 16;
	static constexpr int OFF_RAW = 0;
	static constexpr int OFF_PXBUF = NBUF * RAW_STRIDE;
	static constexpr int PXBUF_BYTES = EXPAND ? (16 + QOI_TILE_PX * 4 + 16) : 0;
	static constexpr int OFF_CELLS = OFF_PXBUF + PXBUF_BYTES;         /* bit matrix; the staging buffer aliases it */
	static constexpr int CELLS_BYTES = 64 * QOI_CELL_ROW;
	static constexpr int STAGE_BYTES = 16 + QOI_TILE_PX * (CH + 1) + 16;
	static constexpr int OFF_SUMMARY = OFF_CELLS + CELLS_BYTES;       /* u64 [64]: bit g of row h = word g of the row is non-zero */
	static constexpr int OFF_TINVAL = OFF_SUMMARY + 64 * 8;           /* u32 [64] */
	static constexpr int OFF_HITW = OFF_TINVAL + 64 * 4;              /* u32 [128] */
	static constexpr int OFF_FPOS = OFF_HITW + 128 * 4;               /* u16 [64] */
	static constexpr int OFF_TINFLAG = OFF_FPOS + 64 * 2;             /* u8 [64] */
	static constexpr int OFF_MBAR = OFF_TINFLAG + 64;                 /* u64 [2] */
	static constexpr int OFF_INFO = OFF_MBAR + 16;                    /* TileInfo [2] */
	static constexpr int OFF_NBR = OFF_INFO + 2 * (int)sizeof(TileInfo); /* u32 [2][2] */
	static constexpr int OFF_WTOT = OFF_NBR + 16;                     /* u32 [8] */
	static constexpr int OFF_WLNR = OFF_WTOT + 32;                    /* i32 [8] */
	static constexpr int OFF_WLNRIN = OFF_WLNR + 32;                  /* i32 [8] */
	static constexpr int OFF_MISC = OFF_WLNRIN + 32;                  /* u32 [8] */
	static constexpr int TOTAL = OFF_MISC + 32;
	static_assert(STAGE_BYTES <= CELLS_BYTES, "staging must fit into the bit matrix it aliases");
	static_assert(sizeof(TileInfo) % 16 == 0, "TileInfo must keep 16-byte alignment of what follows");
	static_assert(OFF_CELLS % 16 == 0 && OFF_SUMMARY % 8 == 0 && OFF_TINVAL % 16 == 0 && OFF_MBAR % 8 == 0, "alignment");
};

/* misc[] slots */
enum { MISC_TILE_BASE = 0, MISC_RS_CAND = 1, MISC_TILE_TOTAL = 2 };

__device__ __forceinline__ uint64_t pack_bytes(uint32_t epoch, uint32_t state, uint32_t v) {
	return ((uint64_t)epoch << 34) | ((uint64_t)state << 32) | v;
}
__device__ __forceinline__ uint64_t pack_ent(uint32_t epoch, uint32_t value, bool written) {
	return ((uint64_t)value << 32) | (uint64_t)((epoch << 1) | (written ? 1u : 0u));
}
__device__ __forceinline__ uint64_t pack_sum(uint32_t epoch, uint32_t payload) { return ((uint64_t)payload << 32) | epoch; }
__device__ __forceinline__ int hi_bit64(uint64_t w) { /* index of the highest set bit, w != 0 */
	const uint32_t hi = (uint32_t)(w >> 32);
	return hi ? 63 - __clz((int)hi) : 31 - __clz((int)(uint32_t)w);
}

/* ---------------------------------------------------------------- tile fetch (loader warp) */
struct Fetch {
	uint32_t tile;
	bool any;
	QoiImgMeta m;
	uint32_t img;
};

/* step 1: take a ticket and read the image record (touches no shared memory) */
__device__ __forceinline__ void fetch_ticket(const QoiEncParams &P, Fetch &F) {
	const uint32_t lane = lane_id();
	uint32_t tile = 0;
	if (lane == 0) tile = atomicAdd(&P.st.ticket[0], 1u);
	tile = __shfl_sync(QOI_FULL_MASK, tile, 0);
	F.any = tile < P.n_tiles;
	F.tile = 0;
	if (F.any) {
		F.tile = P.tile_order[tile]; /* ascending inside every image, images interleaved */
		F.img = P.tile_img[F.tile];
		F.m = P.imgs[F.img];
	}
}

/* step 2: start the bulk copy into raw buffer `rb`, fill fragments / neighbours, publish the info record `ib` */
template <int CH, bool EXPAND>
__device__ __forceinline__ void fetch_issue(const QoiEncParams &P, unsigned char *smem, const Fetch &F, int rb, int ib) {
	using L = Layout<CH, EXPAND>;
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

/* ---------------------------------------------------------------- op classification of one non-run, non-hit pixel */
/* returns size | b0 << 8 | b1 << 16 ; reference qoi.h:438-474 */
__device__ __forceinline__ uint32_t classify(uint32_t v, uint32_t p) {
	const uint32_t d = __vsub4(v, p);                 /* per-channel difference mod 256 (qoi.h:439-441) */
	const uint32_t t = __vadd4(d, 0x00020202u);       /* dr+2, dg+2, db+2, da */
	if ((t & 0xFFFCFCFCu) == 0u) {                    /* alpha equal and all three in [-2,1] (qoi.h:446-451) */
		return 1u | ((0x40u | ((t & 3u) << 4) | (((t >> 8) & 3u) << 2) | ((t >> 16) & 3u)) << 8);
	}
	const uint32_t dg = (d >> 8) & 0xFFu;
	const uint32_t e = __vsub4(d, dg * 0x00010101u);  /* dr-dg, 0, db-dg, da-dg (qoi.h:443-444) */
	const uint32_t t2 = __vadd4(e, 0x00080008u);      /* dr-dg+8, ., db-dg+8, . */
	const uint32_t g32 = (dg + 32u) & 0xFFu;
	if ((d >> 24) == 0u && (t2 & 0x00F000F0u) == 0u && (g32 & 0xC0u) == 0u) { /* qoi.h:453-459 */
		return 2u | ((0x80u | g32) << 8) | ((((t2 & 0xFu) << 4) | ((t2 >> 16) & 0xFu)) << 16);
	}
	if ((d >> 24) == 0u) return 4u | (0xFEu << 8) | ((v & 0xFFu) << 16); /* qoi.h:461-465 */
	return 5u | (0xFFu << 8) | ((v & 0xFFu) << 16);                       /* qoi.h:468-473 */
}

/* ---------------------------------------------------------------- scout: table / run look-back over a 32-tile window */
struct LookBack {
	uint32_t in_val[2];
	bool in_ok[2];
	int32_t lnr_in;
};

__device__ __forceinline__ void table_lookback(const QoiEncParams &P, uint32_t t, uint32_t first_tile, uint32_t epoch, LookBack &R) {
	const uint32_t lane = lane_id();
	uint32_t need_lo = 0xFFFFFFFFu, need_hi = 0xFFFFFFFFu; /* slots still unresolved (uniform) */
	bool need_lnr = true;
	uint32_t found_tile[2] = {0, 0};
	bool found_inc[2] = {false, false};
	R.in_ok[0] = R.in_ok[1] = false;
	R.in_val[0] = R.in_val[1] = 0;
	R.lnr_in = -1;
	int64_t base = (int64_t)t - 1;
	while (true) {
		const int64_t tp = base - lane;
		uint32_t state = QOI_ST_NONE, vlo = 0, vhi = 0;
		int32_t lnr = -1;
		if (tp < (int64_t)first_tile) {
			state = QOI_ST_INC; /* before the image: nothing written, no non-run pixel yet */
		} else { /* one look, no waiting: an unpublished tile only matters if nothing nearer resolves the slot */
			const uint64_t *is = P.st.inc_sum + (size_t)tp * QOI_SUM_WORDS;
			const uint64_t *as = P.st.agg_sum + (size_t)tp * QOI_SUM_WORDS;
			uint64_t w0 = ld_relaxed_u64(is + 0), w1 = ld_relaxed_u64(is + 1), w2 = ld_relaxed_u64(is + 2);
			if ((uint32_t)w0 == epoch && (uint32_t)w1 == epoch && (uint32_t)w2 == epoch) {
				state = QOI_ST_INC;
			} else {
				w0 = ld_relaxed_u64(as + 0); w1 = ld_relaxed_u64(as + 1); w2 = ld_relaxed_u64(as + 2);
				if ((uint32_t)w0 == epoch && (uint32_t)w1 == epoch && (uint32_t)w2 == epoch) state = QOI_ST_AGG;
			}
			if (state != QOI_ST_NONE) {
				vlo = (uint32_t)(w0 >> 32); vhi = (uint32_t)(w1 >> 32); lnr = (int32_t)(uint32_t)(w2 >> 32);
			}
		}
		/* lanes [0, usable) hold published tiles; only they (up to the nearest inclusive record) may answer */
		const uint32_t known = __ballot_sync(QOI_FULL_MASK, state != QOI_ST_NONE);
		const uint32_t usable = (known == 0xFFFFFFFFu) ? 32u : (uint32_t)__ffs((int)~known) - 1u;
		const uint32_t umask = (usable == 32u) ? 0xFFFFFFFFu : ((1u << usable) - 1u);
		const uint32_t incm = __ballot_sync(QOI_FULL_MASK, state == QOI_ST_INC) & umask;
		const uint32_t first_inc = incm ? (uint32_t)__ffs((int)incm) - 1u : 32u;
		const bool part = lane < usable && lane <= first_inc;
		if (!part) {
			vlo = vhi = 0;
			lnr = QOI_LNR_NONE;
		}
		if (need_lnr) {
			const bool has = part && (state == QOI_ST_INC || lnr != QOI_LNR_NONE);
			const uint32_t hb = __ballot_sync(QOI_FULL_MASK, has);
			if (hb) {
				R.lnr_in = __shfl_sync(QOI_FULL_MASK, lnr, __ffs((int)hb) - 1);
				need_lnr = false;
			}
		}
#pragma unroll
		for (int half = 0; half < 2; half++) {
			const uint32_t need = half ? need_hi : need_lo;
			const uint32_t vm = half ? vhi : vlo;
			const uint32_t any = __reduce_or_sync(QOI_FULL_MASK, vm) & need;
			uint32_t todo = any;
			while (todo) {
				const uint32_t sb = (uint32_t)__ffs((int)todo) - 1u;
				todo &= todo - 1u;
				const uint32_t b = __ballot_sync(QOI_FULL_MASK, (vm >> sb) & 1u);
				const uint32_t src = (uint32_t)__ffs((int)b) - 1u; /* nearest tile that wrote the slot */
				if (lane == sb) {
					found_tile[half] = (uint32_t)(base - src);
					found_inc[half] = (src == first_inc);
					R.in_ok[half] = true;
				}
			}
			if (half) need_hi &= ~any;
			else need_lo &= ~any;
		}
		if (incm) break;                                  /* whatever is still unresolved was never written before this tile */
		if (!need_lnr && !(need_lo | need_hi)) break;     /* everything answered by nearer tiles */
		if (usable == 0u) spin_pause();                   /* the nearest undecided tile has not published yet */
		base -= usable;
	}
	/* the values themselves */
#pragma unroll
	for (int q = 0; q < 2; q++) {
		if (R.in_ok[q]) {
			const uint64_t *src = (found_inc[q] ? P.st.inc_ent : P.st.agg_ent) + (size_t)found_tile[q] * 64 + lane + 32u * q;
			uint64_t e;
			while (true) {
				e = ld_relaxed_u64(src);
				if (((uint32_t)e >> 1) == epoch) break;
				spin_pause();
			}
			R.in_val[q] = (uint32_t)(e >> 32);
		}
	}
}

/* ---------------------------------------------------------------- the kernel */
/* WALK: look-ups walk back along the slot row word by word instead of using the second-level bits (A/B switch) */
template <int CH, bool EXPAND, bool WALK>
__global__ void __launch_bounds__(QOI_ENC_THREADS, 3) qoi_encode_kernel(const QoiEncParams P) {
	using L = Layout<CH, EXPAND>;
	QOI_DYNAMIC_SMEM(smem);
	const uint32_t tid = threadIdx.x, lane = tid & 31u, warp = tid >> 5;
	const bool scout = warp == QOI_ENC_WARPS;
	const bool loader = warp == QOI_ENC_WARPS + 1;
	uint64_t *mbar = (uint64_t *)(smem + L::OFF_MBAR);
	TileInfo *infos = (TileInfo *)(smem + L::OFF_INFO);
	unsigned char *cells = smem + L::OFF_CELLS;
	uint64_t *summary = (uint64_t *)(smem + L::OFF_SUMMARY);
	unsigned char *stage = cells; /* alias: the matrix is dead once phase R is over */
	uint16_t *Fpos = (uint16_t *)(smem + L::OFF_FPOS);
	uint32_t *Tin_val = (uint32_t *)(smem + L::OFF_TINVAL);
	uint8_t *Tin_flag = (uint8_t *)(smem + L::OFF_TINFLAG);
	uint32_t *hitw = (uint32_t *)(smem + L::OFF_HITW);
	uint32_t *wtot = (uint32_t *)(smem + L::OFF_WTOT);
	int32_t *wlnr = (int32_t *)(smem + L::OFF_WLNR);
	int32_t *wlnr_in = (int32_t *)(smem + L::OFF_WLNRIN);
	uint32_t *misc = (uint32_t *)(smem + L::OFF_MISC);

	if (tid == 0) {
		mbar_init(&mbar[0], 1);
		mbar_init(&mbar[1], 1);
		mbar_fence_init();
		misc[MISC_RS_CAND] = (uint32_t)INT_MAX;
	}
	__syncthreads();
	if (loader) {
		Fetch F;
		fetch_ticket(P, F);
		fetch_issue<CH, EXPAND>(P, smem, F, 0, 0);
		__syncwarp();
		bar_arrive(BAR_INFO_READY, QOI_N_ALL);
		bar_arrive(BAR_STAGE_FREE, QOI_N_CS);
	}

	int cur = 0;
	uint32_t parity0 = 0, parity1 = 0;
	const uint32_t epoch = P.epoch;

	while (true) {
		if (!loader) bar_sync(BAR_INFO_READY, QOI_N_ALL); /* infos[cur] is published */
		const TileInfo ti = infos[cur];
		if (ti.tile < 0) break;
		const uint32_t t = (uint32_t)ti.tile;
		const int rb = EXPAND ? 0 : cur;
		unsigned char *data = smem + L::OFF_RAW + rb * L::RAW_STRIDE + 16;
		uint32_t *pxv = EXPAND ? (uint32_t *)(smem + L::OFF_PXBUF + 16) : (uint32_t *)data; /* pxv[-1 .. 4096] */
		uint8_t *gstream = P.streams + ti.stream_off;

		if (loader) {
			/* =============================== loader warp =============================== */
			Fetch F;
			fetch_ticket(P, F);
			if (EXPAND) bar_sync(BAR_E_DONE, QOI_N_CS); /* single raw buffer: wait until phase E has drained it */
			fetch_issue<CH, EXPAND>(P, smem, F, EXPAND ? 0 : (cur ^ 1), cur ^ 1);

			bar_sync(BAR_W_DONE, QOI_N_CS); /* staging is complete */
			/* Only now may INFO_READY be signalled again: a compute thread arrives at W_DONE, and the scout at the
			   B_READY before it, only after passing this iteration's INFO_READY, so the barrier's previous phase is
			   over.  (Signalling right after fetch_issue could overtake a warp that had not reached its
			   bar.sync yet and complete that phase early.) */
			__syncwarp();
			bar_arrive(BAR_INFO_READY, QOI_N_ALL); /* consumed at the top of the next iteration */
			{
				const uint32_t excl = misc[MISC_TILE_BASE], total = misc[MISC_TILE_TOTAL];
				uint8_t *gdst = gstream + excl;
				const uint32_t padb = (uint32_t)((uintptr_t)gdst & 15u);
				const uint32_t lo_off = (padb + 15u) & ~15u;
				const uint32_t end = padb + total;
				const uint32_t hi_off = end & ~15u;
				const bool bulk = hi_off > lo_off;
				if (lane == 0 && bulk) {
					fence_proxy_async_smem();
					tma_store_1d_hint(gdst + (lo_off - padb), stage + lo_off, hi_off - lo_off, l2_evict_first_policy());
					tma_store_commit();
				}
				const uint32_t a_hi = bulk ? lo_off : end; /* [padb, a_hi) */
				const uint32_t b_lo = bulk ? hi_off : end; /* [b_lo, end)  */
				for (uint32_t o = padb + lane; o < a_hi; o += 32) gdst[o - padb] = stage[o];
				for (uint32_t o = b_lo + lane; o < end; o += 32) gdst[o - padb] = stage[o];
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
				if (lane == 0) {
					misc[MISC_RS_CAND] = (uint32_t)INT_MAX; /* for the next tile */
					tma_store_wait_read();                  /* the bulk store has finished reading the staging buffer */
				}
			}
			__syncwarp();
			bar_arrive(BAR_STAGE_FREE, QOI_N_CS);
		} else if (scout) {
			/* =============================== scout warp =============================== */
			bar_sync(BAR_I_DONE, QOI_N_CS); /* the bit matrix and wlnr[] of this tile are complete */
			/* ---- phase T: tile aggregate = last set bit of each slot row (lane handles slots lane, lane+32) ---- */
			uint32_t own_val[2];
			bool own_ok[2];
#pragma unroll
			for (int q = 0; q < 2; q++) {
				const uint32_t sl = lane + 32u * q;
				const uint64_t *row = (const uint64_t *)(cells + sl * QOI_CELL_ROW);
				own_ok[q] = false;
				own_val[q] = 0;
				if (WALK) {
					for (int g = 63; g >= 0; g--) {
						const uint64_t w = row[g];
						if (w) {
							own_val[q] = pxv[g * 64 + hi_bit64(w)];
							own_ok[q] = true;
							break;
						}
					}
				} else {
					const uint64_t sm = summary[sl];
					if (sm) {
						const int g = hi_bit64(sm);
						own_val[q] = pxv[g * 64 + hi_bit64(row[g])];
						own_ok[q] = true;
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
			/* inclusive record = what qoi.h's index[] / run bookkeeping hold after this tile */
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
			/* ---- phase B: output offset look-back, 32 predecessors per probe ---- */
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
				misc[MISC_TILE_BASE] = excl;
				misc[MISC_TILE_TOTAL] = total;
			}
			__syncwarp();
			bar_arrive(BAR_B_READY, QOI_N_CS);
		} else {
			/* =============================== compute warps =============================== */
			if (rb == 0) { mbar_wait(&mbar[0], parity0); parity0 ^= 1u; }
			else { mbar_wait(&mbar[1], parity1); parity1 ^= 1u; }

			/* clear the bit matrix (it aliases the staging buffer of the previous tile) and the per-tile side tables */
			bar_sync(BAR_STAGE_FREE, QOI_N_CS);
			for (uint32_t k = tid; k < (uint32_t)(L::CELLS_BYTES / 16); k += QOI_ENC_COMPUTE) ((uint4 *)cells)[k] = make_uint4(0u, 0u, 0u, 0u);
			if (tid < 128) hitw[tid] = 0u;
			if (tid < 64) Fpos[tid] = 0xFFFFu;
			if (EXPAND) {
				/* ---- phase E: 16 pixels per thread, raw bytes at data[sh + CH*i] -> RGBA words ---- */
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
			}
			bar_sync(BAR_COMPUTE, QOI_ENC_COMPUTE);
			if (EXPAND) bar_arrive(BAR_E_DONE, QOI_N_CS); /* the raw buffer may be refilled */

			/* ---- phase A: my 16 consecutive pixels: non-run flags, bits into the slot x position matrix ---- */
			uint32_t v[16];
			uint32_t code[16];
			uint32_t ttotal = 0;
			const uint32_t i0 = tid * 16u;
			uint32_t pprev, pnext;
			{
				const uint4 *src = (const uint4 *)(pxv + i0);
				const uint4 q0 = src[0], q1 = src[1], q2 = src[2], q3 = src[3];
				v[0] = q0.x; v[1] = q0.y; v[2] = q0.z; v[3] = q0.w;
				v[4] = q1.x; v[5] = q1.y; v[6] = q1.z; v[7] = q1.w;
				v[8] = q2.x; v[9] = q2.y; v[10] = q2.z; v[11] = q2.w;
				v[12] = q3.x; v[13] = q3.y; v[14] = q3.z; v[15] = q3.w;
				pprev = pxv[(int)i0 - 1];
				pnext = pxv[i0 + 16];
			}
			const bool full = ti.n_valid == (uint32_t)QOI_TILE_PX; /* uniform: only an image's last tile can be partial */
			const bool last_tile = ti.px0 + ti.n_valid == ti.npx;
			const bool track_cand = tid < 4;                        /* the restart point lies in the first 62 pixels */
			uint32_t nrbits = 0;
			{
				uint32_t before = pprev;
#pragma unroll
				for (int j = 0; j < 16; j++) {
					if ((full || i0 + j < ti.n_valid) && v[j] != before) {
						uint16_t *cell = (uint16_t *)(cells + qoi_slot(v[j]) * QOI_CELL_ROW) + tid;
						*cell = (uint16_t)(*cell | (1u << j));
						nrbits |= 1u << j;
					}
					before = v[j];
				}
			}
			/* last non-run pixel (tile coordinates) before my first pixel: warp max-scan + per-warp totals */
			int32_t lnr_local;
			{
				const int32_t mine = nrbits ? (int32_t)(i0 + 31u - (uint32_t)__clz((int)nrbits)) : -1;
				int32_t incl = mine;
#pragma unroll
				for (int dlt = 1; dlt < 32; dlt <<= 1) {
					const int32_t o = __shfl_up_sync(QOI_FULL_MASK, incl, dlt);
					if (lane >= (uint32_t)dlt && o > incl) incl = o;
				}
				lnr_local = __shfl_up_sync(QOI_FULL_MASK, incl, 1);
				if (lane == 0) lnr_local = -1;
				if (lane == 31) wlnr[warp] = incl;
			}
			bar_sync(BAR_COMPUTE, QOI_ENC_COMPUTE); /* matrix and wlnr[] complete */
			/* second level: one bit per 64-pixel word of every row, so that a look-up never has to walk */
#pragma unroll
			for (int r = 0; r < (WALK ? 0 : 8); r++) {
				const uint64_t *row = (const uint64_t *)(cells + (warp * 8u + r) * QOI_CELL_ROW);
				const uint32_t b0 = __ballot_sync(QOI_FULL_MASK, row[lane] != 0ull);
				const uint32_t b1 = __ballot_sync(QOI_FULL_MASK, row[lane + 32] != 0ull);
				if (lane == 0) summary[warp * 8u + r] = (uint64_t)b0 | ((uint64_t)b1 << 32);
			}
			if (!WALK) bar_sync(BAR_COMPUTE, QOI_ENC_COMPUTE);
			bar_arrive(BAR_I_DONE, QOI_N_CS);
			if (lnr_local < 0) {
				for (int w = (int)warp - 1; w >= 0; w--) {
					if (wlnr[w] >= 0) {
						lnr_local = wlnr[w];
						break;
					}
				}
			}
			/* a thread that starts inside a run which began before this tile needs the carry from the look-back */
			const bool defer = (full || i0 < ti.n_valid) && !(nrbits & 1u) && lnr_local < 0;
			uint32_t hitbits = 0;
			int32_t cand = INT_MAX;
			const uint32_t grp = tid >> 2, bitbase = (tid & 3u) * 16u;

			/* ---- phase C: classify pass; pass 0 with what the tile knows, pass 1 after the look-back ---- */
#pragma unroll 1
			for (int pass = 0; pass < 2; pass++) {
				int32_t lnr_glob = (lnr_local >= 0) ? (int32_t)ti.px0 + lnr_local : (int32_t)(ti.px0 + i0) - 1;
				if (pass == 1) {
					bar_sync(BAR_T_READY, QOI_N_CS); /* T_in and the run carry are in shared memory */
					/* ---- phase R: a slot's first pixel in the tile against the incoming table ---- */
					if (tid < 64) {
						const uint32_t pos = Fpos[tid];
						if (pos != 0xFFFFu) {
							const uint32_t pvv = pxv[pos];
							/* an untouched index[] slot holds 0 (qoi.h:393): only the all-zero pixel can hit it */
							const bool hit = Tin_flag[tid] ? (Tin_val[tid] == pvv) : (pvv == 0u);
							if (hit) atomicOr(&hitw[pos >> 5], 1u << (pos & 31u));
						}
					}
					bar_sync(BAR_COMPUTE, QOI_ENC_COMPUTE);
					const uint32_t changed = (hitw[i0 >> 5] >> (i0 & 31u)) & 0xFFFFu; /* pixels that turned into hits */
					hitbits |= changed;
					if (!defer) {
						/* patch the pixels that turned into INDEX ops (qoi.h:432-434) */
#pragma unroll
						for (int j = 0; j < 16; j++) {
							if ((changed >> j) & 1u) {
								ttotal -= (code[j] & 7u) - 1u;
								code[j] = 1u | (qoi_slot(v[j]) << 8);
							}
						}
						break;
					}
					lnr_glob = wlnr_in[0];
				}
				/* run pixels between the last non-run pixel and my first pixel */
				const uint32_t c = (uint32_t)((int32_t)(ti.px0 + i0) - 1 - lnr_glob);
				int32_t r = (c == 0u) ? -1 : (int32_t)((c - 1u) % 62u); /* (k-1) % 62 of the previous pixel */
				uint32_t before = pprev;
				ttotal = 0;
				cand = INT_MAX;
#pragma unroll
				for (int j = 0; j < 16; j++) {
					const uint32_t gi = ti.px0 + i0 + j;
					uint32_t cd = 0;
					if (full || i0 + j < ti.n_valid) {
						if ((nrbits >> j) & 1u) {
							r = -1;
							if (track_cand && cand == INT_MAX) cand = (int32_t)(i0 + j);
							const uint32_t h = qoi_slot(v[j]);
							bool hit;
							if (pass == 0) {
								/* nearest earlier non-run pixel with my slot: my own 64-pixel word first, else the nearest
								   earlier non-empty word of row h (second-level bits) */
								const uint64_t *row = (const uint64_t *)(cells + h * QOI_CELL_ROW);
								int g = (int)grp;
								uint64_t w = row[g] & ((1ull << (bitbase + j)) - 1ull);
								if (WALK) {
									while (w == 0ull && g > 0) {
										g--;
										w = row[g];
									}
								} else if (w == 0ull) {
									const uint64_t sm = summary[h] & ((1ull << grp) - 1ull);
									if (sm) {
										g = hi_bit64(sm);
										w = row[g];
									}
								}
								if (w) {
									hit = pxv[g * 64 + hi_bit64(w)] == v[j];
								} else {
									Fpos[h] = (uint16_t)(i0 + j); /* first pixel of this slot in the tile: phase R decides */
									hit = false;
								}
								hitbits |= (hit ? 1u : 0u) << j;
							} else {
								hit = (hitbits >> j) & 1u;
							}
							if (hit) cd = 1u | (h << 8); /* qoi.h:432-434 */
							else cd = classify(v[j], before);
						} else {
							/* run pixel (qoi.h:415-421): a byte leaves at run length 62, at the image end, or when the run ends */
							r = (r == 61) ? 0 : r + 1;
							if (track_cand && r == 0 && cand == INT_MAX) cand = (int32_t)(i0 + j);
							const uint32_t nx = (j == 15) ? pnext : v[j < 15 ? j + 1 : 15];
							if (r == 61 || (last_tile && gi == ti.npx - 1u) || nx != v[j]) cd = 1u | ((0xC0u | (uint32_t)r) << 8);
						}
					}
					code[j] = cd;
					ttotal += cd & 7u;
					before = v[j];
				}
			}
			if (track_cand && cand != INT_MAX) atomicMin((int *)&misc[MISC_RS_CAND], cand);
			/* block scan of the per-thread byte counts */
			uint32_t incl = ttotal;
#pragma unroll
			for (int dlt = 1; dlt < 32; dlt <<= 1) {
				const uint32_t o = __shfl_up_sync(QOI_FULL_MASK, incl, dlt);
				if (lane >= (uint32_t)dlt) incl += o;
			}
			if (lane == 31) wtot[warp] = incl;
			__syncwarp();
			bar_arrive(BAR_C_DONE, QOI_N_CS);

			bar_sync(BAR_B_READY, QOI_N_CS);
			/* ---- phase W: op bytes -> staging (alignment congruent to the destination) ---- */
			const uint32_t tile_base = misc[MISC_TILE_BASE];
			const int32_t rs_cand = (int32_t)misc[MISC_RS_CAND];
			const uint32_t padb = (uint32_t)((uintptr_t)(gstream + tile_base) & 15u);
			{
				uint32_t woff = incl - ttotal; /* exclusive inside the warp */
				for (uint32_t w = 0; w < warp; w++) woff += wtot[w];
				unsigned char *o = stage + padb + woff;
				if (track_cand) { /* decoder restart point of this tile: the owner of pixel rs_cand records where its op starts */
					const int32_t jr = rs_cand - (int32_t)i0;
					if (jr >= 0 && jr < 16) {
						uint32_t off = woff, before = pprev;
#pragma unroll
						for (int j = 0; j < 16; j++) {
							if (j < jr) {
								off += code[j] & 7u;
								before = v[j];
							}
						}
						P.st.rs_px[t] = (int32_t)(ti.px0 + (uint32_t)rs_cand);
						P.st.rs_byte[t] = (int32_t)(tile_base + off);
						P.st.rs_prev[t] = before;
					}
					if (tid == 0 && rs_cand == INT_MAX) P.st.rs_px[t] = -1;
				}
#pragma unroll
				for (int j = 0; j < 16; j++) {
					const uint32_t cd = code[j];
					const uint32_t sz = cd & 7u;
					if (sz) {
						o[0] = (uint8_t)(cd >> 8);
						if (sz >= 2) {
							o[1] = (uint8_t)(cd >> 16);
							if (sz >= 4) {
								o[2] = (uint8_t)(v[j] >> 8);
								o[3] = (uint8_t)(v[j] >> 16);
								if (sz == 5) o[4] = (uint8_t)(v[j] >> 24);
							}
						}
						o += sz;
					}
				}
			}
			__syncwarp();
			bar_arrive(BAR_W_DONE, QOI_N_CS);
		}
		cur ^= 1;
	}

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

}  // namespace qoienc
