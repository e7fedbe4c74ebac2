/*
 * qoi_enc.cuh -- single-pass chained-scan QOI encoder for sm_100a.
 *
 * Replaces the sequential loop of the reference encoder (qoi.h:406-478) for a
 * whole batch of images in one launch.  The per-pixel state machine is carried
 * by px_prev, a run counter, index[64] and the output cursor p; here each of
 * these becomes a scan over 4096-pixel tiles (order-free form):
 *
 *   run position      k(i) = i - (last non-run pixel before i)         max-scan
 *   INDEX hit         value of the previous NON-RUN pixel with the same
 *                     hash slot equals this pixel (64 "last writer" lanes) 64-wide last-writer scan
 *   output offset     exclusive prefix sum of the bytes each pixel emits  +-scan
 *
 * Persistent CTAs take tiles from an atomic ticket (so every predecessor of a
 * running tile is itself running or finished) and publish / look back through
 * per-tile records in HBM (decoupled look-back).  Pixels are read ONCE: a 1-D
 * TMA bulk copy (cp.async.bulk, SASS UBLKCP) brings the next tile into shared
 * memory while the current one is processed; op bytes are assembled in a shared
 * staging buffer whose alignment is congruent to the destination, and leave
 * with one TMA bulk store plus <16-byte edges.
 *
 * Phases per tile (256 threads, 8 warps, 512 px per warp, 16 px per thread):
 *   E  (EXPAND only) raw RGB / unaligned RGBA bytes -> aligned RGBA words
 *   I  "index pass", striped: 32 consecutive pixels per step across the lanes;
 *      __match_any_sync on the hash slot finds the in-step predecessor, a
 *      per-warp 64-entry table carries across steps; first occurrences of a
 *      slot in the warp chunk are deferred (at most 64 per warp)
 *   T  warp 0: tile aggregate (last writer per slot, last non-run index) is
 *      published, predecessors are looked back, inclusive record published
 *   R  deferred first occurrences are resolved against earlier warps / T_in
 *   C  "classify pass", blocked: each thread walks its 16 consecutive pixels,
 *      decides RUN / INDEX / DIFF / LUMA / RGB / RGBA exactly like qoi.h:415-474,
 *      block-scans the byte counts
 *   B  warp 0: byte-offset look-back (32 predecessors per probe)
 *   W  bytes -> staging -> global; header / end marker / length / restart entry
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

template <int CH, bool EXPAND>
struct Layout {
	static constexpr int RAW_STRIDE = 16 + 16 + QOI_TILE_PX * CH + 16;
	static constexpr int OFF_RAW = 0;
	static constexpr int OFF_PXBUF = 2 * RAW_STRIDE;
	static constexpr int PXBUF_BYTES = EXPAND ? (16 + QOI_TILE_PX * 4 + 16) : 0;
	static constexpr int OFF_STAGE = OFF_PXBUF + PXBUF_BYTES;
	static constexpr int STAGE_BYTES = 16 + QOI_TILE_PX * (CH + 1) + 16;
	static constexpr int OFF_LVAL = OFF_STAGE + STAGE_BYTES;          /* u32 [8][64] */
	static constexpr int OFF_TINVAL = OFF_LVAL + 8 * 64 * 4;          /* u32 [64] */
	static constexpr int OFF_HITW = OFF_TINVAL + 64 * 4;              /* u32 [128] */
	static constexpr int OFF_NRW = OFF_HITW + 128 * 4;                /* u32 [128] */
	static constexpr int OFF_LNRS = OFF_NRW + 128 * 4;                /* i32 [128] */
	static constexpr int OFF_FPOS = OFF_LNRS + 128 * 4;               /* u16 [8][64] */
	static constexpr int OFF_LFLAG = OFF_FPOS + 8 * 64 * 2;           /* u8 [8][64] */
	static constexpr int OFF_TINFLAG = OFF_LFLAG + 8 * 64;            /* u8 [64] */
	static constexpr int OFF_MBAR = OFF_TINFLAG + 64;                 /* u64 [2] */
	static constexpr int OFF_INFO = OFF_MBAR + 16;                    /* TileInfo [2] */
	static constexpr int OFF_NBR = OFF_INFO + 2 * (int)sizeof(TileInfo); /* u32 [2][2] */
	static constexpr int OFF_WTOT = OFF_NBR + 16;                     /* u32 [8] */
	static constexpr int OFF_WLNR = OFF_WTOT + 32;                    /* i32 [8] */
	static constexpr int OFF_WLNRIN = OFF_WLNR + 32;                  /* i32 [8] */
	static constexpr int OFF_MISC = OFF_WLNRIN + 32;                  /* u32 [8] */
	static constexpr int TOTAL = OFF_MISC + 32;
	static_assert(sizeof(TileInfo) % 16 == 0, "TileInfo must keep 16-byte alignment of what follows");
	static_assert(OFF_STAGE % 16 == 0 && OFF_LVAL % 16 == 0 && OFF_MBAR % 8 == 0, "alignment");
};

/* misc[] slots */
enum { MISC_TILE_BASE = 0, MISC_RS_CAND = 1, MISC_TILE_TOTAL = 2 };

__device__ __forceinline__ uint64_t pack_bytes(uint32_t epoch, uint32_t state, uint32_t v) {
	return ((uint64_t)epoch << 34) | ((uint64_t)state << 32) | v;
}

/* ---------------------------------------------------------------- tile fetch (warp 0) */
template <int CH, bool EXPAND>
__device__ __forceinline__ void fetch_tile(const QoiEncParams &P, unsigned char *smem, int b) {
	using L = Layout<CH, EXPAND>;
	const uint32_t lane = lane_id();
	TileInfo *info = (TileInfo *)(smem + L::OFF_INFO) + b;
	uint64_t *mbar = (uint64_t *)(smem + L::OFF_MBAR) + b;
	uint32_t tile = 0;
	if (lane == 0) tile = atomicAdd(&P.st.ticket[0], 1u);
	tile = __shfl_sync(QOI_FULL_MASK, tile, 0);
	if (tile >= P.n_tiles) {
		if (lane == 0) info->tile = -1;
		return;
	}
	/* image that owns the tile: largest i with first_tile[i] <= tile */
	uint32_t lo = 0, hi = P.n_images - 1;
	while (lo < hi) {
		uint32_t mid = (lo + hi + 1) >> 1;
		if (P.imgs[mid].first_tile <= tile) lo = mid;
		else hi = mid - 1;
	}
	const QoiImgMeta m = P.imgs[lo];
	const uint32_t t_in = tile - m.first_tile;
	const uint32_t px0 = t_in * QOI_TILE_PX;
	const uint32_t n_valid = (m.npx - px0) < (uint32_t)QOI_TILE_PX ? (m.npx - px0) : (uint32_t)QOI_TILE_PX;
	const uint8_t *gsrc = P.pixels + m.pix_off + (uint64_t)px0 * CH;
	const uint32_t nbytes = n_valid * CH;
	const uint32_t sh = EXPAND ? (uint32_t)((uintptr_t)gsrc & 15u) : 0u;
	unsigned char *data = smem + L::OFF_RAW + b * L::RAW_STRIDE + 16; /* data[sh + k] = gsrc[k] */
	const uint32_t lo_off = (sh + 15u) & ~15u;
	const uint32_t hi_off = (sh + nbytes) & ~15u;
	const uint32_t tma_bytes = hi_off > lo_off ? hi_off - lo_off : 0u;
	if (lane == 0) {
		mbar_arrive_expect_tx(mbar, tma_bytes);
		if (tma_bytes) tma_load_1d(data + lo_off, gsrc + (lo_off - sh), tma_bytes, mbar);
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
			nb = (uint32_t)q[0] | ((uint32_t)q[1] << 8) | ((uint32_t)q[2] << 16) | (CH == 4 ? ((uint32_t)q[3] << 24) : 0xFF000000u);
		}
	} else if (lane == 1) {
		if (px0 + n_valid < m.npx) {
			const uint8_t *q = gsrc + nbytes;
			nb = (uint32_t)q[0] | ((uint32_t)q[1] << 8) | ((uint32_t)q[2] << 16) | (CH == 4 ? ((uint32_t)q[3] << 24) : 0xFF000000u);
		}
	}
	if (lane < 2) {
		if (EXPAND) ((uint32_t *)(smem + L::OFF_NBR))[b * 2 + lane] = nb;
		else ((uint32_t *)data)[lane == 0 ? -1 : QOI_TILE_PX] = nb;
	}
	if (lane == 0) {
		info->tile = (int32_t)tile;
		info->img = lo;
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

/* ---------------------------------------------------------------- the kernel */
template <int CH, bool EXPAND>
__global__ void __launch_bounds__(QOI_ENC_THREADS, 3) qoi_encode_kernel(const QoiEncParams P) {
	using L = Layout<CH, EXPAND>;
	QOI_DYNAMIC_SMEM(smem);
	const uint32_t tid = threadIdx.x, lane = tid & 31u, warp = tid >> 5;
	uint64_t *mbar = (uint64_t *)(smem + L::OFF_MBAR);
	TileInfo *infos = (TileInfo *)(smem + L::OFF_INFO);
	uint32_t *Lval_all = (uint32_t *)(smem + L::OFF_LVAL);
	uint8_t *Lflag_all = (uint8_t *)(smem + L::OFF_LFLAG);
	uint16_t *Fpos_all = (uint16_t *)(smem + L::OFF_FPOS);
	uint32_t *Tin_val = (uint32_t *)(smem + L::OFF_TINVAL);
	uint8_t *Tin_flag = (uint8_t *)(smem + L::OFF_TINFLAG);
	uint32_t *hitw = (uint32_t *)(smem + L::OFF_HITW);
	uint32_t *nrw = (uint32_t *)(smem + L::OFF_NRW);
	int32_t *lnrs = (int32_t *)(smem + L::OFF_LNRS);
	uint32_t *wtot = (uint32_t *)(smem + L::OFF_WTOT);
	int32_t *wlnr = (int32_t *)(smem + L::OFF_WLNR);
	int32_t *wlnr_in = (int32_t *)(smem + L::OFF_WLNRIN);
	uint32_t *misc = (uint32_t *)(smem + L::OFF_MISC);
	unsigned char *stage = smem + L::OFF_STAGE;

	if (tid == 0) {
		mbar_init(&mbar[0], 1);
		mbar_init(&mbar[1], 1);
		mbar_fence_init();
	}
	__syncthreads();
	if (warp == 0) fetch_tile<CH, EXPAND>(P, smem, 0);
	__syncthreads();

	int cur = 0;
	uint32_t parity0 = 0, parity1 = 0;
	const uint32_t epoch = P.epoch;

	while (true) {
		const TileInfo ti = infos[cur];
		if (ti.tile < 0) break;
		const uint32_t t = (uint32_t)ti.tile;
		if (warp == 0) fetch_tile<CH, EXPAND>(P, smem, cur ^ 1); /* prefetch: overlaps everything below */

		/* wait for this tile's bulk copy */
		if (cur == 0) { mbar_wait(&mbar[0], parity0); parity0 ^= 1u; }
		else { mbar_wait(&mbar[1], parity1); parity1 ^= 1u; }

		unsigned char *data = smem + L::OFF_RAW + cur * L::RAW_STRIDE + 16;
		uint32_t *pxv; /* pxv[-1] = pixel before the tile, pxv[0..4095], pxv[4096] = pixel after */
		if (EXPAND) {
			/* ---- phase E: 16 pixels per thread, raw bytes at data[sh + CH*i] -> RGBA words ---- */
			pxv = (uint32_t *)(smem + L::OFF_PXBUF + 16);
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
			__syncthreads();
		} else {
			pxv = (uint32_t *)data;
		}

		/* ---- phase I: striped index pass, one warp per 512-pixel chunk ---- */
		{
			uint32_t *Lval = Lval_all + warp * 64;
			uint8_t *Lflag = Lflag_all + warp * 64;
			uint16_t *Fpos = Fpos_all + warp * 64;
			Lflag[lane] = 0;
			Lflag[lane + 32] = 0;
			Fpos[lane] = 0xFFFFu;
			Fpos[lane + 32] = 0xFFFFu;
			if (tid == 0) misc[MISC_RS_CAND] = (uint32_t)INT_MAX;
			__syncwarp();
			int32_t last_nr = -1; /* tile-local index of the last non-run pixel seen so far in this chunk */
			const uint32_t lt = lanemask_lt(), gt = lanemask_gt();
#pragma unroll 1
			for (int s = 0; s < QOI_WARP_STEPS; s++) {
				const uint32_t i = warp * QOI_WARP_PX + s * 32 + lane;
				const uint32_t v = pxv[i];
				const uint32_t p = pxv[(int)i - 1];
				const bool nonrun = (i < ti.n_valid) && (v != p);
				const uint32_t h = qoi_slot(v);
				const uint32_t m = __match_any_sync(QOI_FULL_MASK, nonrun ? h : (64u + lane));
				const uint32_t nr = __ballot_sync(QOI_FULL_MASK, nonrun);
				const uint32_t lower = m & lt;
				const int src = lower ? (31 - __clz((int)lower)) : (int)lane;
				const uint32_t pv = __shfl_sync(QOI_FULL_MASK, v, src);
				bool hit = false;
				if (nonrun) {
					if (lower) hit = (pv == v);             /* predecessor with the same slot inside this step */
					else if (Lflag[h]) hit = (Lval[h] == v); /* ... inside this warp chunk */
					else Fpos[h] = (uint16_t)i;             /* first occurrence in the chunk: resolved in phase R */
				}
				const uint32_t hb = __ballot_sync(QOI_FULL_MASK, hit);
				__syncwarp();
				if (nonrun && (m & gt) == 0u) { /* last pixel of its slot in this step becomes the table entry */
					Lval[h] = v;
					Lflag[h] = 1;
				}
				if (lane == 0) {
					hitw[warp * QOI_WARP_STEPS + s] = hb;
					nrw[warp * QOI_WARP_STEPS + s] = nr;
					lnrs[warp * QOI_WARP_STEPS + s] = last_nr;
				}
				if (nr) last_nr = (int32_t)(warp * QOI_WARP_PX + s * 32 + 31 - __clz((int)nr));
				__syncwarp();
			}
			if (lane == 0) wlnr[warp] = last_nr;
		}
		__syncthreads();

		/* ---- phase T: tile aggregate, look-back, inclusive record (warp 0; lane handles slots lane, lane+32) ---- */
		if (warp == 0) {
			uint32_t own_val[2];
			bool own_ok[2];
#pragma unroll
			for (int q = 0; q < 2; q++) {
				const uint32_t s = lane + 32u * q;
				own_ok[q] = false;
				own_val[q] = 0;
				for (int w = QOI_ENC_WARPS - 1; w >= 0; w--) {
					if (Lflag_all[w * 64 + s]) {
						own_val[q] = Lval_all[w * 64 + s];
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
			uint32_t in_val[2] = {0, 0};
			bool in_ok[2] = {false, false};
			int32_t lnr_in = -1;
			if (t != ti.first_tile) {
				P.st.agg_tab[(size_t)t * 64 + lane] = own_val[0];
				P.st.agg_tab[(size_t)t * 64 + lane + 32] = own_val[1];
				if (lane == 0) {
					P.st.agg_valid[t] = (uint64_t)own_lo | ((uint64_t)own_hi << 32);
					P.st.agg_lnr[t] = own_lnr;
				}
				__syncwarp();
				if (lane == 0) {
					__threadfence();
					st_release_u32(&P.st.tab_status[t], (epoch << 2) | QOI_ST_AGG);
				}
				/* look back: nearest predecessor that wrote the slot, or the first inclusive record */
#pragma unroll
				for (int q = 0; q < 2; q++) {
					const uint32_t s = lane + 32u * q;
					uint32_t tp = t - 1;
					while (true) {
						const uint32_t st = ld_acquire_u32(&P.st.tab_status[tp]);
						if ((st >> 2) != epoch || (st & 3u) == QOI_ST_NONE) {
							spin_pause();
							continue;
						}
						if ((st & 3u) == QOI_ST_INC) {
							in_ok[q] = (__ldcg(&P.st.inc_valid[tp]) >> s) & 1ull;
							in_val[q] = __ldcg(&P.st.inc_tab[(size_t)tp * 64 + s]);
							break;
						}
						if ((__ldcg(&P.st.agg_valid[tp]) >> s) & 1ull) {
							in_ok[q] = true;
							in_val[q] = __ldcg(&P.st.agg_tab[(size_t)tp * 64 + s]);
							break;
						}
						tp--;
					}
				}
				if (lane == 0) {
					uint32_t tp = t - 1;
					while (true) {
						const uint32_t st = ld_acquire_u32(&P.st.tab_status[tp]);
						if ((st >> 2) != epoch || (st & 3u) == QOI_ST_NONE) {
							spin_pause();
							continue;
						}
						if ((st & 3u) == QOI_ST_INC) {
							lnr_in = __ldcg(&P.st.inc_lnr[tp]);
							break;
						}
						const int32_t a = __ldcg(&P.st.agg_lnr[tp]);
						if (a != QOI_LNR_NONE) {
							lnr_in = a;
							break;
						}
						tp--;
					}
				}
			}
			Tin_val[lane] = in_val[0];
			Tin_val[lane + 32] = in_val[1];
			Tin_flag[lane] = in_ok[0] ? 1 : 0;
			Tin_flag[lane + 32] = in_ok[1] ? 1 : 0;
			/* inclusive record = what qoi.h's index[] / run bookkeeping hold after this tile */
			const uint32_t inc_lo = __ballot_sync(QOI_FULL_MASK, own_ok[0] || in_ok[0]);
			const uint32_t inc_hi = __ballot_sync(QOI_FULL_MASK, own_ok[1] || in_ok[1]);
			P.st.inc_tab[(size_t)t * 64 + lane] = own_ok[0] ? own_val[0] : in_val[0];
			P.st.inc_tab[(size_t)t * 64 + lane + 32] = own_ok[1] ? own_val[1] : in_val[1];
			if (lane == 0) {
				P.st.inc_valid[t] = (uint64_t)inc_lo | ((uint64_t)inc_hi << 32);
				P.st.inc_lnr[t] = (own_lnr != QOI_LNR_NONE) ? own_lnr : lnr_in;
				/* last non-run pixel index before each warp chunk (image coordinates) */
				int32_t run = lnr_in;
				for (int w = 0; w < QOI_ENC_WARPS; w++) {
					wlnr_in[w] = run;
					if (wlnr[w] >= 0) run = (int32_t)ti.px0 + wlnr[w];
				}
			}
			__syncwarp();
			if (lane == 0) {
				__threadfence();
				st_release_u32(&P.st.tab_status[t], (epoch << 2) | QOI_ST_INC);
			}
		}
		__syncthreads();

		/* ---- phase R: resolve the deferred first occurrences (<= 64 per warp chunk) ---- */
#pragma unroll
		for (int j = 0; j < 2; j++) {
			const uint32_t q = tid * 2 + j;
			const int w = (int)(q >> 6);
			const uint32_t s = q & 63u;
			const uint32_t pos = Fpos_all[w * 64 + s];
			if (pos != 0xFFFFu) {
				const uint32_t v = pxv[pos];
				bool found = false;
				uint32_t val = 0;
				for (int w2 = w - 1; w2 >= 0; w2--) {
					if (Lflag_all[w2 * 64 + s]) {
						val = Lval_all[w2 * 64 + s];
						found = true;
						break;
					}
				}
				if (!found && Tin_flag[s]) {
					val = Tin_val[s];
					found = true;
				}
				/* an untouched index[] slot holds 0 (qoi.h:393): only the all-zero pixel can hit it */
				const bool hit = found ? (val == v) : (v == 0u);
				if (hit) atomicOr(&hitw[pos >> 5], 1u << (pos & 31u));
			}
		}
		__syncthreads();

		/* ---- phase C: blocked classify pass, 16 consecutive pixels per thread ---- */
		uint32_t v[16];
		uint32_t code[16];
		uint32_t ttotal = 0;
		const uint32_t i0 = tid * 16u;
		uint32_t pprev;
		{
			const uint4 *src = (const uint4 *)(pxv + i0);
			const uint4 q0 = src[0], q1 = src[1], q2 = src[2], q3 = src[3];
			v[0] = q0.x; v[1] = q0.y; v[2] = q0.z; v[3] = q0.w;
			v[4] = q1.x; v[5] = q1.y; v[6] = q1.z; v[7] = q1.w;
			v[8] = q2.x; v[9] = q2.y; v[10] = q2.z; v[11] = q2.w;
			v[12] = q3.x; v[13] = q3.y; v[14] = q3.z; v[15] = q3.w;
			pprev = pxv[(int)i0 - 1];
			const uint32_t pnext = pxv[i0 + 16];
			const uint32_t wi = i0 >> 5, bo = i0 & 31u;
			const uint32_t hitbits = (hitw[wi] >> bo) & 0xFFFFu;
			const uint32_t nrword = nrw[wi];
			const uint32_t nrbits = (nrword >> bo) & 0xFFFFu;
			/* run position of the pixel before my first one */
			const uint32_t below = nrword & ((1u << bo) - 1u);
			int32_t lnr_glob;
			if (below) lnr_glob = (int32_t)(ti.px0 + (i0 & ~31u) + 31u - (uint32_t)__clz((int)below));
			else {
				const int32_t l = lnrs[wi];
				lnr_glob = (l >= 0) ? (int32_t)ti.px0 + l : wlnr_in[warp];
			}
			const uint32_t c = (uint32_t)((int32_t)(ti.px0 + i0) - 1 - lnr_glob); /* run pixels between them */
			int32_t r = (c == 0u) ? -1 : (int32_t)((c - 1u) % 62u);               /* (k-1) % 62 of the previous pixel */
			int32_t cand = INT_MAX;
			uint32_t before = pprev;
#pragma unroll
			for (int j = 0; j < 16; j++) {
				const uint32_t gi = ti.px0 + i0 + j;
				uint32_t cd = 0;
				if (i0 + j < ti.n_valid) {
					if ((nrbits >> j) & 1u) {
						r = -1;
						if (cand == INT_MAX) cand = (int32_t)(i0 + j);
						if ((hitbits >> j) & 1u) cd = 1u | (qoi_slot(v[j]) << 8); /* qoi.h:432-434 */
						else cd = classify(v[j], before);
					} else {
						/* run pixel (qoi.h:415-421): a byte leaves at run length 62, at the image end, or when the run ends */
						r = (r == 61) ? 0 : r + 1;
						if (r == 0 && cand == INT_MAX) cand = (int32_t)(i0 + j);
						const uint32_t nx = (j == 15) ? pnext : v[j < 15 ? j + 1 : 15];
						if (r == 61 || gi == ti.npx - 1u || nx != v[j]) cd = 1u | ((0xC0u | (uint32_t)r) << 8);
					}
				}
				code[j] = cd;
				ttotal += cd & 7u;
				before = v[j];
			}
			if (tid < 4 && cand != INT_MAX) atomicMin((int *)&misc[MISC_RS_CAND], cand);
		}
		/* block scan of the per-thread byte counts */
		uint32_t incl = ttotal;
#pragma unroll
		for (int dlt = 1; dlt < 32; dlt <<= 1) {
			const uint32_t o = __shfl_up_sync(QOI_FULL_MASK, incl, dlt);
			if (lane >= (uint32_t)dlt) incl += o;
		}
		if (lane == 31) wtot[warp] = incl;
		__syncthreads();

		/* ---- phase B: output offset look-back (warp 0) ---- */
		if (warp == 0) {
			uint32_t total = 0;
#pragma unroll
			for (int w = 0; w < QOI_ENC_WARPS; w++) total += wtot[w];
			uint32_t excl = QOI_HEADER_BYTES;
			if (t != ti.first_tile) {
				if (lane == 0) st_release_u64(&P.st.byte_state[t], pack_bytes(epoch, QOI_ST_AGG, total));
				uint32_t sum = 0;
				int64_t base = (int64_t)t - 1;
				while (true) {
					const int64_t tp = base - lane;
					uint32_t state = QOI_ST_AGG, val = 0;
					if (tp >= (int64_t)ti.first_tile) {
						while (true) {
							const uint64_t st = ld_acquire_u64(&P.st.byte_state[tp]);
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
				st_release_u64(&P.st.byte_state[t], pack_bytes(epoch, QOI_ST_INC, excl + total));
				misc[MISC_TILE_BASE] = excl;
				misc[MISC_TILE_TOTAL] = total;
				tma_store_wait_read(); /* the previous tile's bulk store has finished reading the staging buffer */
			}
		}
		__syncthreads();

		/* ---- phase W: op bytes -> staging (alignment congruent to the destination) ---- */
		const uint32_t tile_base = misc[MISC_TILE_BASE];
		const uint32_t tile_total = misc[MISC_TILE_TOTAL];
		const int32_t rs_cand = (int32_t)misc[MISC_RS_CAND];
		uint8_t *gstream = P.streams + ti.stream_off;
		uint8_t *gdst = gstream + tile_base;
		const uint32_t padb = (uint32_t)((uintptr_t)gdst & 15u);
		{
			uint32_t woff = incl - ttotal; /* exclusive inside the warp */
			for (uint32_t w = 0; w < warp; w++) woff += wtot[w];
			unsigned char *o = stage + padb + woff;
			uint32_t before = pprev;
#pragma unroll
			for (int j = 0; j < 16; j++) {
				const uint32_t cd = code[j];
				const uint32_t sz = cd & 7u;
				if ((int32_t)(i0 + j) == rs_cand) { /* decoder restart point of this tile */
					P.st.rs_px[t] = (int32_t)(ti.px0 + i0 + j);
					P.st.rs_byte[t] = (int32_t)(tile_base + (uint32_t)(o - (stage + padb)));
					P.st.rs_prev[t] = before;
				}
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
				before = v[j];
			}
			if (tid == 0 && rs_cand == INT_MAX) P.st.rs_px[t] = -1;
		}
		__syncthreads();

		/* ---- staging -> global: one bulk store for the 16-byte aligned interior, byte stores for the edges ---- */
		{
			const uint32_t lo_off = (padb + 15u) & ~15u;
			const uint32_t end = padb + tile_total;
			const uint32_t hi_off = end & ~15u;
			const bool bulk = hi_off > lo_off;
			if (tid == 0 && bulk) {
				fence_proxy_async_smem();
				tma_store_1d(gdst + (lo_off - padb), stage + lo_off, hi_off - lo_off);
				tma_store_commit();
			}
			const uint32_t a_hi = bulk ? lo_off : end;          /* [padb, a_hi) */
			const uint32_t b_lo = bulk ? hi_off : end;          /* [b_lo, end)  */
			if (tid >= 32 && tid < 64) {
				for (uint32_t o = padb + (tid - 32); o < a_hi; o += 32) gdst[o - padb] = stage[o];
			} else if (tid >= 64 && tid < 96) {
				for (uint32_t o = b_lo + (tid - 64); o < end; o += 32) gdst[o - padb] = stage[o];
			} else if (tid >= 96 && tid < 96 + QOI_HEADER_BYTES) {
				if (t == ti.first_tile) { /* header, qoi.h:384-388 */
					const uint32_t k = tid - 96;
					uint8_t b;
					if (k < 4) b = (uint8_t)("qoif"[k]);
					else if (k < 8) b = (uint8_t)(ti.width >> (8 * (7 - k)));
					else if (k < 12) b = (uint8_t)(ti.height >> (8 * (11 - k)));
					else if (k == 12) b = (uint8_t)(ti.chan_cs & 0xFFu);
					else b = (uint8_t)(ti.chan_cs >> 8);
					gstream[k] = b;
				}
			} else if (tid >= 128 && tid < 128 + QOI_END_BYTES + 1) {
				if (ti.px0 + ti.n_valid == ti.npx) { /* end marker and length, qoi.h:480-484 */
					const uint32_t k = tid - 128;
					const uint32_t fin = tile_base + tile_total;
					if (k < QOI_END_BYTES) gstream[fin + k] = (k == QOI_END_BYTES - 1) ? 1 : 0;
					else P.lens[ti.img] = (int32_t)(fin + QOI_END_BYTES);
				}
			}
		}
		__syncthreads();
		cur ^= 1;
	}

	/* the last CTA to finish rearms the ticket for the next launch */
	__syncthreads();
	if (tid == 0) {
		tma_store_wait_all();
		const uint32_t done = atomicAdd(&P.st.ticket[1], 1u);
		if (done == gridDim.x - 1u) {
			P.st.ticket[0] = 0;
			P.st.ticket[1] = 0;
		}
	}
}

}  // namespace qoienc
