/*
 * qoi_enc.cuh -- single-pass chained-scan QOI encoder for sm_100a (round-2 rewrite).
 *
 * Replaces the sequential loop of the reference encoder (qoi.h:406-478) for a whole batch of images in one launch.
 * The reference carries px_prev, a run counter, index[64] and the output cursor p from pixel to pixel; in the
 * order-free form each of them is a scan over 2048-pixel tiles:
 *
 *   run position      k(i) = i - (last non-run pixel before i)                      max-scan
 *   INDEX hit         the previous NON-RUN pixel with the same hash slot has the
 *                     same value (no such pixel: only the all-zero pixel hits)     64-wide last-writer scan
 *   output offset     exclusive prefix sum of the bytes each pixel emits            +-scan
 *
 * Persistent CTAs take tiles from an atomic ticket (every predecessor of a running tile is itself running or
 * finished, so look-back cannot deadlock) and publish / look back through per-tile records in HBM (decoupled
 * look-back, fence-free: every 64-bit record word carries its launch epoch).  Pixels are read ONCE: a 1-D TMA bulk
 * copy (cp.async.bulk, SASS UBLKCP) brings a tile into shared memory.
 *
 * What round 1 taught (profiles/r1_final, profiles/r2_ab): nine variants of a 4096-pixel, 10-warp CTA all took 7-8 ms
 * for 32 x 4096^2 whatever their instruction count -- three or four CTAs per SM, each spending most of a tile waiting
 * at hand-off barriers, leave the schedulers without ready warps.  This kernel is built the other way round:
 *
 *   - small CTAs (4 compute warps + 1 service warp, ~22 KB of shared memory, <= 40 registers): ~10 CTAs = 10 tiles
 *     in flight per SM, so one CTA's look-back wait is another CTA's compute time;
 *   - lane = pixel.  A warp owns 512 consecutive pixels and walks them as 16 rows of 32; every step of qoi.h's state
 *     machine is a warp-wide operation (ballot = non-run mask; lanes with my hash slot = match.any or six ballots of the
 *     slot's bits, alternating so that neither the ADU nor the ALU pipe carries them alone; shuffle = the same-slot
 *     predecessor inside the row, one shared-memory word per slot for the rows before);
 *   - two light passes instead of defer-and-patch.  Pass A only finds the last writer of every slot in each warp's
 *     chunk (no look-ups); the tile aggregate is published, the service warp looks back for the incoming table, and
 *     pass B runs with an EXACT table from its first pixel: look-up, op selection, sizes (three ballots + popc, no
 *     shuffle scan) and the op bytes straight into the warp's own slab of shared memory at a running offset;
 *   - the slab leaves with a byte-shifted 16-byte copy as soon as the tile's output offset is known.
 *
 *   compute warps (4)                                     service warp
 *   -----------------                                     ------------
 *   wait tile k (mbarrier: info record + pixels)
 *   E  (EXPAND) raw bytes -> RGBA words
 *   A  per row: non-run mask, slot, same-slot lanes
 *      (six ballots), hit / miss / unknown per pixel,
 *      last writer of each slot -> the warp's table
 *        ------- A_DONE ------->                          T  tile aggregate published; 32-tile window look-back for
 *   C  copy-out of tile k-1: its slab -> stream              the incoming table and the last non-run index;
 *      (shifted 16-byte copy), restart record             <- B_READY(k-1)
 *        <------ T_READY -------                             inclusive record published; ticket for tile k+1
 *   M  my chunk's incoming table = nearest earlier
 *      warp that wrote the slot, else the tile's
 *   B  per row: INDEX / RUN / DIFF / LUMA / RGB / RGBA
 *      exactly like qoi.h:415-474, bytes -> slab
 *        ------- B_DONE ------->                          bulk copy of tile k+1's pixels; output-offset look-back of
 *                                                          tile k (32 tiles per probe) -> B_READY(k); header / end marker
 *
 * The copy-out of a tile waits for its output offset; it runs one tile late, behind the NEXT tile's pass A, so that the
 * offset look-back (and most of the table look-back) is hidden behind work instead of a barrier (profiles/r2_enc_v2: a third
 * of all warp samples sat at B_READY, a fifth at T_READY).
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
	uint32_t n_valid;   /* pixels of this tile that exist */
	uint32_t px0;       /* image pixel index of the tile's first pixel */
	uint32_t npx;       /* pixels in the image */
	uint32_t first_tile;
	uint32_t sh;        /* (global source address) & 15, 0 on the zero-copy path */
	uint32_t chan_cs;   /* channels | colorspace << 8 */
	uint64_t stream_off;
	uint32_t width, height;
	uint32_t nbr_prev, nbr_next; /* the pixels either side of the tile (EXPAND; the zero-copy path puts them next to the pixels) */
	uint32_t pad[2];
};
static_assert(sizeof(TileInfo) == 64, "TileInfo layout");

enum {
	BAR_COMPUTE = 1, /* the compute threads among themselves */
	BAR_A_DONE = 2,  /* compute -> service: per-warp tables of this tile are complete */
	BAR_T_READY = 3, /* service -> compute: incoming table and run carry are in shared memory */
	BAR_B_DONE = 4,  /* compute -> service: op bytes are in the slabs, the pixel buffer is free */
	BAR_B_READY = 5  /* service -> compute: the tile's output offset is known */
};
#define QOI_N_CS QOI_ENC_THREADS /* compute warps + service warp */

template <int CH, bool EXPAND>
struct Layout {
	/* RGBA words pix[-1 .. QOI_TILE_PX]; on the zero-copy path this is the bulk copy's destination */
	static constexpr int OFF_PIX = 0;
	static constexpr int PIX_BYTES = 16 + QOI_TILE_PX * 4 + 16;
	/* EXPAND: raw[sh + k] = source byte k (sh = source address & 15), one spare word behind the end */
	static constexpr int OFF_RAW = OFF_PIX + PIX_BYTES;
	static constexpr int RAW_BYTES = EXPAND ? ((16 + QOI_TILE_PX * CH + 16 + 15) & ~15) : 0;
	/* one slab of op bytes per compute warp: worst case of its chunk + room for the shifted copy's last word */
	static constexpr int OFF_SLAB = OFF_RAW + RAW_BYTES;
	static constexpr int SLAB_BYTES = QOI_WARP_PX * (CH + 1) + 32;
	static constexpr int OFF_WTAB = OFF_SLAB + QOI_ENC_WARPS * SLAB_BYTES; /* uint2 [warps][64]: last non-run pixel per slot in the warp's chunk {value, written} */
	static constexpr int OFF_TIN = OFF_WTAB + QOI_ENC_WARPS * 64 * 8;      /* u32 [warps][64]: the warp's running index[] in pass B */
	static constexpr int OFF_TINS = OFF_TIN + QOI_ENC_WARPS * 64 * 4;      /* u32 [64]: index[] before the tile (service warp) */
	static constexpr int OFF_ROWS = OFF_TINS + 64 * 4;                     /* uint4 [warps][rows]: what pass A found out about each row: non-run, hit, unknown lanes */
	static constexpr int OFF_MBAR = OFF_ROWS + QOI_ENC_WARPS * QOI_WARP_ROWS * 16; /* u64 */
	static constexpr int OFF_INFO = OFF_MBAR + 16;                         /* TileInfo [2] */
	static constexpr int OFF_WTOT = OFF_INFO + 2 * (int)sizeof(TileInfo);  /* u32 [2][warps] bytes per slab, by tile parity (a tile's copy-out runs during the next tile) */
	static constexpr int OFF_WLNR = OFF_WTOT + 2 * QOI_ENC_WARPS * 4;      /* i32 [warps] last non-run pixel of the chunk (tile coordinates, -1: none) */
	static constexpr int OFF_MISC = OFF_WLNR + QOI_ENC_WARPS * 4;          /* u32 [4] */
	static constexpr int TOTAL = OFF_MISC + 16;
	static_assert(SLAB_BYTES % 16 == 0 && OFF_SLAB % 16 == 0 && OFF_WTAB % 8 == 0 && OFF_ROWS % 16 == 0 && OFF_MBAR % 8 == 0 && OFF_INFO % 16 == 0, "alignment");
};

/* misc[] slots */
enum { MISC_TILE_BASE = 0 /* [2], by tile parity */, MISC_LNR_IN = 2 };

__device__ __forceinline__ uint64_t pack_bytes(uint32_t epoch, uint32_t state, uint32_t v) {
	return ((uint64_t)epoch << 34) | ((uint64_t)state << 32) | v;
}
__device__ __forceinline__ uint64_t pack_ent(uint32_t epoch, uint32_t value) { return ((uint64_t)value << 32) | (uint64_t)((epoch << 1) | 1u); }
__device__ __forceinline__ uint64_t pack_sum(uint32_t epoch, uint32_t payload) { return ((uint64_t)payload << 32) | epoch; }

/* ---------------------------------------------------------------- tile fetch (service warp) */
struct Fetch {
	uint32_t tile, img;
	bool any;
	QoiImgMeta m; /* read with the ticket, a pass ahead of the bulk copy */
};

/* step 1: take a ticket -- only when the tile can start as soon as its pixels are there.  (Round 1 and the first version of
   this kernel took it a whole tile ahead: the predecessor of a running tile could then still be queued behind another tile
   of its CTA, and every look-back waited for tiles that had not even started -- ~20 us per tile whatever the content.) */
__device__ __forceinline__ uint32_t ticket_take(const QoiEncParams &P) {
	uint32_t tk = 0;
	if (lane_id() == 0) tk = atomicAdd(&P.st.ticket[0], 1u);
	return tk; /* valid in lane 0 */
}
template <int CH>
__device__ __forceinline__ void ticket_resolve(const QoiEncParams &P, uint32_t tk_lane0, Fetch &F) {
	const uint32_t tk = __shfl_sync(QOI_FULL_MASK, tk_lane0, 0);
	F.any = tk < P.n_tiles;
	F.tile = 0;
	F.img = 0;
	if (F.any) {
		const uint64_t rec = P.tile_order[tk]; /* ascending inside every image, images interleaved */
		F.tile = (uint32_t)rec;
		F.img = (uint32_t)(rec >> 32);
		F.m = P.imgs[F.img];
		/* the tile's pixels (and the pixel before / after them) on their way into L2: the bulk copy that follows a pass
		   later then takes an L2 round trip, not a DRAM one, while the compute warps wait for it */
		const uint32_t px0 = (F.tile - F.m.first_tile) * QOI_TILE_PX;
		const uint32_t n_valid = (F.m.npx - px0) < (uint32_t)QOI_TILE_PX ? (F.m.npx - px0) : (uint32_t)QOI_TILE_PX;
		const uint8_t *gsrc = P.pixels + F.m.pix_off + (uint64_t)px0 * CH;
		const uint8_t *lo = px0 ? gsrc - CH : gsrc, *hi = gsrc + (size_t)n_valid * CH + (px0 + n_valid < F.m.npx ? CH : 0);
		const uintptr_t l0 = (uintptr_t)lo & ~(uintptr_t)127;
		for (uintptr_t a = l0 + 128u * lane_id(); a < (uintptr_t)hi; a += 128u * 32u) prefetch_l2((const void *)a);
	}
}

/* step 2: publish the info record `ib`, fill fragments / neighbours, start the bulk copy.  The mbarrier's phase
   completes when the copy has landed; its arrive (release) also makes the info record visible to the waiters. */
template <int CH, bool EXPAND>
__device__ __forceinline__ void fetch_issue(const QoiEncParams &P, unsigned char *smem, const Fetch &F, int ib) {
	using L = Layout<CH, EXPAND>;
	const uint32_t lane = lane_id();
	TileInfo *info = (TileInfo *)(smem + L::OFF_INFO) + ib;
	uint64_t *mbar = (uint64_t *)(smem + L::OFF_MBAR);
	if (!F.any) {
		if (lane == 0) {
			info->tile = -1;
			mbar_arrive_expect_tx(mbar, 0u);
		}
		return;
	}
	const QoiImgMeta &m = F.m;
	const uint32_t t_in = F.tile - m.first_tile;
	const uint32_t px0 = t_in * QOI_TILE_PX;
	const uint32_t n_valid = (m.npx - px0) < (uint32_t)QOI_TILE_PX ? (m.npx - px0) : (uint32_t)QOI_TILE_PX;
	const uint8_t *gsrc = P.pixels + m.pix_off + (uint64_t)px0 * CH;
	const uint32_t nbytes = n_valid * CH;
	const uint32_t sh = EXPAND ? (uint32_t)((uintptr_t)gsrc & 15u) : 0u;
	unsigned char *data = EXPAND ? smem + L::OFF_RAW : smem + L::OFF_PIX + 16; /* data[sh + k] = gsrc[k] */
	const uint32_t lo_off = (sh + 15u) & ~15u;
	const uint32_t hi_off = (sh + nbytes) & ~15u;
	const uint32_t tma_bytes = hi_off > lo_off ? hi_off - lo_off : 0u;
	/* the bulk copy first: everything below runs while it is in flight.  The phase cannot complete before lane 0's arrive. */
	if (lane == 0 && tma_bytes) {
		mbar_expect_tx(mbar, tma_bytes);
		tma_load_1d_hint(data + lo_off, gsrc + (lo_off - sh), tma_bytes, mbar, l2_evict_first_policy());
	}
	/* the <16-byte fragments on either side of the 16-byte aligned interior */
	const uint32_t a_lo = sh, a_hi = tma_bytes ? lo_off : sh + nbytes;
	const uint32_t b_lo = tma_bytes ? hi_off : 0u, b_hi = tma_bytes ? sh + nbytes : 0u;
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
		if (EXPAND) (lane == 0 ? info->nbr_prev : info->nbr_next) = nb;
		else ((uint32_t *)data)[lane == 0 ? -1 : QOI_TILE_PX] = nb;
	}
	if (lane == 0) {
		info->tile = (int32_t)F.tile;
		info->img = F.img;
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
	__syncwarp(); /* lanes 0/1 and the fragment lanes wrote shared memory that lane 0's arrive releases */
	if (lane == 0) mbar_arrive(mbar);
}

/* the lanes of the row whose 6-bit key equals mine: six ballots.  (match.any does this in one instruction, but that
   instruction holds the ADU pipe for ~50 cycles per warp on sm_100a -- measured, profiles/r2_enc_v1 -- and two of them per
   row made the whole kernel ADU-bound; pass A now uses it on every other row, see there.) */
#define QOI_ENC_MATCH_ROWS 2 /* rows of every four whose slot match is a match.any */
__device__ __forceinline__ uint32_t same_key_lanes(uint32_t key) {
#ifdef QOI_EMU
	uint32_t differ = 0;
	for (int b = 0; b < 6; b++) {
		const bool bit = (key & (1u << b)) != 0u;
		const uint32_t plane = __ballot_sync(QOI_FULL_MASK, bit);
		differ |= plane ^ (bit ? 0xFFFFFFFFu : 0u);
	}
	return ~differ;
#else
	/* four instructions per bit (mask test with a predicate result, vote, select, one LOP3 for differ |= plane ^ mine); left to
	   itself the compiler shifts the bit down, masks it, compares it and negates it -- six */
	uint32_t differ = 0;
#define QOI_KEY_BIT(MASK)                                                                    \
	asm volatile(                                                                            \
	    "{\n\t.reg .pred p;\n\t.reg .b32 t, pl, mn;\n\t"                                     \
	    "and.b32 t, %1, " #MASK ";\n\t"                                                       \
	    "setp.ne.u32 p, t, 0;\n\t"                                                            \
	    "vote.sync.ballot.b32 pl, p, 0xffffffff;\n\t"                                         \
	    "selp.b32 mn, 0xffffffff, 0, p;\n\t"                                                  \
	    "lop3.b32 %0, %0, pl, mn, 0xF6;\n\t}"                                                 \
	    : "+r"(differ)                                                                       \
	    : "r"(key))
	QOI_KEY_BIT(1);
	QOI_KEY_BIT(2);
	QOI_KEY_BIT(4);
	QOI_KEY_BIT(8);
	QOI_KEY_BIT(16);
	QOI_KEY_BIT(32);
#undef QOI_KEY_BIT
	return ~differ;
#endif
}

/* ---------------------------------------------------------------- op of one non-run pixel that is not an INDEX hit */
/* qoi.h:438-474, branch free, on packed r | g<<8 | b<<16 | a<<24 words.  The op class comes back as predicates (the byte
   offsets are ballots of them): DIFF (1 byte), LUMA (2), neither = RGB (4) or, if alpha changed, RGBA (5).
   b0 = first byte, b1 = second byte (LUMA's second byte, or r for RGB / RGBA whose g, b, a follow from v). */
struct OpClass {
	bool diff, luma, a_eq;
	uint32_t diff_b, luma_b0, luma_b1;
};
__device__ __forceinline__ OpClass classify(uint32_t v, uint32_t p) {
	OpClass c;
	/* per-channel difference mod 256 (qoi.h:439-441), no borrows between bytes */
	const uint32_t d = ((v | 0x80808080u) - (p & 0x7F7F7F7Fu)) ^ ((v ^ ~p) & 0x80808080u);
	c.a_eq = d < 0x01000000u; /* alpha unchanged (qoi.h:438) */
	/* DIFF (qoi.h:446-451): dr+2, dg+2, db+2 all in 0..3 */
	const uint32_t t = ((d & 0x7F7F7F7Fu) + 0x00020202u) ^ (d & 0x80808080u);
	c.diff = (t & 0xFFFCFCFCu) == 0u;
	/* three 2-bit fields at bits 0, 8, 16 -> one 6-bit field: the multiplier drops them next to each other at bits 16..21 */
	c.diff_b = 0x40u | ((((t & 0x00030303u) * 0x00100401u) >> 16) & 0x3Fu);
	/* LUMA (qoi.h:443-444, 453-459): vg in -32..31, vr-vg and vb-vg in -8..7.  Signed dot products of the difference
	   bytes give vg + 32, vr - vg + 8, vb - vg + 8 in one instruction each (FMA pipe; the ALU pipe is this kernel's limit).
	   qoi.h wraps vr - vg to a signed char; with vg in -32..31 the wrapped value is in -8..7 exactly when the unwrapped
	   one is, and the other cases are not LUMA anyway. */
	const uint32_t g32 = (uint32_t)__dp4a((int)d, 0x00000100, 32);
	const uint32_t tr = (uint32_t)__dp4a((int)d, 0x0000FF01, 8), tb = (uint32_t)__dp4a((int)d, 0x0001FF00, 8);
	c.luma = !c.diff && c.a_eq && g32 < 64u && (tr | tb) < 16u;
	c.luma_b0 = 0x80u | g32;
	c.luma_b1 = tr * 16u + tb; /* (dr-dg+8) << 4 | (db-dg+8) */
	return c;
}

/* ---------------------------------------------------------------- service warp: table / run look-back over a 32-tile window */
struct LookBack {
	uint32_t val[2]; /* index[] before the tile, slots lane and lane + 32 */
	int32_t lnr_in;  /* last non-run pixel before the tile (image coordinates), -1: none yet */
};

__device__ __forceinline__ void table_lookback(const QoiEncParams &P, uint32_t t, uint32_t first_tile, uint32_t epoch, LookBack &R) {
	const uint32_t lane = lane_id();
	bool need0 = true, need1 = true; /* my two slots */
	bool need_lnr = true;            /* uniform */
	int64_t found0 = -1, found1 = -1; /* tile whose record holds the slot's value; -1: never written (value 0) */
	R.lnr_in = -1;
	int64_t base = (int64_t)t - 1;
	while (true) {
		const int64_t tp = base - lane;
		uint32_t state = QOI_ST_NONE, vlo = 0, vhi = 0;
		int32_t lnr = -1;
		if (tp < (int64_t)first_tile) {
			state = QOI_ST_INC; /* before the image: nothing written, no non-run pixel yet */
		} else { /* one look, no waiting: an unpublished tile only matters if nothing nearer resolves the slot */
			const uint64_t *sw = P.st.sum + (size_t)tp * QOI_SUM_WORDS;
			const uint64_t w3 = ld_relaxed_u64(sw + QOI_SUM_INC_LNR);
			if ((uint32_t)w3 == epoch) {
				state = QOI_ST_INC;
				lnr = (int32_t)(uint32_t)(w3 >> 32);
			} else {
				const uint64_t w0 = ld_relaxed_u64(sw + QOI_SUM_OWN_LO), w1 = ld_relaxed_u64(sw + QOI_SUM_OWN_HI), w2 = ld_relaxed_u64(sw + QOI_SUM_OWN_LNR);
				if ((uint32_t)w0 == epoch && (uint32_t)w1 == epoch && (uint32_t)w2 == epoch) {
					state = QOI_ST_AGG;
					vlo = (uint32_t)(w0 >> 32);
					vhi = (uint32_t)(w1 >> 32);
					lnr = (int32_t)(uint32_t)(w2 >> 32);
				}
			}
		}
		/* lanes [0, usable) hold published tiles; they answer in order, up to the nearest inclusive record */
		const uint32_t known = __ballot_sync(QOI_FULL_MASK, state != QOI_ST_NONE);
		const uint32_t usable = (known == 0xFFFFFFFFu) ? 32u : (uint32_t)__ffs((int)~known) - 1u;
		bool done = false;
		for (uint32_t src = 0; src < usable; src++) { /* uniform */
			const uint32_t st = __shfl_sync(QOI_FULL_MASK, state, (int)src);
			const uint32_t mlo = __shfl_sync(QOI_FULL_MASK, vlo, (int)src);
			const uint32_t mhi = __shfl_sync(QOI_FULL_MASK, vhi, (int)src);
			const int32_t ll = __shfl_sync(QOI_FULL_MASK, lnr, (int)src);
			const int64_t tile = base - (int64_t)src;
			const bool inc = st == QOI_ST_INC;
			const int64_t holder = tile >= (int64_t)first_tile ? tile : -1;
			if (need0 && (inc || ((mlo >> lane) & 1u))) { need0 = false; found0 = holder; }
			if (need1 && (inc || ((mhi >> lane) & 1u))) { need1 = false; found1 = holder; }
			if (need_lnr && (inc || ll != QOI_LNR_NONE)) { need_lnr = false; R.lnr_in = ll; }
			if (inc || (!need_lnr && !__ballot_sync(QOI_FULL_MASK, need0 || need1))) { /* an inclusive record answers everything */
				done = true;
				break;
			}
		}
		if (done) break;
		if (usable == 0u) spin_pause(); /* the nearest undecided tile has not published yet */
		base -= usable;
	}
	/* the values themselves (a slot a tile did not write itself appears with its inclusive record) */
	R.val[0] = R.val[1] = 0u;
#pragma unroll
	for (int q = 0; q < 2; q++) {
		const int64_t f = q ? found1 : found0;
		if (f >= 0) {
			const uint64_t *src = P.st.ent + (size_t)f * 64 + lane + 32u * q;
			uint64_t e;
			while (true) {
				e = ld_relaxed_u64(src);
				if (((uint32_t)e >> 1) == epoch) break;
				spin_pause();
			}
			R.val[q] = (uint32_t)(e >> 32);
		}
	}
}

/* compile-time flags for the row bodies of pass B */
struct Yes { static constexpr bool value = true; };
struct No { static constexpr bool value = false; };

/* ---------------------------------------------------------------- the kernel */
#ifndef QOI_ENC_MIN_BLOCKS
#define QOI_ENC_MIN_BLOCKS 9 /* register budget: 40 (45 warps fit in 64 K registers) */
#endif

template <int CH, bool EXPAND>
__global__ void __launch_bounds__(QOI_ENC_THREADS, QOI_ENC_MIN_BLOCKS) qoi_encode_kernel(const QoiEncParams P) {
	using L = Layout<CH, EXPAND>;
	QOI_DYNAMIC_SMEM(smem);
	const uint32_t tid = threadIdx.x, lane = tid & 31u;
	/* through a shuffle so that the compiler knows the warp index (and every branch on it) is uniform: without that it guards
	   each warp-wide instruction of the row loops with a divergence check */
	const uint32_t warp = __shfl_sync(QOI_FULL_MASK, tid >> 5, 0);
	const bool service = warp == QOI_ENC_WARPS;
	uint64_t *mbar = (uint64_t *)(smem + L::OFF_MBAR);
	TileInfo *infos = (TileInfo *)(smem + L::OFF_INFO);
	uint32_t *pix = (uint32_t *)(smem + L::OFF_PIX + 16); /* pix[-1 .. QOI_TILE_PX] */
	uint2 *wtab_all = (uint2 *)(smem + L::OFF_WTAB);
	uint32_t *Tins = (uint32_t *)(smem + L::OFF_TINS);
	uint32_t *wtot = (uint32_t *)(smem + L::OFF_WTOT);
	int32_t *wlnr = (int32_t *)(smem + L::OFF_WLNR);
	uint32_t *misc = (uint32_t *)(smem + L::OFF_MISC);
	const uint32_t epoch = P.epoch;

	if (tid == 0) {
		mbar_init(mbar, 1);
		mbar_fence_init();
	}
	__syncthreads();

	if (service) {
		/* =============================== service warp =============================== */
		Fetch F;
		ticket_resolve<CH>(P, ticket_take(P), F);
		fetch_issue<CH, EXPAND>(P, smem, F, 0);
		int cur = 0;
		uint32_t par = 0; /* tile parity of this CTA: which copy of wtot[] / misc[] belongs to the tile */
		while (F.any) {
			const TileInfo *ti = infos + cur;
			const uint32_t t = F.tile;
			const uint32_t first_tile = ti->first_tile, px0 = ti->px0;

			bar_sync(BAR_A_DONE, QOI_N_CS); /* the per-warp tables and wlnr[] of this tile are complete */
			/* ---- phase T: tile aggregate = the last warp that wrote each slot (lane handles slots lane, lane + 32) ---- */
			uint32_t own_val[2];
			bool own_ok[2];
#pragma unroll
			for (int q = 0; q < 2; q++) {
				own_ok[q] = false;
				own_val[q] = 0;
#pragma unroll
				for (int w = 0; w < QOI_ENC_WARPS; w++) {
					const uint2 e = wtab_all[w * 64 + lane + 32 * q];
					if (e.y) {
						own_val[q] = e.x;
						own_ok[q] = true;
					}
				}
			}
			int32_t own_lnr = QOI_LNR_NONE;
#pragma unroll
			for (int w = 0; w < QOI_ENC_WARPS; w++)
				if (wlnr[w] >= 0) own_lnr = (int32_t)px0 + wlnr[w];
			LookBack lb;
			lb.val[0] = lb.val[1] = 0u;
			lb.lnr_in = -1;
			if (t != first_tile) {
				const uint32_t own_lo = __ballot_sync(QOI_FULL_MASK, own_ok[0]);
				const uint32_t own_hi = __ballot_sync(QOI_FULL_MASK, own_ok[1]);
				if (own_ok[0]) st_relaxed_u64(&P.st.ent[(size_t)t * 64 + lane], pack_ent(epoch, own_val[0]));
				if (own_ok[1]) st_relaxed_u64(&P.st.ent[(size_t)t * 64 + lane + 32], pack_ent(epoch, own_val[1]));
				if (lane < 3) {
					const uint32_t payload = lane == 0 ? own_lo : (lane == 1 ? own_hi : (uint32_t)own_lnr);
					st_relaxed_u64(&P.st.sum[(size_t)t * QOI_SUM_WORDS + lane], pack_sum(epoch, payload));
				}
				table_lookback(P, t, first_tile, epoch, lb);
			}
			Tins[lane] = lb.val[0];
			Tins[lane + 32] = lb.val[1];
			if (lane == 0) misc[MISC_LNR_IN] = (uint32_t)lb.lnr_in;
			__syncwarp();
			bar_arrive(BAR_T_READY, QOI_N_CS);
			/* inclusive record = what qoi.h's index[] / run bookkeeping hold after this tile */
			if (t == first_tile || !own_ok[0]) st_relaxed_u64(&P.st.ent[(size_t)t * 64 + lane], pack_ent(epoch, own_ok[0] ? own_val[0] : lb.val[0]));
			if (t == first_tile || !own_ok[1]) st_relaxed_u64(&P.st.ent[(size_t)t * 64 + lane + 32], pack_ent(epoch, own_ok[1] ? own_val[1] : lb.val[1]));
			if (lane == 0) {
				const int32_t inc_lnr = (own_lnr != QOI_LNR_NONE) ? own_lnr : lb.lnr_in;
				st_relaxed_u64(&P.st.sum[(size_t)t * QOI_SUM_WORDS + QOI_SUM_INC_LNR], pack_sum(epoch, (uint32_t)inc_lnr));
			}

			/* the next tile's ticket: taken while the compute warps are in pass B, so the tile starts a fixed, short time
			   after its ticket (a ticket taken a whole tile ahead lets a predecessor tile sit unstarted behind another tile of
			   its CTA while its successors already look back for it) */
			Fetch Fn;
			ticket_resolve<CH>(P, ticket_take(P), Fn);
			const uint32_t img_cur = ti->img, npx_cur = ti->npx, nv_cur = ti->n_valid;
			const uint64_t so_cur = ti->stream_off;
			const uint32_t w_cur = ti->width, h_cur = ti->height, cc_cur = ti->chan_cs;

			bar_sync(BAR_B_DONE, QOI_N_CS); /* slab sizes are in wtot[par]; nobody reads the pixel buffer any more */
			fetch_issue<CH, EXPAND>(P, smem, Fn, cur ^ 1); /* its info record, its pixels on their way */
			/* ---- phase B: output offset look-back, 32 predecessors per probe ---- */
			uint32_t total = 0;
#pragma unroll
			for (int w = 0; w < QOI_ENC_WARPS; w++) total += wtot[par * QOI_ENC_WARPS + w];
			uint32_t excl = QOI_HEADER_BYTES;
			if (t != first_tile) {
				if (lane == 0) st_relaxed_u64(&P.st.byte_state[t], pack_bytes(epoch, QOI_ST_AGG, total));
				uint32_t sum = 0;
				int64_t base = (int64_t)t - 1;
				while (true) {
					const int64_t tp = base - lane;
					uint32_t state = QOI_ST_AGG, val = 0;
					if (tp >= (int64_t)first_tile) {
						while (true) {
							const uint64_t st = ld_relaxed_u64(&P.st.byte_state[tp]);
							if ((uint32_t)(st >> 34) == epoch && ((st >> 32) & 3u) != QOI_ST_NONE) {
								state = (uint32_t)(st >> 32) & 3u;
								val = (uint32_t)st;
								break;
							}
							spin_pause();
						}
					} /* tiles before the image count as empty aggregates; the image's first tile always publishes an inclusive record */
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
				misc[MISC_TILE_BASE + par] = excl;
			}
			__syncwarp();
			bar_arrive(BAR_B_READY, QOI_N_CS);
			{
				uint8_t *gstream = P.streams + so_cur;
				if (t == first_tile && lane < QOI_HEADER_BYTES) { /* header, qoi.h:384-388 */
					const uint32_t k = lane;
					uint8_t b;
					if (k < 4) b = (uint8_t)("qoif"[k]);
					else if (k < 8) b = (uint8_t)(w_cur >> (8 * (7 - k)));
					else if (k < 12) b = (uint8_t)(h_cur >> (8 * (11 - k)));
					else if (k == 12) b = (uint8_t)(cc_cur & 0xFFu);
					else b = (uint8_t)(cc_cur >> 8);
					gstream[k] = b;
				}
				if (px0 + nv_cur == npx_cur && lane >= 16 && lane < 16 + QOI_END_BYTES + 1) { /* end marker, length: qoi.h:480-484 */
					const uint32_t k = lane - 16;
					const uint32_t fin = excl + total;
					if (k < QOI_END_BYTES) gstream[fin + k] = (k == QOI_END_BYTES - 1) ? 1 : 0;
					else P.lens[img_cur] = (int32_t)(fin + QOI_END_BYTES);
				}
			}
			par ^= 1u;
			F = Fn;
			cur ^= 1;
		}
		/* the last CTA to finish rearms the ticket for the next launch */
		if (lane == 0) {
			const uint32_t done = atomicAdd(&P.st.ticket[1], 1u);
			if (done == gridDim.x - 1u) {
				P.st.ticket[0] = 0;
				P.st.ticket[1] = 0;
			}
		}
		return;
	}

	/* =============================== compute warps =============================== */
	const uint32_t lt = lanemask_lt(), gt = lanemask_gt();
	const uint32_t wbase = warp * (uint32_t)QOI_WARP_PX;
	uint2 *wtab = wtab_all + warp * 64;
	uint32_t *tin = (uint32_t *)(smem + L::OFF_TIN) + warp * 64;
	unsigned char *slab = smem + L::OFF_SLAB + warp * L::SLAB_BYTES;
	uint4 *rows = (uint4 *)(smem + L::OFF_ROWS) + warp * QOI_WARP_ROWS;
	const uint32_t lanebit = 1u << lane;
	int cur = 0;
	uint32_t parity = 0;
	/* the tile whose op bytes sit in my slab, waiting for its output offset (copied out one tile late, see the header) */
	bool c_valid = false;
	uint32_t c_t = 0, c_bytes = 0, c_px0 = 0, c_par = 0, c_rs_off = 0xFFFFFFFFu, c_rs_prev = 0;
	int32_t c_rs_cand = INT_MAX;
	uint8_t *c_gstream = nullptr;
	/* ---- phase C: my slab -> tile base + bytes of the warps before me (byte-shifted 16-byte copy), restart record ---- */
	auto copy_out = [&]() {
		bar_sync(BAR_B_READY, QOI_N_CS); /* that tile's output offset is in misc[] */
		uint32_t woff = misc[MISC_TILE_BASE + c_par];
#pragma unroll
		for (int w2 = 0; w2 < QOI_ENC_WARPS - 1; w2++)
			if ((uint32_t)w2 < warp) woff += wtot[c_par * QOI_ENC_WARPS + w2];
		if (warp == 0) { /* restart record: the lane that owns the restart pixel knows where its op starts */
			if (c_rs_off != 0xFFFFFFFFu) {
				P.st.rs_px[c_t] = (int32_t)c_px0 + c_rs_cand;
				P.st.rs_byte[c_t] = (int32_t)(woff + c_rs_off);
				P.st.rs_prev[c_t] = c_rs_prev;
			}
			if (lane == 0 && c_rs_cand == INT_MAX) P.st.rs_px[c_t] = -1;
		}
		uint8_t *gdst = c_gstream + woff;
		const uint32_t total = c_bytes;
		uint32_t head = (16u - (uint32_t)((uintptr_t)gdst & 15u)) & 15u; /* bytes up to the first aligned 16 */
		if (head > total) head = total;
		const uint32_t nvec = (total - head) >> 4;
		if (lane < head) gdst[lane] = slab[lane];
		const uint32_t *s32 = (const uint32_t *)slab + (head >> 2);
		const uint32_t shf = (head & 3u) * 8u;
		for (uint32_t k = lane; k < nvec; k += 32u) {
			const uint32_t *w = s32 + 4u * k;
			const uint32_t w0 = w[0], w1 = w[1], w2 = w[2], w3 = w[3], w4 = w[4];
			uint4 q;
			q.x = __funnelshift_r(w0, w1, shf);
			q.y = __funnelshift_r(w1, w2, shf);
			q.z = __funnelshift_r(w2, w3, shf);
			q.w = __funnelshift_r(w3, w4, shf);
			st_stream_u4((uint4 *)(gdst + head + 16u * k), q);
		}
		const uint32_t tail0 = head + 16u * nvec;
		if (tail0 + lane < total) gdst[tail0 + lane] = slab[tail0 + lane]; /* fewer than 16 bytes */
		__syncwarp(); /* the slab is free for the next tile's pass B */
	};
	while (true) {
		mbar_wait(mbar, parity); /* info record and pixels of the tile have landed */
		parity ^= 1u;
		const TileInfo *ti = infos + cur;
		if (__ballot_sync(QOI_FULL_MASK, ti->tile < 0)) break; /* (a vote, so that the exit is visibly uniform) */
		const uint32_t t = (uint32_t)ti->tile;
		const uint32_t nv = ti->n_valid, px0 = ti->px0, npx = ti->npx;

		if (EXPAND) {
			/* ---- phase E: raw bytes at raw[sh + CH * i] -> RGBA words (lane = pixel, two aligned words + funnel shift) ---- */
			const uint32_t *raw32 = (const uint32_t *)(smem + L::OFF_RAW);
			const uint32_t sh = ti->sh;
			for (uint32_t i = tid; i < nv; i += QOI_ENC_COMPUTE) {
				const uint32_t b = sh + i * CH;
				const uint32_t x = __funnelshift_r(raw32[b >> 2], raw32[(b >> 2) + 1], (b & 3u) * 8u);
				pix[i] = CH == 4 ? x : ((x & 0x00FFFFFFu) | 0xFF000000u);
			}
			if (tid == 0) {
				pix[-1] = ti->nbr_prev;
				pix[QOI_TILE_PX] = ti->nbr_next;
			}
			bar_sync(BAR_COMPUTE, QOI_ENC_COMPUTE);
		}

		/* ---- pass A: everything about the INDEX look-up that my chunk can know on its own.  Per row: non-run mask, hash slot,
		   the row's lanes with my slot; a pixel with an earlier same-slot pixel in the chunk (this row: shuffle, rows before:
		   the warp's table) is a known hit or miss, a slot's first pixel in the chunk stays UNKNOWN until the table before
		   the chunk is there.  The table ends up holding the chunk's last non-run pixel per slot = its aggregate. ---- */
		wtab[lane] = make_uint2(0u, 0u);
		wtab[lane + 32] = make_uint2(0u, 0u);
		__syncwarp();
		{
			int32_t last_nr = -1; /* tile-local index of the chunk's last non-run pixel so far (uniform) */
			const uint32_t *px = pix + wbase + lane;
			uint32_t il = wbase + lane;
#pragma unroll 4
			for (int s = 0; s < QOI_WARP_ROWS; s++, px += 32, il += 32u) {
				const uint32_t v = px[0], p = px[-1];
				const bool valid = il < nv;
				const bool nonrun = valid && v != p; /* qoi.h:415 */
				const uint32_t nr = __ballot_sync(QOI_FULL_MASK, nonrun);
				const uint32_t vm = __ballot_sync(QOI_FULL_MASK, valid);
				uint32_t hb = 0, ub = 0;
				if (nr) { /* uniform: a row of run pixels touches nothing */
					const uint32_t h = qoi_slot(v);
#ifdef QOI_EMU
					const uint32_t m = same_key_lanes(h) & nr;
#else
					/* the row's lanes with my slot: match.any on two rows of every four, six ballots on the other two.  Both give
					   the same mask; one runs on the ADU pipe (~50 cycles per warp), the other is 19 ALU-pipe instructions --
					   alone either pipe is the kernel's limit (ALU 75 % with ballots only, ADU 73 % with match.any only:
					   4.42 / 4.37 ms), mixed 1:1 they share the load: 4.12 ms (profiles/r2_enc_v4) */
					const uint32_t m = ((s & 3) < QOI_ENC_MATCH_ROWS ? __match_any_sync(QOI_FULL_MASK, h) : same_key_lanes(h)) & nr;
#endif
					const uint32_t lower = m & lt;
					const uint2 e = wtab[h];                                                 /* as the rows before left it */
					const uint32_t pv = __shfl_sync(QOI_FULL_MASK, v, 31 - __clz((int)lower)); /* nearest same-slot pixel of this row (garbage if none) */
					const bool known = lower != 0u || e.y != 0u;
					hb = __ballot_sync(QOI_FULL_MASK, nonrun && known && (lower ? pv : e.x) == v); /* qoi.h:432 */
					ub = __ballot_sync(QOI_FULL_MASK, nonrun && !known);
					__syncwarp(); /* every lane has read the table (the votes above need each lane's e, but only this orders memory: racecheck) */
					if (nonrun && (m & gt) == 0u) wtab[h] = make_uint2(v, 1u); /* qoi.h:436: the row's last pixel of its slot */
					last_nr = (int32_t)(il - lane + 31u - (uint32_t)__clz((int)nr));
				}
				if (lane == 0) rows[s] = make_uint4(nr, hb, ub, vm);
				__syncwarp(); /* the next row reads the table after these stores */
			}
			if (lane == 0) wlnr[warp] = last_nr;
		}
		__syncwarp();
		bar_arrive(BAR_A_DONE, QOI_N_CS);

		if (c_valid) copy_out(); /* the tile before this one leaves while the service warp looks back for this one */

		bar_sync(BAR_T_READY, QOI_N_CS); /* every warp's table is final; index[] before the tile and the run carry are in shared memory */
		/* ---- phase M: index[] before my chunk = nearest earlier warp that wrote the slot, else the tile's incoming table ---- */
#pragma unroll
		for (int q = 0; q < 2; q++) {
			const uint32_t sl = lane + 32u * q;
			uint32_t val = Tins[sl];
#pragma unroll
			for (int w2 = 0; w2 < QOI_ENC_WARPS - 1; w2++) {
				if ((uint32_t)w2 < warp) {
					const uint2 e = wtab_all[w2 * 64 + sl];
					if (e.y) val = e.x;
				}
			}
			tin[sl] = val;
		}
		int32_t lnr = (int32_t)misc[MISC_LNR_IN]; /* last non-run pixel before the current row (image coordinates; -1: none since the image start) */
#pragma unroll
		for (int w2 = 0; w2 < QOI_ENC_WARPS - 1; w2++)
			if ((uint32_t)w2 < warp && wlnr[w2] >= 0) lnr = (int32_t)px0 + wlnr[w2];
		__syncwarp();

		/* ---- pass B: ops and bytes.  The rows are independent now (tin[] is read only); only the byte offset and the run
		   carry go from row to row.  One straight-line body per kind of row (non-run pixels only / run pixels only / both),
		   picked by a uniform branch, so that the op classes stay predicates from classify() to the stores. ---- */
		uint32_t off = 0;               /* bytes in my slab so far (uniform) */
		int32_t rs_cand = INT_MAX;      /* warp 0: the tile's first pixel at which an op starts = decoder restart point */
		uint32_t rs_off = 0xFFFFFFFFu;  /* its byte offset in the slab, in the lane that owns the pixel */
		uint32_t rs_prev = 0;
		const uint32_t last_px = px0 + nv == npx ? npx - 1u : 0xFFFFFFFFu; /* the image's last pixel, if it lies in this tile */
		{
			const uint32_t *px = pix + wbase + lane;
			uint32_t gi = px0 + wbase + lane;      /* my pixel's index in the image */
			auto row = [&](auto has_nonrun, auto has_run, auto is_head, int s, const uint4 ri) {
				constexpr bool NR = decltype(has_nonrun)::value, RUN = decltype(has_run)::value, HEAD = decltype(is_head)::value;
				const uint32_t nr = ri.x, vmask = ri.w;
				const uint32_t v = px[0], p = px[-1];
				const bool nonrun = (nr & lanebit) != 0u;
				/* op classes of my pixel (at most one is set): INDEX hit, DIFF, LUMA, RGB / RGBA, RUN byte */
				bool c_hit = false, c_diff = false, c_luma = false, c_abs = false, c_a5 = false, c_run = false;
				bool opstart = true;
				uint32_t h = 0, b_diff = 0, b_luma0 = 0, b1 = 0, b_run = 0;
				if (NR) {
					h = qoi_slot(v);
					bool hit = (ri.y & lanebit) != 0u;
					if (ri.z) { /* uniform: a slot's first pixel in my chunk -- qoi.h:432 against index[] before the chunk */
						if ((ri.z & lanebit) && tin[h] == v) hit = true;
					}
					const OpClass c = classify(v, p);
					c_hit = nonrun && hit;
					c_diff = nonrun && !hit && c.diff;
					c_luma = nonrun && !hit && c.luma;
					c_abs = nonrun && !hit && !c.diff && !c.luma;
					c_a5 = c_abs && !c.a_eq;
					b_diff = c.diff_b;
					b_luma0 = c.luma_b0;
					b1 = c.luma ? c.luma_b1 : v; /* second byte: LUMA's, or r of RGB / RGBA (only the low byte is stored) */
				}
				if (RUN) { /* qoi.h:415-421 */
					const uint32_t nx = px[1];
					const uint32_t below = nr & lt;
					const uint32_t lnr_i = below ? (gi - lane + 31u - (uint32_t)__clz((int)below)) : (uint32_t)lnr;
					const uint32_t r = (gi - lnr_i - 1u) % 62u; /* (position in the run - 1) % 62 */
					const bool isrun = !nonrun && (vmask & lanebit) != 0u;
					/* a byte leaves at run length 62, at the image end, or when the run ends */
					c_run = isrun && (r == 61u || nx != v || gi == last_px);
					b_run = 0xC0u | r;
					if (isrun) opstart = r == 0u;
				}
				/* byte offsets inside the row: sizes are 0, 1, 2, 4 or 5 -> three bit planes */
				const uint32_t s1 = __ballot_sync(QOI_FULL_MASK, c_hit || c_diff || c_run || c_a5);
				const uint32_t s2 = NR ? __ballot_sync(QOI_FULL_MASK, c_luma) : 0u;
				const uint32_t s4 = NR ? __ballot_sync(QOI_FULL_MASK, c_abs) : 0u;
				const uint32_t ex = (uint32_t)__popc(s1 & lt) + (NR ? 2u * (uint32_t)__popc(s2 & lt) + 4u * (uint32_t)__popc(s4 & lt) : 0u);
				if (HEAD && rs_cand == INT_MAX) { /* a RUN op covers <= 62 pixels: the restart point lies in the tile's first two rows */
					const uint32_t starts = __ballot_sync(QOI_FULL_MASK, opstart && (vmask & lanebit));
					if (starts) {
						const uint32_t first = (uint32_t)__ffs((int)starts) - 1u;
						rs_cand = (int32_t)((uint32_t)s * 32u + first);
						if (lane == first) {
							rs_off = off + ex;
							rs_prev = p;
						}
					}
				}
				/* the op bytes: one predicated byte store per case (the LSU pipe has room, the ALU pipe that a chain of
				   selects would use has not) */
				unsigned char *o = slab + off + ex;
				if (RUN && c_run) o[0] = (uint8_t)b_run;
				if (NR) {
					if (c_hit) o[0] = (uint8_t)h; /* QOI_OP_INDEX | slot, qoi.h:433 */
					if (c_diff) o[0] = (uint8_t)b_diff;
					if (c_luma) o[0] = (uint8_t)b_luma0;
					if (c_abs) o[0] = c_a5 ? (uint8_t)0xFF : (uint8_t)0xFE; /* qoi.h:461-473 */
					if (c_luma || c_abs) o[1] = (uint8_t)b1;
					if (c_abs) o[2] = (uint8_t)__umulhi(v, 1u << 24); /* v >> 8 on the FMA pipe */
					if (c_abs) o[3] = (uint8_t)__umulhi(v, 1u << 16);
					if (c_a5) o[4] = (uint8_t)__umulhi(v, 1u << 8);
				}
				off += (uint32_t)__popc(s1) + (NR ? 2u * (uint32_t)__popc(s2) + 4u * (uint32_t)__popc(s4) : 0u);
				if (NR) lnr = (int32_t)(gi - lane + 31u - (uint32_t)__clz((int)nr));
			};
			auto pick = [&](auto is_head, int s) {
				const uint4 ri = rows[s];
				if (ri.x == ri.w) {
					if (ri.w) row(Yes(), No(), is_head, s, ri); /* every pixel starts an op of its own */
				} else if (ri.x == 0u) {
					row(No(), Yes(), is_head, s, ri);
				} else {
					row(Yes(), Yes(), is_head, s, ri);
				}
				px += 32;
				gi += 32u;
			};
			int s0 = 0;
			if (warp == 0) { /* uniform: the decoder's restart point of the tile is found in warp 0's first two rows */
				pick(Yes(), 0);
				pick(Yes(), 1);
				s0 = 2;
			}
#pragma unroll 1
			for (int s = s0; s < QOI_WARP_ROWS; s++) pick(No(), s);
		}
		const uint32_t tpar = (parity ^ 1u) & 1u; /* 0 for the CTA's first tile, like the service warp's `par` */
		if (lane == 0) wtot[tpar * QOI_ENC_WARPS + warp] = off;
		c_valid = true;
		c_t = t;
		c_bytes = off;
		c_px0 = px0;
		c_par = tpar;
		c_rs_off = rs_off;
		c_rs_prev = rs_prev;
		c_rs_cand = rs_cand;
		c_gstream = P.streams + ti->stream_off;
		__syncwarp();
		bar_arrive(BAR_B_DONE, QOI_N_CS);
		cur ^= 1;
	}
	if (c_valid) copy_out(); /* the CTA's last tile */
}


}  // namespace qoienc
