/*
 * qoi_common.h -- structures shared by the host C ABI (qoi_gpu.cu) and the kernels.
 */
#pragma once
#include <stdint.h>

#define QOI_TILE_PX 2048        /* pixels per encoder tile == pixels per decoder segment (== QOI_GPU_TILE_PX) */
#define QOI_ENC_WARPS 4         /* compute warps per CTA */
#define QOI_ENC_COMPUTE (QOI_ENC_WARPS * 32)
#define QOI_ENC_THREADS (QOI_ENC_COMPUTE + 32) /* + one service warp: tickets, bulk loads, look-backs */
#define QOI_WARP_PX (QOI_TILE_PX / QOI_ENC_WARPS) /* pixels per warp chunk */
#define QOI_WARP_ROWS (QOI_WARP_PX / 32)          /* 32-pixel rows per warp chunk (lane = pixel) */
#define QOI_HEADER_BYTES 14
#define QOI_END_BYTES 8
#define QOI_LNR_NONE INT32_MIN  /* "no non-run pixel in this tile" in an aggregate record */

/* one image of a plan (device copy) */
struct QoiImgMeta {
	uint64_t pix_off;     /* raw pixels inside the pixel buffer */
	uint64_t stream_off;  /* stream slot inside the stream buffer */
	uint64_t dec_off;     /* decoded pixels inside the decode output buffer */
	uint32_t npx;         /* width * height */
	uint32_t first_tile;  /* global index of this image's tile 0 */
	uint32_t width, height;
	uint8_t channels, colorspace, pad[6];
};

/*
 * Look-back records are self-validating 64-bit words: every word carries the epoch of the launch that
 * wrote it, so a reader needs no fence and no ordering between words -- it just re-reads a word until
 * the epoch matches (relaxed gpu-scope loads/stores, L2 coherent).
 *
 *   table entry   [63:32] pixel value     [31:1] epoch  [0] 1
 *   summary word  [63:32] payload         [31:0] epoch
 *   byte state    [63:34] epoch  [33:32] state  [31:0] bytes
 *
 * One table record per tile, ent[t][64].  A tile publishes the slots it wrote itself as soon as its own pixels are
 * hashed (they are final: nothing before the tile can change them), and the slots it did not write once its look-back
 * has found their value (0 for a slot no pixel of the image has touched yet: qoi.h:393 zero-fills index[]).  After the
 * launch ent[t] is exactly what qoi.h's index[] holds behind the last pixel of tile t.
 */
#define QOI_ST_NONE 0u
#define QOI_ST_AGG 1u  /* the tile's own aggregate is published */
#define QOI_ST_INC 2u  /* the inclusive prefix (everything up to and including this tile) is published */

#define QOI_SUM_OWN_LO 0   /* slots 0..31 written by the tile itself */
#define QOI_SUM_OWN_HI 1   /* slots 32..63 */
#define QOI_SUM_OWN_LNR 2  /* image index of the tile's last non-run pixel (QOI_LNR_NONE: it has none) */
#define QOI_SUM_INC_LNR 3  /* published with the inclusive record: last non-run pixel up to the tile's end (-1: none yet) */
#define QOI_SUM_WORDS 4    /* stride per tile (32 bytes) */

/*
 * Per-tile chained-scan state; after an encode it doubles as the decoder's restart index:
 * the state a decoder needs at the start of tile t is ent of tile t-1 plus rs_* of tile t.
 */
struct QoiTileState {
	uint32_t *ticket;      /* [2]: next tile to hand out, number of CTAs that finished */
	uint64_t *ent;         /* [tiles][64] index[] behind the tile's last pixel (see above) */
	uint64_t *sum;         /* [tiles][4]  summary words of the look-back */
	uint64_t *byte_state;  /* [tiles] output-offset look-back */
	int32_t *rs_px;        /* [tiles] first pixel of the tile at which an op starts (-1: none) */
	int32_t *rs_byte;      /* [tiles] stream offset of that op */
	uint32_t *rs_prev;     /* [tiles] value of the pixel before it */
};

struct QoiEncParams {
	const uint8_t *pixels;
	uint8_t *streams;
	int32_t *lens;
	const QoiImgMeta *imgs;
	const uint32_t *tile_img; /* [n_tiles] image that owns each tile */
	const uint64_t *tile_order; /* [n_tiles] ticket -> tile | image << 32: images interleaved so that in-flight tiles spread over the batch */
	uint32_t n_images;
	uint32_t n_tiles;
	uint32_t epoch; /* 1 .. 2^30-1, different for every launch that shares the state arrays */
	QoiTileState st;
};

struct QoiDecParams {
	const uint8_t *streams;
	const int32_t *lens;
	uint8_t *out;
	int32_t *status; /* may be NULL */
	const QoiImgMeta *imgs;
	const uint32_t *tile_img; /* [tiles] image that owns each tile (indexed decode) */
	uint32_t n_images;
	uint32_t n_jobs;  /* tiles if use_index, else images */
	int32_t out_channels;
	int32_t use_index;
	QoiTileState st;
};
