/*
 * qoi_common.h -- structures shared by the host C ABI (qoi_gpu.cu) and the kernels.
 */
#pragma once
#include <stdint.h>

#define QOI_TILE_PX 4096        /* pixels per encoder tile (== QOI_GPU_TILE_PX) */
#define QOI_ENC_THREADS 320      /* 8 compute warps + scout warp + loader warp */
#define QOI_ENC_COMPUTE 256
#define QOI_ENC_WARPS 8         /* compute warps */
#define QOI_PX_PER_THREAD 16    /* QOI_TILE_PX / QOI_ENC_COMPUTE */
#define QOI_WARP_PX 512         /* pixels per warp chunk */
#define QOI_WARP_STEPS 16       /* 32-pixel steps per warp chunk */
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
 *   table entry   [63:32] pixel value     [31:1] epoch  [0] slot written
 *   summary word  [63:32] payload         [31:0] epoch           (payload: valid-mask half / last non-run index)
 *   byte state    [63:34] epoch  [33:32] state  [31:0] bytes
 */
#define QOI_ST_NONE 0u
#define QOI_ST_AGG 1u  /* the tile's own aggregate is published */
#define QOI_ST_INC 2u  /* the inclusive prefix (everything up to and including this tile) is published */

#define QOI_SUM_VALID_LO 0
#define QOI_SUM_VALID_HI 1
#define QOI_SUM_LNR 2
#define QOI_SUM_WORDS 4 /* stride per tile (32 bytes) */

/*
 * Per-tile chained-scan state; after an encode it doubles as the decoder's restart index:
 * the state a decoder needs at the start of tile t is inc_ent of tile t-1 plus rs_* of tile t.
 */
struct QoiTileState {
	uint32_t *ticket;      /* [2]: next tile to hand out, number of CTAs that finished */
	uint64_t *agg_ent;     /* [tiles][64] last non-run pixel per hash slot INSIDE the tile */
	uint64_t *agg_sum;     /* [tiles][4]  valid mask halves, last non-run pixel index (QOI_LNR_NONE: none) */
	uint64_t *inc_ent;     /* [tiles][64] table after the tile == what qoi.h's index[] holds there */
	uint64_t *inc_sum;     /* [tiles][4]  same, inclusive (last non-run index: -1 if none yet) */
	uint64_t *byte_state;  /* [tiles] output-offset look-back */
	int32_t *rs_px;        /* [tiles] first pixel of the tile at which an op starts (-1: none) */
	int32_t *rs_byte;      /* [tiles] stream offset of that op */
	uint32_t *rs_prev;     /* [tiles] value of the pixel before it */
	/* a second restart point per tile, at its first op start at or after pixel QOI_TILE_PX / 2 -- written only by the
	   slab-mode encoders of qoi_enc4.cuh (experimental); it lets the indexed decoder run two lanes per tile */
	uint64_t *mid_ent;     /* [tiles][64] table before that op: [63:32] value, bit 0 slot written */
	int32_t *mid_px;       /* [tiles] its pixel (-1: none) */
	int32_t *mid_byte;     /* [tiles] stream offset of the op */
	uint32_t *mid_prev;    /* [tiles] value of the pixel before it */
};

struct QoiEncParams {
	const uint8_t *pixels;
	uint8_t *streams;
	int32_t *lens;
	const QoiImgMeta *imgs;
	const uint32_t *tile_img; /* [n_tiles] image that owns each tile */
	const uint32_t *tile_order; /* [n_tiles] ticket -> tile: images interleaved so that in-flight tiles spread over the batch */
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
