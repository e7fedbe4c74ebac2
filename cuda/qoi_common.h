/*
 * qoi_common.h -- structures shared by the host C ABI (qoi_gpu.cu) and the kernels.
 */
#pragma once
#include <stdint.h>

#define QOI_TILE_PX 4096        /* pixels per encoder tile (== QOI_GPU_TILE_PX) */
#define QOI_ENC_THREADS 256
#define QOI_ENC_WARPS 8
#define QOI_PX_PER_THREAD 16    /* QOI_TILE_PX / QOI_ENC_THREADS */
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

/* look-back status word: (epoch << 2) | state */
#define QOI_ST_NONE 0u
#define QOI_ST_AGG 1u  /* the tile's own aggregate is published */
#define QOI_ST_INC 2u  /* the inclusive prefix (everything up to and including this tile) is published */

/*
 * Per-tile chained-scan state; after an encode it doubles as the decoder's restart index:
 * the state a decoder needs at the start of tile t is inc_* of tile t-1 plus rs_* of tile t.
 */
struct QoiTileState {
	uint32_t *ticket;      /* [2]: next tile to hand out, number of CTAs that finished */
	uint32_t *tab_status;  /* [tiles] table/run look-back status */
	uint32_t *agg_tab;     /* [tiles][64] last non-run pixel value per hash slot inside the tile */
	uint64_t *agg_valid;   /* [tiles] bit s: slot s written inside the tile */
	int32_t *agg_lnr;      /* [tiles] image pixel index of the tile's last non-run pixel, or QOI_LNR_NONE */
	uint32_t *inc_tab;     /* [tiles][64] table after the tile (what qoi.h's index[] holds there) */
	uint64_t *inc_valid;   /* [tiles] */
	int32_t *inc_lnr;      /* [tiles] last non-run pixel index up to the tile's end, -1 if none yet */
	uint64_t *byte_state;  /* [tiles] (epoch<<34)|(state<<32)|bytes : output-offset look-back */
	int32_t *rs_px;        /* [tiles] first pixel of the tile at which an op starts (-1: none) */
	int32_t *rs_byte;      /* [tiles] stream offset of that op */
	uint32_t *rs_prev;     /* [tiles] value of the pixel before it */
};

struct QoiEncParams {
	const uint8_t *pixels;
	uint8_t *streams;
	int32_t *lens;
	const QoiImgMeta *imgs;
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
	uint32_t n_images;
	uint32_t n_jobs;  /* tiles if use_index, else images */
	int32_t out_channels;
	int32_t use_index;
	QoiTileState st;
};
