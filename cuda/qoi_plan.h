/*
 * qoi_plan.h -- host-side geometry of a batch: per-image metadata, tile numbering and the layout
 * of the per-tile chained-scan state.  Plain C++ (no CUDA) so that the C ABI (qoi_gpu.cu) and the
 * CPU emulation driver (tests/emu) share one definition.
 */
#pragma once
#include <stddef.h>
#include <stdint.h>

#include <vector>

#include "qoi_common.h"

struct QoiDescPOD { /* same layout as qoi_desc (reference qoi.h:236-241) */
	unsigned int width, height;
	unsigned char channels, colorspace;
};

/* the reference's argument checks for one image: qoi.h:366-369 (encode) == qoi.h:514-518 (decode) */
static inline bool qoi_geometry_ok(const QoiDescPOD &d) {
	if (d.width == 0 || d.height == 0) return false;
	if (d.channels < 3 || d.channels > 4) return false;
	if (d.colorspace > 1) return false;
	if (d.height >= 400000000u / d.width) return false; /* QOI_PIXELS_MAX, qoi.h:332 */
	return true;
}

static inline size_t qoi_max_stream(const QoiDescPOD &d) { /* qoi.h:374-376 */
	return (size_t)d.width * d.height * (d.channels + 1u) + QOI_HEADER_BYTES + QOI_END_BYTES;
}

/* the longest stream a DEcoder can consume: a 3-channel file may still hold 5-byte RGBA ops (qoi.h:552-557) */
static inline size_t qoi_max_foreign_stream(const QoiDescPOD &d) {
	return (size_t)d.width * d.height * 5u + QOI_HEADER_BYTES + QOI_END_BYTES;
}

struct QoiPlanGeom {
	std::vector<QoiImgMeta> metas;
	std::vector<uint32_t> tile_img; /* image index of every tile */
	std::vector<uint64_t> tile_order; /* ticket -> tile | image << 32 (round robin over the images, ascending inside an image) */
	uint32_t n_tiles = 0;
	uint64_t total_px = 0;
	int channels = 0;
	int dec_channels = 0;
	uint64_t pixel_bytes = 0;   /* extent of the pixel buffer (max pix_off + size) */
	uint64_t stream_bytes = 0;  /* extent of the stream buffer */
	uint64_t decoded_bytes = 0; /* extent of the decode output buffer */
};

/* returns 0 or QOI_GPU_E_ARG (1) */
static inline int qoi_plan_geometry(QoiPlanGeom &g, int n, const QoiDescPOD *descs, const uint64_t *pixel_off,
                                    const uint64_t *stream_off, const uint64_t *decoded_off, int decoded_channels) {
	if (n <= 0 || !descs || !pixel_off || !stream_off) return 1;
	if (decoded_channels != 0 && decoded_channels != 3 && decoded_channels != 4) return 1;
	g.metas.resize((size_t)n);
	g.channels = descs[0].channels;
	g.dec_channels = decoded_channels ? decoded_channels : g.channels;
	uint64_t tiles = 0, packed = 0;
	for (int i = 0; i < n; i++) {
		if (!qoi_geometry_ok(descs[i]) || descs[i].channels != g.channels) return 1;
		QoiImgMeta &m = g.metas[(size_t)i];
		m.npx = descs[i].width * descs[i].height;
		m.width = descs[i].width;
		m.height = descs[i].height;
		m.channels = descs[i].channels;
		m.colorspace = descs[i].colorspace;
		m.pix_off = pixel_off[i];
		m.stream_off = stream_off[i];
		m.dec_off = decoded_off ? decoded_off[i] : packed;
		if (g.dec_channels == 4 && (m.dec_off & 3u)) return 1; /* 4-byte pixel stores */
		packed += (uint64_t)m.npx * g.dec_channels;
		m.first_tile = (uint32_t)tiles;
		tiles += (m.npx + QOI_TILE_PX - 1) / QOI_TILE_PX;
		if (tiles > 0x7FFFFFFFull) return 1;
		g.total_px += m.npx;
		uint64_t e;
		e = m.pix_off + (uint64_t)m.npx * m.channels; if (e > g.pixel_bytes) g.pixel_bytes = e;
		e = m.stream_off + qoi_max_foreign_stream(descs[i]); if (e > g.stream_bytes) g.stream_bytes = e;
		e = m.dec_off + (uint64_t)m.npx * g.dec_channels; if (e > g.decoded_bytes) g.decoded_bytes = e;
	}
	g.n_tiles = (uint32_t)tiles;
	g.tile_img.resize(g.n_tiles);
	for (int i = 0; i < n; i++) {
		const QoiImgMeta &m = g.metas[(size_t)i];
		const uint32_t nt = (m.npx + QOI_TILE_PX - 1) / QOI_TILE_PX;
		for (uint32_t k = 0; k < nt; k++) g.tile_img[m.first_tile + k] = (uint32_t)i;
	}
	/* Ticket order.  A tile only ever waits for EARLIER tiles of its own image, so any order that is ascending
	   inside each image is safe; interleaving the images keeps the look-back chains short (the ~450 tiles in
	   flight then belong to many images instead of one). */
	g.tile_order.clear();
	g.tile_order.reserve(g.n_tiles);
	{
		uint32_t max_tiles = 0;
		std::vector<uint32_t> nts((size_t)n);
		for (int i = 0; i < n; i++) {
			nts[(size_t)i] = (g.metas[(size_t)i].npx + QOI_TILE_PX - 1) / QOI_TILE_PX;
			if (nts[(size_t)i] > max_tiles) max_tiles = nts[(size_t)i];
		}
		for (uint32_t k = 0; k < max_tiles; k++)
			for (int i = 0; i < n; i++)
				if (k < nts[(size_t)i]) g.tile_order.push_back((uint64_t)(g.metas[(size_t)i].first_tile + k) | ((uint64_t)(uint32_t)i << 32));
	}
	return 0;
}

#define QOI_STATE_ARRAYS 7
struct QoiStateLayout {
	size_t off[QOI_STATE_ARRAYS];
	size_t total;
};

static inline QoiStateLayout qoi_state_layout(uint32_t n_tiles) {
	QoiStateLayout L;
	size_t o = 0;
	const size_t T = n_tiles;
	const size_t sizes[QOI_STATE_ARRAYS] = {16, T * 512, T * 32, T * 8, T * 4, T * 4, T * 4};
	for (int i = 0; i < QOI_STATE_ARRAYS; i++) {
		L.off[i] = o;
		o += (sizes[i] + 255) & ~(size_t)255;
	}
	L.total = o;
	return L;
}

static inline void qoi_state_bind(QoiTileState &st, unsigned char *base, const QoiStateLayout &L) {
	st.ticket = (uint32_t *)(base + L.off[0]);
	st.ent = (uint64_t *)(base + L.off[1]);
	st.sum = (uint64_t *)(base + L.off[2]);
	st.byte_state = (uint64_t *)(base + L.off[3]);
	st.rs_px = (int32_t *)(base + L.off[4]);
	st.rs_byte = (int32_t *)(base + L.off[5]);
	st.rs_prev = (uint32_t *)(base + L.off[6]);
}

/* ------------------------------------------------------------------ foreign-stream decoder scratch (qoi_fdec.cuh) */
#define QOI_FSEG_BYTES_HOST 1024 /* must equal QOI_FSEG_BYTES */
#define QOI_FCHUNK_HOST 64       /* must equal QOI_FCHUNK: segments per scan chunk */
#define QOI_F_ARRAYS 33

struct QoiFLayout {
	std::vector<uint32_t> seg_base;   /* [n + 1] first segment of each image (capacity) */
	std::vector<uint32_t> chunk_base; /* [n + 1] first scan chunk of each image */
	uint32_t n_segs = 0, n_chunks = 0;
	size_t off[QOI_F_ARRAYS];
	size_t total = 0;
};

static inline QoiFLayout qoi_fstate_layout(const QoiPlanGeom &g) {
	QoiFLayout L;
	const size_t n = g.metas.size();
	L.seg_base.resize(n + 1);
	L.chunk_base.resize(n + 1);
	uint64_t segs = 0, chunks = 0;
	for (size_t i = 0; i < n; i++) {
		L.seg_base[i] = (uint32_t)segs;
		L.chunk_base[i] = (uint32_t)chunks;
		const uint64_t worst = (uint64_t)g.metas[i].npx * 5u; /* longest chunk area a decoder can consume */
		const uint64_t s_img = (worst + QOI_FSEG_BYTES_HOST - 1) / QOI_FSEG_BYTES_HOST + 1;
		segs += s_img;
		chunks += (s_img + QOI_FCHUNK_HOST - 1) / QOI_FCHUNK_HOST;
	}
	L.seg_base[n] = (uint32_t)segs;
	L.chunk_base[n] = (uint32_t)chunks;
	L.n_segs = (uint32_t)segs;
	L.n_chunks = (uint32_t)chunks;
	const size_t S = segs, N = n, C = chunks, TL = g.n_tiles;
	const size_t sizes[QOI_F_ARRAYS] = {(N + 1) * 4, (N + 1) * 4, S * 20, S * 8,      S,         S * 4,     N * 4,     S * 4,
	                                    S * 256,     S * 4,       S * 256, S * 8,      S * 8,     S * 4,     S,         S * 4,
	                                    64 * 4,      C * 256,     C * 8,   C * 8,      C * 256,   C * 4,     N * 4 * 4, N * 4 * 2,
	                                    S * 8,       S * 20,      S * 4,   TL * 8,     TL * 4,    N * 65 * 4,
	                                    S * 4,       64 * 4,      S * 2};
	size_t o = 0;
	for (int k = 0; k < QOI_F_ARRAYS; k++) {
		L.off[k] = o;
		o += (sizes[k] + 255) & ~(size_t)255;
	}
	L.total = o;
	return L;
}
