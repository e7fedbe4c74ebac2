/*
 * qoi_dec.cuh -- segment-parallel QOI decoder for sm_100a.
 *
 * Replaces the reference decode loop (qoi.h:540-587).  The op stream of one
 * image is inherently sequential (p, px, run and index[64] all chain), so the
 * unit of parallelism is a *segment*: a position in the stream at which the
 * complete decoder state is known.
 *
 *   foreign stream (use_index = 0): one segment = the whole image, state = the
 *       reference's initial state (qoi.h:533-537);
 *   stream written by qoi_encode_kernel (use_index = 1): one segment per
 *       encoder tile.  The chained scan already materialised the state at every
 *       tile boundary: index[] after tile t-1 is inc_tab[t-1], and rs_px /
 *       rs_byte / rs_prev give the first op that starts inside tile t, where it
 *       lies in the stream and the pixel before it.  The index is a side table
 *       in HBM; the stream itself carries nothing extra and stays bit exact.
 *
 * One LANE per segment (not one warp): each lane runs the scalar state machine
 * with its 64-entry table in shared memory, laid out [slot][thread] so that any
 * slot pattern across the lanes is bank-conflict free.  All lanes produce one
 * pixel per loop trip, so pixel stores stay converged.
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

template <int OUT_CH, int BLOCK>
__global__ void __launch_bounds__(BLOCK) qoi_decode_kernel(const QoiDecParams P) {
	QOI_DYNAMIC_SMEM(smem);
	uint32_t *tab = (uint32_t *)smem + threadIdx.x; /* tab[slot * BLOCK] */
	const uint32_t job = blockIdx.x * BLOCK + threadIdx.x;
	if (job >= P.n_jobs) return;

	/* ---- which image, which pixels ---- */
	uint32_t img;
	if (P.use_index) {
		uint32_t lo = 0, hi = P.n_images - 1;
		while (lo < hi) {
			const uint32_t mid = (lo + hi + 1) >> 1;
			if (P.imgs[mid].first_tile <= job) lo = mid;
			else hi = mid - 1;
		}
		img = lo;
	} else {
		img = job;
	}
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
	int32_t p = QOI_HEADER_BYTES;
	uint32_t px = 0xFF000000u; /* {0,0,0,255}, qoi.h:534-537 */
	if (P.use_index) {
		const int32_t b = P.st.rs_px[job];
		if (b < 0) return; /* the image ends before an op starts in this tile */
		px_begin = (uint32_t)b;
		p = P.st.rs_byte[job];
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
		const uint64_t valid = P.st.inc_valid[job - 1];
		const uint32_t *src = P.st.inc_tab + (size_t)(job - 1) * 64;
#pragma unroll 8
		for (int k = 0; k < 64; k++) tab[k * BLOCK] = ((valid >> k) & 1ull) ? src[k] : 0u;
		/* the reference DEcoder also stores the current colour after a RUN op (qoi.h:577); the ENcoder's
		   table does not see the leading run of the start colour.  Only slot 53 can differ, and only
		   until a pixel hashing to 53 is coded. */
		const uint32_t op0 = s[QOI_HEADER_BYTES];
		if (!((valid >> 53) & 1ull) && (op0 & 0xC0u) == 0xC0u && op0 < 0xFEu) tab[53 * BLOCK] = 0xFF000000u;
	}

	const int32_t stop = size - QOI_END_BYTES; /* qoi.h:539 */
	uint8_t *out = P.out + m.dec_off + (uint64_t)px_begin * OUT_CH;
	uint32_t run = 0;
	for (uint32_t i = px_begin; i < px_end; i++) {
		if (run > 0) {
			run--;
		} else if (p < stop) {
			const uint32_t b1 = s[p++];
			if (b1 == 0xFEu) { /* qoi.h:547-551 */
				px = (px & 0xFF000000u) | (uint32_t)s[p] | ((uint32_t)s[p + 1] << 8) | ((uint32_t)s[p + 2] << 16);
				p += 3;
			} else if (b1 == 0xFFu) { /* qoi.h:552-557 */
				px = (uint32_t)s[p] | ((uint32_t)s[p + 1] << 8) | ((uint32_t)s[p + 2] << 16) | ((uint32_t)s[p + 3] << 24);
				p += 4;
			} else if ((b1 & 0xC0u) == 0x00u) { /* qoi.h:558-560 */
				px = tab[b1 * BLOCK];
			} else if ((b1 & 0xC0u) == 0x40u) { /* qoi.h:561-565: +(-2..1) per channel, wrapping */
				const uint32_t add = ((((b1 >> 4) & 3u) + 254u) & 0xFFu) | (((((b1 >> 2) & 3u) + 254u) & 0xFFu) << 8) | ((((b1 & 3u) + 254u) & 0xFFu) << 16);
				px = __vadd4(px, add);
			} else if ((b1 & 0xC0u) == 0x80u) { /* qoi.h:566-572 */
				const uint32_t b2 = s[p++];
				const uint32_t vg = (b1 & 0x3Fu) + 224u; /* -32 mod 256 */
				const uint32_t add = ((vg + 248u + ((b2 >> 4) & 0xFu)) & 0xFFu) | ((vg & 0xFFu) << 8) | (((vg + 248u + (b2 & 0xFu)) & 0xFFu) << 16);
				px = __vadd4(px, add);
			} else { /* qoi.h:573-575 */
				run = b1 & 0x3Fu;
			}
			tab[qoi_slot(px) * BLOCK] = px; /* qoi.h:577 */
		}
		if (OUT_CH == 4) {
			*(uint32_t *)out = px;
		} else {
			out[0] = (uint8_t)px;
			out[1] = (uint8_t)(px >> 8);
			out[2] = (uint8_t)(px >> 16);
		}
		out += OUT_CH;
	}
}

}  // namespace qoidec
