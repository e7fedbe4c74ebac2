/*
 * qoi_gpu.cu -- C ABI of the B200-native QOI codec (libqoi_gpu.so).
 *
 * Exports (all extern "C", plain pointers and sizes):
 *   - the reference's own entry points  qoi_encode / qoi_decode   (reference qoi.h:278,289; bodies qoi.h:356-590)
 *   - the batch / device-resident API declared in include/qoi_gpu.h
 *
 * There is NO CPU implementation in this library.  Every pixel is coded by the sm_100a kernels in
 * qoi_enc.cuh / qoi_dec.cuh; if CUDA is unavailable the functions fail (NULL / QOI_GPU_E_CUDA).
 */
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <atomic>
#include <chrono>
#include <condition_variable>
#include <functional>
#include <mutex>
#include <new>
#include <thread>
#include <vector>

#include "../include/qoi_gpu.h"
#include "qoi_dec.cuh"
#include "qoi_enc.cuh"
#include "qoi_fdec.cuh"
#include "qoi_plan.h"

static_assert(sizeof(qoi_desc) == sizeof(QoiDescPOD), "qoi_desc layout");
static_assert(QOI_GPU_TILE_PX == QOI_TILE_PX, "tile size");

static std::atomic<uint64_t> g_launches{0};

struct qoi_gpu_ctx {
	int device = 0;
	int sm_count = 0;
	cudaStream_t stream = nullptr;                 /* kernels of the host-buffer entry points */
	cudaStream_t s_h2d = nullptr, s_d2h = nullptr; /* one stream per copy direction: PCIe is full duplex */
	char err[256] = {0};
	int enc_blocks_per_sm[3] = {0, 0, 0}; /* resident CTAs per SM of qoi_encode_kernel<4,false>, <4,true>, <3,true> */
	/* pinned staging of the four reference entry points (pageable caller memory <-> device), grown on demand */
	unsigned char *h_stage = nullptr;
	size_t h_stage_bytes = 0;
	std::vector<cudaEvent_t> stage_events;
};

struct qoi_gpu_plan {
	qoi_gpu_ctx *ctx = nullptr;
	QoiPlanGeom geom;
	QoiStateLayout layout;
	QoiImgMeta *d_metas = nullptr;
	uint32_t *d_tile_img = nullptr;
	uint64_t *d_tile_order = nullptr;
	unsigned char *d_state = nullptr;
	QoiTileState st;
	uint32_t epoch = 0;
	uint32_t fepoch = 0; /* any-stream decodes on this plan (QoiFParams::epoch) */
	bool pix_off_aligned16 = true;
	/* lazily allocated device mirrors for the *_host entry points */
	unsigned char *d_pixels = nullptr, *d_streams = nullptr, *d_decoded = nullptr;
	int32_t *d_lens = nullptr;
	int32_t *h_lens_pinned = nullptr;
	/* the host-buffer entry points cut the batch into chunks (sub-plans over the same buffers) and pipeline
	   H2D copy / kernel / D2H copy of successive chunks */
	std::vector<qoi_gpu_plan *> chunks;
	std::vector<int> chunk_first;
	std::vector<cudaEvent_t> ev_in, ev_k;
	int index_owner = 0; /* who holds the restart index of the last encode: 0 nobody, 1 this plan, 2 the chunks */
	/* scratch of the foreign-stream decoder (qoi_fdec.cuh), allocated on first use */
	unsigned char *d_fstate = nullptr;
	QoiFLayout flayout;
};

/* every entry point runs on its context's device and leaves the caller's current device as it found it */
struct DeviceGuard {
	int prev = -1;
	bool changed = false;
	explicit DeviceGuard(int want) {
		if (cudaGetDevice(&prev) == cudaSuccess && prev != want) changed = cudaSetDevice(want) == cudaSuccess;
	}
	~DeviceGuard() {
		if (changed) cudaSetDevice(prev);
	}
	DeviceGuard(const DeviceGuard &) = delete;
	DeviceGuard &operator=(const DeviceGuard &) = delete;
};

#define CK(ctx, call)                                                                         \
	do {                                                                                      \
		cudaError_t e_ = (call);                                                              \
		if (e_ != cudaSuccess) {                                                              \
			if (ctx) snprintf((ctx)->err, sizeof((ctx)->err), "%s: %s", #call, cudaGetErrorString(e_)); \
			return QOI_GPU_E_CUDA;                                                            \
		}                                                                                     \
	} while (0)

static_assert(QOI_FFALL_SMEM <= 48 * 1024 && QOI_FCHAIN_SMEM <= 48 * 1024 && QOI_FCHUNK_HOST == QOI_FCHUNK && QOI_FROUNDS < 60, "these kernels are launched without a shared-memory opt-in");
/* the state-round blocks need more than the default 48 KB of dynamic shared memory */
static int fdec_setup(qoi_gpu_ctx *ctx) {
	if (QOI_FDEC_SMEM > 48 * 1024) CK(ctx, cudaFuncSetAttribute(qoifdec::fstate_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, QOI_FDEC_SMEM));
	return QOI_GPU_OK;
}

template <int CH, bool EXPAND>
static int enc_occupancy(qoi_gpu_ctx *ctx, int *out) {
	using L = qoienc::Layout<CH, EXPAND>;
	CK(ctx, cudaFuncSetAttribute(qoienc::qoi_encode_kernel<CH, EXPAND>, cudaFuncAttributeMaxDynamicSharedMemorySize, L::TOTAL));
	/* many small CTAs per SM: ask for the largest shared-memory carve-out */
	CK(ctx, cudaFuncSetAttribute(qoienc::qoi_encode_kernel<CH, EXPAND>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
	int nb = 0;
	CK(ctx, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, qoienc::qoi_encode_kernel<CH, EXPAND>, QOI_ENC_THREADS, L::TOTAL));
	*out = nb > 0 ? nb : 1; /* the grid must not exceed what is co-resident (look-back) */
	return QOI_GPU_OK;
}

template <int OUT_CH, int BLOCK>
static void launch_decode(const QoiDecParams &D, cudaStream_t s) {
	const uint32_t grid = (D.n_jobs + BLOCK - 1) / BLOCK;
	qoidec::qoi_decode_kernel<OUT_CH, BLOCK><<<grid, BLOCK, QOI_DEC_SMEM_BYTES(BLOCK), s>>>(D);
}

extern "C" {

int qoi_gpu_device_count(void) {
	int n = 0;
	if (cudaGetDeviceCount(&n) != cudaSuccess) return 0;
	return n;
}

uint64_t qoi_gpu_launch_count(void) { return g_launches.load(); }

int qoi_gpu_ctx_create(qoi_gpu_ctx **out, int device) {
	if (!out) return QOI_GPU_E_ARG;
	*out = nullptr;
	int n = qoi_gpu_device_count();
	if (device < 0 || device >= n) return QOI_GPU_E_CUDA;
	qoi_gpu_ctx *c = new (std::nothrow) qoi_gpu_ctx;
	if (!c) return QOI_GPU_E_NOMEM;
	c->device = device;
	DeviceGuard guard(device);
	cudaError_t e = cudaSuccess;
	if (e == cudaSuccess) e = cudaDeviceGetAttribute(&c->sm_count, cudaDevAttrMultiProcessorCount, device);
	if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking);
	if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&c->s_h2d, cudaStreamNonBlocking);
	if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&c->s_d2h, cudaStreamNonBlocking);
	if (e != cudaSuccess) {
		qoi_gpu_ctx_destroy(c);
		return QOI_GPU_E_CUDA;
	}
	int rc = enc_occupancy<4, false>(c, &c->enc_blocks_per_sm[0]);
	if (!rc) rc = enc_occupancy<4, true>(c, &c->enc_blocks_per_sm[1]);
	if (!rc) rc = enc_occupancy<3, true>(c, &c->enc_blocks_per_sm[2]);
	if (!rc) rc = fdec_setup(c);
	if (rc) {
		fprintf(stderr, "qoi_gpu: kernel setup failed: %s\n", c->err);
		qoi_gpu_ctx_destroy(c);
		return rc;
	}
	*out = c;
	return QOI_GPU_OK;
}

void qoi_gpu_ctx_destroy(qoi_gpu_ctx *c) {
	if (!c) return;
	DeviceGuard guard(c->device);
	if (c->stream) cudaStreamDestroy(c->stream);
	if (c->s_h2d) cudaStreamDestroy(c->s_h2d);
	if (c->s_d2h) cudaStreamDestroy(c->s_d2h);
	if (c->h_stage) cudaFreeHost(c->h_stage);
	for (cudaEvent_t e : c->stage_events) cudaEventDestroy(e);
	delete c;
}

const char *qoi_gpu_last_error(const qoi_gpu_ctx *c) { return c ? c->err : ""; }

size_t qoi_gpu_max_stream_size(const qoi_desc *d) {
	if (!d) return 0;
	QoiDescPOD p = {d->width, d->height, d->channels, d->colorspace};
	return qoi_max_stream(p);
}

int qoi_gpu_plan_create(qoi_gpu_ctx *ctx, qoi_gpu_plan **out, int n, const qoi_desc *descs, const uint64_t *pixel_off,
                        const uint64_t *stream_off, const uint64_t *decoded_off, int decoded_channels) {
	if (!ctx || !out) return QOI_GPU_E_ARG;
	*out = nullptr;
	qoi_gpu_plan *p = new (std::nothrow) qoi_gpu_plan;
	if (!p) return QOI_GPU_E_NOMEM;
	p->ctx = ctx;
	if (qoi_plan_geometry(p->geom, n, (const QoiDescPOD *)descs, pixel_off, stream_off, decoded_off, decoded_channels)) {
		delete p;
		return QOI_GPU_E_ARG;
	}
	for (const QoiImgMeta &m : p->geom.metas)
		if (m.pix_off & 15u) p->pix_off_aligned16 = false;
	p->layout = qoi_state_layout(p->geom.n_tiles);
	DeviceGuard guard(ctx->device);
	cudaError_t e = cudaMalloc(&p->d_metas, sizeof(QoiImgMeta) * (size_t)n);
	if (e == cudaSuccess) e = cudaMalloc(&p->d_tile_img, sizeof(uint32_t) * (size_t)p->geom.n_tiles);
	if (e == cudaSuccess) e = cudaMemcpy(p->d_tile_img, p->geom.tile_img.data(), sizeof(uint32_t) * (size_t)p->geom.n_tiles, cudaMemcpyHostToDevice);
	if (e == cudaSuccess) e = cudaMalloc(&p->d_tile_order, sizeof(uint64_t) * (size_t)p->geom.n_tiles);
	if (e == cudaSuccess) e = cudaMemcpy(p->d_tile_order, p->geom.tile_order.data(), sizeof(uint64_t) * (size_t)p->geom.n_tiles, cudaMemcpyHostToDevice);
	if (e == cudaSuccess) e = cudaMalloc(&p->d_state, p->layout.total);
	if (e == cudaSuccess) e = cudaMemset(p->d_state, 0, p->layout.total);
	if (e == cudaSuccess) e = cudaMemcpy(p->d_metas, p->geom.metas.data(), sizeof(QoiImgMeta) * (size_t)n, cudaMemcpyHostToDevice);
	/* the copies above ran on the legacy stream; the kernels run on non-blocking streams that are not ordered against it */
	if (e == cudaSuccess) e = cudaDeviceSynchronize();
	if (e != cudaSuccess) {
		snprintf(ctx->err, sizeof ctx->err, "plan allocation: %s", cudaGetErrorString(e));
		qoi_gpu_plan_destroy(p);
		return e == cudaErrorMemoryAllocation ? QOI_GPU_E_NOMEM : QOI_GPU_E_CUDA;
	}
	qoi_state_bind(p->st, p->d_state, p->layout);
	*out = p;
	return QOI_GPU_OK;
}

void qoi_gpu_plan_destroy(qoi_gpu_plan *p) {
	if (!p) return;
	DeviceGuard guard(p->ctx->device);
	cudaFree(p->d_metas);
	cudaFree(p->d_tile_img);
	cudaFree(p->d_fstate);
	cudaFree(p->d_tile_order);
	cudaFree(p->d_state);
	cudaFree(p->d_pixels);
	cudaFree(p->d_streams);
	cudaFree(p->d_decoded);
	cudaFree(p->d_lens);
	if (p->h_lens_pinned) cudaFreeHost(p->h_lens_pinned);
	for (qoi_gpu_plan *c : p->chunks) qoi_gpu_plan_destroy(c);
	for (cudaEvent_t e : p->ev_in) cudaEventDestroy(e);
	for (cudaEvent_t e : p->ev_k) cudaEventDestroy(e);
	delete p;
}

uint64_t qoi_gpu_plan_pixels(const qoi_gpu_plan *p) { return p ? p->geom.total_px : 0; }
uint32_t qoi_gpu_plan_tiles(const qoi_gpu_plan *p) { return p ? p->geom.n_tiles : 0; }

int qoi_gpu_encode(qoi_gpu_plan *p, const void *d_pixels, void *d_streams, int32_t *d_lens, void *cuda_stream) {
	if (!p || !d_pixels || !d_streams || !d_lens) return QOI_GPU_E_ARG;
	qoi_gpu_ctx *ctx = p->ctx;
	cudaStream_t s = (cudaStream_t)cuda_stream;
	DeviceGuard guard(ctx->device);
	if (++p->epoch >= (1u << 30) - 1u) { /* status words carry a 30-bit epoch */
		CK(ctx, cudaMemsetAsync(p->d_state, 0, p->layout.total, s));
		p->epoch = 1;
	}
	QoiEncParams P;
	memset(&P, 0, sizeof P);
	P.pixels = (const uint8_t *)d_pixels;
	P.streams = (uint8_t *)d_streams;
	P.lens = d_lens;
	P.imgs = p->d_metas;
	P.tile_img = p->d_tile_img;
	P.tile_order = p->d_tile_order;
	P.n_images = (uint32_t)p->geom.metas.size();
	P.n_tiles = p->geom.n_tiles;
	P.epoch = p->epoch;
	P.st = p->st;
	const bool zero_copy = p->geom.channels == 4 && p->pix_off_aligned16 && (((uintptr_t)d_pixels & 15u) == 0);
	const int variant = p->geom.channels == 3 ? 2 : (zero_copy ? 0 : 1);
	uint32_t grid = (uint32_t)(ctx->sm_count * ctx->enc_blocks_per_sm[variant]);
	if (grid > P.n_tiles) grid = P.n_tiles;
	if (variant == 0) qoienc::qoi_encode_kernel<4, false><<<grid, QOI_ENC_THREADS, qoienc::Layout<4, false>::TOTAL, s>>>(P);
	else if (variant == 1) qoienc::qoi_encode_kernel<4, true><<<grid, QOI_ENC_THREADS, qoienc::Layout<4, true>::TOTAL, s>>>(P);
	else qoienc::qoi_encode_kernel<3, true><<<grid, QOI_ENC_THREADS, qoienc::Layout<3, true>::TOTAL, s>>>(P);
	g_launches++;
	CK(ctx, cudaGetLastError());
	p->index_owner = 1;
	return QOI_GPU_OK;
}

int qoi_gpu_decode(qoi_gpu_plan *p, const void *d_streams, const int32_t *d_lens, void *d_decoded, int use_index,
                   int32_t *d_status, void *cuda_stream) {
	if (!p || !d_streams || !d_lens || !d_decoded) return QOI_GPU_E_ARG;
	/* the restart index must be the one the last encode of THESE streams left in THIS plan (qoi_gpu_encode, or an index
	   imported with qoi_gpu_index_import); qoi_gpu_encode_host leaves it in the chunk sub-plans instead */
	if (use_index && p->index_owner != 1) return QOI_GPU_E_ARG;
	qoi_gpu_ctx *ctx = p->ctx;
	cudaStream_t s = (cudaStream_t)cuda_stream;
	DeviceGuard guard(ctx->device);
	if (!use_index) {
		/* any stream: build the restart index (parse, state rounds, scans, fall-back; qoi_fdec.cuh), then decode with it below */
		static_assert(QOI_FSEG_BYTES_HOST == QOI_FSEG_BYTES, "segment size");
		if (!p->d_fstate) {
			p->flayout = qoi_fstate_layout(p->geom);
			CK(ctx, cudaMalloc(&p->d_fstate, p->flayout.total));
			/* on the launching stream, so that the kernels below are ordered behind the initialisation (the host arrays live in the plan) */
			CK(ctx, cudaMemsetAsync(p->d_fstate, 0, p->flayout.total, s));
			CK(ctx, cudaMemcpyAsync(p->d_fstate + p->flayout.off[0], p->flayout.seg_base.data(), p->flayout.seg_base.size() * sizeof(uint32_t),
			                        cudaMemcpyHostToDevice, s));
			CK(ctx, cudaMemcpyAsync(p->d_fstate + p->flayout.off[1], p->flayout.chunk_base.data(), p->flayout.chunk_base.size() * sizeof(uint32_t),
			                        cudaMemcpyHostToDevice, s));
		}
		qoifdec::QoiFParams F;
		memset(&F, 0, sizeof F);
		F.streams = (const uint8_t *)d_streams;
		F.lens = d_lens;
		F.status = d_status;
		F.imgs = p->d_metas;
		F.n_images = (uint32_t)p->geom.metas.size();
		F.n_segs = p->flayout.n_segs;
		F.n_chunks = p->flayout.n_chunks;
		F.st = p->st;
		F.epoch = (++p->fepoch) & 0x00FFFFFFu;
		qoifdec::fstate_bind(F.f, p->d_fstate, p->flayout.off);
		const uint32_t gs = (F.n_segs + QOI_FDEC_THREADS - 1) / QOI_FDEC_THREADS;
		qoifdec::fparse_kernel<<<gs, QOI_FDEC_THREADS, 0, s>>>(F);
		qoifdec::fchain_kernel<<<F.n_images, QOI_FCHAIN_THREADS, QOI_FCHAIN_SMEM, s>>>(F);
		qoifdec::fzero_kernel<<<gs, QOI_FDEC_THREADS, 0, s>>>(F);
		/* rounds 0..2 move most of the segments; later rounds usually find every stream settled (each launch then returns at
		   once: active[]) or walk a handful of segments: a grid of a few CTAs per SM striding over the work keeps an idle
		   round at a few microseconds */
		const uint32_t small = (uint32_t)ctx->sm_count * 4u;
		for (int r = 0; r < QOI_FROUNDS; r++) {
			F.round = (uint32_t)r;
			const int last = r + 1 == QOI_FROUNDS;
			const uint32_t gst = (F.n_segs + QOI_FSTATE_THREADS - 1) / QOI_FSTATE_THREADS;
			const uint32_t g_st = r < 3 ? gst : (gst < small ? gst : small), g_chk = r < 3 ? F.n_chunks : (F.n_chunks < small ? F.n_chunks : small);
			qoifdec::fstate_kernel<<<g_st, QOI_FSTATE_THREADS, QOI_FDEC_SMEM, s>>>(F);
			qoifdec::fscan_a_kernel<<<g_chk, QOI_FSCAN_THREADS, QOI_FMETA_SMEM, s>>>(F);
			qoifdec::fscan_b_kernel<<<F.n_images, QOI_FSCANB_THREADS, QOI_FSCANB_SMEM, s>>>(F);
			qoifdec::fverify_kernel<<<g_chk, QOI_FSCAN_THREADS, QOI_FVERIFY_SMEM, s>>>(F, last);
		}
		qoifdec::ffinal_kernel<<<F.n_chunks, QOI_FSCAN_THREADS, QOI_FMETA_SMEM, s>>>(F);
		qoifdec::ffall_kernel<<<F.n_images, 32, QOI_FFALL_SMEM, s>>>(F);
		qoifdec::fcompose_kernel<<<(p->geom.n_tiles * 64u + 127u) / 128u, 128, 0, s>>>(F, p->geom.n_tiles, p->d_tile_img);
		g_launches += 6 + 4 * QOI_FROUNDS;
		CK(ctx, cudaGetLastError());
		p->index_owner = 1; /* the plan now holds the restart index of THESE streams (it can be exported as a side-car) */
		d_status = nullptr; /* reported by fchain_kernel */
	}
	QoiDecParams D;
	memset(&D, 0, sizeof D);
	D.streams = (const uint8_t *)d_streams;
	D.lens = d_lens;
	D.out = (uint8_t *)d_decoded;
	D.status = d_status;
	D.imgs = p->d_metas;
	D.tile_img = p->d_tile_img;
	D.n_images = (uint32_t)p->geom.metas.size();
	D.n_jobs = p->geom.n_tiles;
	D.out_channels = p->geom.dec_channels;
	D.use_index = 1;
	D.st = p->st;
	/* few segments: narrow blocks so that they spread over the SMs */
	const bool narrow = D.n_jobs < (uint32_t)ctx->sm_count * 128u * 2u;
	if (D.out_channels == 4) {
		if (narrow) launch_decode<4, 32>(D, s);
		else launch_decode<4, 128>(D, s);
	} else {
		if (narrow) launch_decode<3, 32>(D, s);
		else launch_decode<3, 128>(D, s);
	}
	g_launches++;
	CK(ctx, cudaGetLastError());
	return QOI_GPU_OK;
}

int qoi_gpu_plan_fstats(qoi_gpu_plan *p, uint32_t out[128]) {
	if (!p || !out || !p->d_fstate) return QOI_GPU_E_ARG;
	qoi_gpu_ctx *ctx = p->ctx;
	DeviceGuard guard(ctx->device);
	qoifdec::QoiFState f;
	qoifdec::fstate_bind(f, p->d_fstate, p->flayout.off);
	CK(ctx, cudaDeviceSynchronize());
	CK(ctx, cudaMemcpy(out, f.n_list, 64 * sizeof(uint32_t), cudaMemcpyDeviceToHost));
	CK(ctx, cudaMemcpy(out + 64, f.active, 64 * sizeof(uint32_t), cudaMemcpyDeviceToHost));
	/* out[120..127]: the first images whose stream went to the fall-back (1 + image index; 0: none) */
	const size_t n = p->geom.metas.size();
	std::vector<uint32_t> planes(4 * n), fbad(2 * n), live(n);
	CK(ctx, cudaMemcpy(planes.data(), f.bad_cnt, 4 * n * sizeof(uint32_t), cudaMemcpyDeviceToHost));
	CK(ctx, cudaMemcpy(fbad.data(), f.first_bad, 2 * n * sizeof(uint32_t), cudaMemcpyDeviceToHost));
	CK(ctx, cudaMemcpy(live.data(), f.n_live, n * sizeof(uint32_t), cudaMemcpyDeviceToHost));
	int k = 0;
	for (int i = 120; i < 128; i++) out[i] = 0;
	for (size_t i = 0; i < n && k < 8; i++) {
		const uint32_t par = planes[3 * n + i] & 1u;
		if ((fbad[par * n + i] >> 1) < live[i]) out[120 + k++] = (uint32_t)i + 1u;
	}
	return QOI_GPU_OK;
}

/* ------------------------------------------------------------------ restart index as a side-car */
struct QoiIndexHeader {
	uint32_t magic;   /* "QOIX" */
	uint32_t version; /* 1 */
	uint32_t tile_px;
	uint32_t n_images;
	uint32_t n_tiles;
	uint32_t reserved;
};
struct QoiIndexImage {
	uint32_t width, height;
	uint32_t channels_cs; /* channels | colorspace << 8 */
	uint32_t first_tile;
};
#define QOI_INDEX_MAGIC 0x58494F51u /* 'Q','O','I','X' little endian */

/* blob: header, n_images x QoiIndexImage, rs_px[n_tiles], rs_byte[n_tiles], rs_prev[n_tiles], index[n_tiles][64] (u32 values) */
size_t qoi_gpu_index_bytes(const qoi_gpu_plan *p) {
	if (!p) return 0;
	return sizeof(QoiIndexHeader) + p->geom.metas.size() * sizeof(QoiIndexImage) + (size_t)p->geom.n_tiles * (12 + 256);
}

int qoi_gpu_index_export(qoi_gpu_plan *p, void *h_index, size_t capacity, void *cuda_stream) {
	if (!p || !h_index || capacity < qoi_gpu_index_bytes(p)) return QOI_GPU_E_ARG;
	if (p->index_owner != 1) return QOI_GPU_E_ARG; /* no encode (or import) has left an index in this plan */
	qoi_gpu_ctx *ctx = p->ctx;
	cudaStream_t s = (cudaStream_t)cuda_stream;
	DeviceGuard guard(ctx->device);
	const size_t T = p->geom.n_tiles, N = p->geom.metas.size();
	unsigned char *o = (unsigned char *)h_index;
	QoiIndexHeader h = {QOI_INDEX_MAGIC, 1u, QOI_TILE_PX, (uint32_t)N, (uint32_t)T, 0u};
	memcpy(o, &h, sizeof h);
	o += sizeof h;
	for (size_t i = 0; i < N; i++) {
		const QoiImgMeta &m = p->geom.metas[i];
		QoiIndexImage im = {m.width, m.height, (uint32_t)m.channels | ((uint32_t)m.colorspace << 8), m.first_tile};
		memcpy(o, &im, sizeof im);
		o += sizeof im;
	}
	std::vector<uint64_t> ent(T * 64);
	CK(ctx, cudaMemcpyAsync(o, p->st.rs_px, T * 4, cudaMemcpyDeviceToHost, s));
	CK(ctx, cudaMemcpyAsync(o + T * 4, p->st.rs_byte, T * 4, cudaMemcpyDeviceToHost, s));
	CK(ctx, cudaMemcpyAsync(o + T * 8, p->st.rs_prev, T * 4, cudaMemcpyDeviceToHost, s));
	CK(ctx, cudaMemcpyAsync(ent.data(), p->st.ent, T * 512, cudaMemcpyDeviceToHost, s));
	CK(ctx, cudaStreamSynchronize(s));
	uint32_t *vals = (uint32_t *)(o + T * 12); /* the blob's arrays are 4-byte aligned if h_index is */
	if ((uintptr_t)vals & 3u) {
		for (size_t k = 0; k < T * 64; k++) {
			const uint32_t v = (uint32_t)(ent[k] >> 32);
			memcpy((unsigned char *)vals + 4 * k, &v, 4);
		}
	} else {
		for (size_t k = 0; k < T * 64; k++) vals[k] = (uint32_t)(ent[k] >> 32);
	}
	return QOI_GPU_OK;
}

int qoi_gpu_index_import(qoi_gpu_plan *p, const void *h_index, size_t bytes, void *cuda_stream) {
	if (!p || !h_index || bytes < sizeof(QoiIndexHeader)) return QOI_GPU_E_ARG;
	const unsigned char *o = (const unsigned char *)h_index;
	QoiIndexHeader h;
	memcpy(&h, o, sizeof h);
	const size_t T = p->geom.n_tiles, N = p->geom.metas.size();
	if (h.magic != QOI_INDEX_MAGIC || h.version != 1u || h.tile_px != QOI_TILE_PX || h.n_images != N || h.n_tiles != T) return QOI_GPU_E_ARG;
	if (bytes < qoi_gpu_index_bytes(p)) return QOI_GPU_E_ARG;
	o += sizeof h;
	for (size_t i = 0; i < N; i++) {
		const QoiImgMeta &m = p->geom.metas[i];
		QoiIndexImage im;
		memcpy(&im, o, sizeof im);
		o += sizeof im;
		if (im.width != m.width || im.height != m.height || im.channels_cs != ((uint32_t)m.channels | ((uint32_t)m.colorspace << 8)) ||
		    im.first_tile != m.first_tile)
			return QOI_GPU_E_ARG;
	}
	/* restart points must lie inside their image (a damaged side-car must not send a lane outside the buffers) */
	std::vector<int32_t> rs((size_t)T * 2);
	memcpy(rs.data(), o, T * 8);
	for (size_t i = 0; i < N; i++) {
		const QoiImgMeta &m = p->geom.metas[i];
		const size_t nt = ((size_t)m.npx + QOI_TILE_PX - 1) / QOI_TILE_PX;
		const QoiDescPOD d = {m.width, m.height, m.channels, m.colorspace};
		for (size_t k = 0; k < nt; k++) {
			const int32_t px = rs[m.first_tile + k], by = rs[T + m.first_tile + k];
			if (px < 0) continue; /* no op starts in this tile */
			if ((uint32_t)px < k * QOI_TILE_PX || (uint32_t)px >= m.npx || by < QOI_HEADER_BYTES || (size_t)by >= qoi_max_foreign_stream(d)) return QOI_GPU_E_ARG;
		}
	}
	qoi_gpu_ctx *ctx = p->ctx;
	cudaStream_t s = (cudaStream_t)cuda_stream;
	DeviceGuard guard(ctx->device);
	std::vector<uint64_t> ent(T * 64);
	const unsigned char *vals = o + T * 12;
	for (size_t k = 0; k < T * 64; k++) {
		uint32_t v;
		memcpy(&v, vals + 4 * k, 4);
		ent[k] = ((uint64_t)v << 32) | 1ull; /* epoch field 0: a later encode on this plan treats the words as unpublished */
	}
	CK(ctx, cudaMemcpyAsync(p->st.rs_px, o, T * 4, cudaMemcpyHostToDevice, s));
	CK(ctx, cudaMemcpyAsync(p->st.rs_byte, o + T * 4, T * 4, cudaMemcpyHostToDevice, s));
	CK(ctx, cudaMemcpyAsync(p->st.rs_prev, o + T * 8, T * 4, cudaMemcpyHostToDevice, s));
	CK(ctx, cudaMemcpyAsync(p->st.ent, ent.data(), T * 512, cudaMemcpyHostToDevice, s));
	CK(ctx, cudaStreamSynchronize(s));
	p->index_owner = 1;
	return QOI_GPU_OK;
}

/* ------------------------------------------------------------------ host-buffer forms */
static int plan_mirrors(qoi_gpu_plan *p) {
	qoi_gpu_ctx *ctx = p->ctx;
	if (!p->d_pixels) CK(ctx, cudaMalloc(&p->d_pixels, p->geom.pixel_bytes + 16));
	if (!p->d_streams) CK(ctx, cudaMalloc(&p->d_streams, p->geom.stream_bytes + 16));
	if (!p->d_decoded) CK(ctx, cudaMalloc(&p->d_decoded, p->geom.decoded_bytes + 16));
	if (!p->d_lens) CK(ctx, cudaMalloc(&p->d_lens, sizeof(int32_t) * p->geom.metas.size()));
	if (!p->h_lens_pinned) CK(ctx, cudaMallocHost(&p->h_lens_pinned, sizeof(int32_t) * p->geom.metas.size()));
	return QOI_GPU_OK;
}

#define QOI_HOST_CHUNKS 8

/* sub-plans over the same buffers, one per chunk of consecutive images */
static int plan_chunks(qoi_gpu_plan *p) {
	if (!p->chunks.empty()) return QOI_GPU_OK;
	qoi_gpu_ctx *ctx = p->ctx;
	const int n = (int)p->geom.metas.size();
	int want = QOI_HOST_CHUNKS;
	{
		const char *e = getenv("QOI_GPU_HOST_CHUNKS"); /* A/B knob: more chunks = shorter pipeline fill and drain, smaller launches */
		if (e && atoi(e) > 0 && atoi(e) <= 64) want = atoi(e);
	}
	const int k = n < want ? n : want;
	std::vector<qoi_desc> descs((size_t)n);
	std::vector<uint64_t> po((size_t)n), so((size_t)n), dof((size_t)n);
	for (int i = 0; i < n; i++) {
		const QoiImgMeta &m = p->geom.metas[(size_t)i];
		memset(&descs[(size_t)i], 0, sizeof(qoi_desc));
		descs[(size_t)i].width = m.width; descs[(size_t)i].height = m.height;
		descs[(size_t)i].channels = m.channels; descs[(size_t)i].colorspace = m.colorspace;
		po[(size_t)i] = m.pix_off; so[(size_t)i] = m.stream_off; dof[(size_t)i] = m.dec_off;
	}
	/* built into locals and committed only when every piece exists: a half-built list would be taken for complete by the next call */
	std::vector<qoi_gpu_plan *> subs;
	std::vector<int> firsts;
	std::vector<cudaEvent_t> e_in, e_k;
	int rc = QOI_GPU_OK;
	for (int c = 0; c < k && !rc; c++) {
		const int first = (int)((int64_t)n * c / k), last = (int)((int64_t)n * (c + 1) / k);
		qoi_gpu_plan *sub = nullptr;
		rc = qoi_gpu_plan_create(ctx, &sub, last - first, descs.data() + first, po.data() + first, so.data() + first, dof.data() + first,
		                         p->geom.dec_channels);
		if (rc) break;
		subs.push_back(sub);
		firsts.push_back(first);
		cudaEvent_t e1 = nullptr, e2 = nullptr;
		if (cudaEventCreateWithFlags(&e1, cudaEventDisableTiming) != cudaSuccess || cudaEventCreateWithFlags(&e2, cudaEventDisableTiming) != cudaSuccess) {
			if (e1) cudaEventDestroy(e1);
			rc = QOI_GPU_E_CUDA;
			break;
		}
		e_in.push_back(e1);
		e_k.push_back(e2);
	}
	if (rc) {
		for (qoi_gpu_plan *sp : subs) qoi_gpu_plan_destroy(sp);
		for (cudaEvent_t e : e_in) cudaEventDestroy(e);
		for (cudaEvent_t e : e_k) cudaEventDestroy(e);
		return rc;
	}
	firsts.push_back(n);
	p->chunks.swap(subs);
	p->chunk_first.swap(firsts);
	p->ev_in.swap(e_in);
	p->ev_k.swap(e_k);
	return QOI_GPU_OK;
}

int qoi_gpu_encode_host(qoi_gpu_plan *p, const void *h_pixels, void *h_streams, int32_t *h_lens) {
	if (!p || !h_pixels || !h_streams || !h_lens) return QOI_GPU_E_ARG;
	qoi_gpu_ctx *ctx = p->ctx;
	DeviceGuard guard(ctx->device);
	int rc = plan_mirrors(p);
	if (!rc) rc = plan_chunks(p);
	if (rc) return rc;
	const size_t nc = p->chunks.size();
	/* queue everything that does not need the lengths: copy in, encode, lengths out -- chunk after chunk */
	for (size_t c = 0; c < nc; c++) {
		const int first = p->chunk_first[c], last = p->chunk_first[c + 1];
		uint64_t lo = ~0ull, hi = 0;
		for (int i = first; i < last; i++) {
			const QoiImgMeta &m = p->geom.metas[(size_t)i];
			const uint64_t e = m.pix_off + (uint64_t)m.npx * m.channels;
			if (m.pix_off < lo) lo = m.pix_off;
			if (e > hi) hi = e;
		}
		CK(ctx, cudaMemcpyAsync(p->d_pixels + lo, (const unsigned char *)h_pixels + lo, (size_t)(hi - lo), cudaMemcpyHostToDevice, ctx->s_h2d));
		CK(ctx, cudaEventRecord(p->ev_in[c], ctx->s_h2d));
		CK(ctx, cudaStreamWaitEvent(ctx->stream, p->ev_in[c], 0));
		rc = qoi_gpu_encode(p->chunks[c], p->d_pixels, p->d_streams, p->d_lens + first, ctx->stream);
		if (rc) return rc;
		CK(ctx, cudaMemcpyAsync(p->h_lens_pinned + first, p->d_lens + first, sizeof(int32_t) * (size_t)(last - first), cudaMemcpyDeviceToHost, ctx->stream));
		CK(ctx, cudaEventRecord(p->ev_k[c], ctx->stream));
	}
	/* streams out as soon as a chunk's lengths are known; these copies overlap the uploads and kernels of later chunks */
	for (size_t c = 0; c < nc; c++) {
		CK(ctx, cudaEventSynchronize(p->ev_k[c]));
		for (int i = p->chunk_first[c]; i < p->chunk_first[c + 1]; i++) {
			const QoiImgMeta &m = p->geom.metas[(size_t)i];
			h_lens[i] = p->h_lens_pinned[i];
			CK(ctx, cudaMemcpyAsync((unsigned char *)h_streams + m.stream_off, p->d_streams + m.stream_off, (size_t)h_lens[i],
			                        cudaMemcpyDeviceToHost, ctx->s_d2h));
		}
	}
	CK(ctx, cudaStreamSynchronize(ctx->s_d2h));
	p->index_owner = 2;
	return QOI_GPU_OK;
}

int qoi_gpu_decode_host(qoi_gpu_plan *p, const void *h_streams, const int32_t *h_lens, void *h_decoded, int use_index) {
	if (!p || !h_streams || !h_lens || !h_decoded) return QOI_GPU_E_ARG;
	if (use_index && p->index_owner == 0) return QOI_GPU_E_ARG;
	qoi_gpu_ctx *ctx = p->ctx;
	DeviceGuard guard(ctx->device);
	int rc = plan_mirrors(p);
	if (!rc) rc = plan_chunks(p);
	if (rc) return rc;
	const size_t n = p->geom.metas.size();
	for (size_t i = 0; i < n; i++) {
		const QoiImgMeta &m = p->geom.metas[i];
		QoiDescPOD d = {m.width, m.height, m.channels, m.colorspace};
		if (h_lens[i] < QOI_HEADER_BYTES + QOI_END_BYTES || (size_t)h_lens[i] > qoi_max_foreign_stream(d)) return QOI_GPU_E_ARG;
		p->h_lens_pinned[i] = h_lens[i];
	}
	CK(ctx, cudaMemcpyAsync(p->d_lens, p->h_lens_pinned, sizeof(int32_t) * n, cudaMemcpyHostToDevice, ctx->s_h2d));
	const bool whole = use_index && p->index_owner == 1; /* the index lives in this plan: one launch over the batch */
	const size_t nc = p->chunks.size();
	for (size_t c = 0; c < nc; c++) {
		const int first = p->chunk_first[c], last = p->chunk_first[c + 1];
		uint64_t lo = ~0ull, hi = 0;
		for (int i = first; i < last; i++) {
			const QoiImgMeta &m = p->geom.metas[(size_t)i];
			CK(ctx, cudaMemcpyAsync(p->d_streams + m.stream_off, (const unsigned char *)h_streams + m.stream_off, (size_t)h_lens[i],
			                        cudaMemcpyHostToDevice, ctx->s_h2d));
			const uint64_t e = m.dec_off + (uint64_t)m.npx * p->geom.dec_channels;
			if (m.dec_off < lo) lo = m.dec_off;
			if (e > hi) hi = e;
		}
		CK(ctx, cudaEventRecord(p->ev_in[c], ctx->s_h2d));
		CK(ctx, cudaStreamWaitEvent(ctx->stream, p->ev_in[c], 0));
		if (whole) {
			if (c + 1 < nc) continue; /* all streams first, then the single launch */
			rc = qoi_gpu_decode(p, p->d_streams, p->d_lens, p->d_decoded, 1, nullptr, ctx->stream);
			if (rc) return rc;
			CK(ctx, cudaMemcpyAsync(h_decoded, p->d_decoded, p->geom.decoded_bytes, cudaMemcpyDeviceToHost, ctx->stream));
			CK(ctx, cudaStreamSynchronize(ctx->stream));
			return QOI_GPU_OK;
		}
		rc = qoi_gpu_decode(p->chunks[c], p->d_streams, p->d_lens + first, p->d_decoded, use_index, nullptr, ctx->stream);
		if (rc) return rc;
		CK(ctx, cudaEventRecord(p->ev_k[c], ctx->stream));
		CK(ctx, cudaStreamWaitEvent(ctx->s_d2h, p->ev_k[c], 0));
		CK(ctx, cudaMemcpyAsync((unsigned char *)h_decoded + lo, p->d_decoded + lo, (size_t)(hi - lo), cudaMemcpyDeviceToHost, ctx->s_d2h));
	}
	CK(ctx, cudaStreamSynchronize(ctx->s_d2h));
	return QOI_GPU_OK;
}

void *qoi_gpu_host_alloc(size_t bytes) {
	void *p = nullptr;
	if (cudaMallocHost(&p, bytes ? bytes : 1) != cudaSuccess) return nullptr;
	return p;
}
void qoi_gpu_host_free(void *p) {
	if (p) cudaFreeHost(p);
}

/* ------------------------------------------------------------------ staged copies for pageable caller memory */
/*
 * qoi_encode / qoi_decode take and return ordinary malloc'ed memory (qoi.h:276,287; callers free() the result).  A plain
 * cudaMemcpy from pageable memory is staged by the driver with one host thread at 5-10 GB/s -- 22 ms for a 4096^2 RGBA
 * image and its stream, against ~2 ms of PCIe time.  Here a few persistent host threads copy 4 MiB chunks between the
 * caller's buffer and a pinned staging buffer while the DMA of the chunks already copied is in flight; the result buffer's
 * first-touch page faults are spread over the same threads.  QOI_GPU_HOST_THREADS (1..64) sets their number (default 16).
 * QOI_GPU_TRACE=1 prints the stage times of every qoi_encode / qoi_decode call to stderr.
 */
class HostPool {
public:
	static HostPool &get() {
		static HostPool pool;
		return pool;
	}
	/* fn(i) for every i in [0, n); the calling thread works too; returns when all are done */
	void run(int n, const std::function<void(int)> &fn) {
		if (n <= 0) return;
		if (workers_.empty() || n == 1) {
			for (int i = 0; i < n; i++) fn(i);
			return;
		}
		std::unique_lock<std::mutex> lk(mu_);
		fn_ = &fn;
		n_ = n;
		next_.store(0);
		left_ = n;
		gen_++;
		cv_.notify_all();
		lk.unlock();
		work();
		lk.lock();
		done_cv_.wait(lk, [&] { return left_ == 0; });
		fn_ = nullptr;
	}

private:
	HostPool() {
		int want = 16;
		const char *e = getenv("QOI_GPU_HOST_THREADS");
		if (e && atoi(e) >= 1 && atoi(e) <= 64) want = atoi(e);
		const unsigned hw = std::thread::hardware_concurrency();
		if (hw && (unsigned)want > hw) want = (int)hw;
		for (int i = 1; i < want; i++) workers_.emplace_back([this] { loop(); });
	}
	~HostPool() {
		{
			std::lock_guard<std::mutex> lk(mu_);
			stop_ = true;
			gen_++;
		}
		cv_.notify_all();
		for (std::thread &t : workers_) t.join();
	}
	void work() {
		while (true) {
			const int i = next_.fetch_add(1);
			if (i >= n_) break;
			(*fn_)(i);
			std::lock_guard<std::mutex> lk(mu_);
			if (--left_ == 0) done_cv_.notify_all();
		}
	}
	void loop() {
		uint64_t seen = 0;
		while (true) {
			{
				std::unique_lock<std::mutex> lk(mu_);
				cv_.wait(lk, [&] { return gen_ != seen; });
				seen = gen_;
				if (stop_) return;
			}
			work();
		}
	}
	std::vector<std::thread> workers_;
	std::mutex mu_;
	std::condition_variable cv_, done_cv_;
	const std::function<void(int)> *fn_ = nullptr;
	int n_ = 0, left_ = 0;
	std::atomic<int> next_{0};
	uint64_t gen_ = 0;
	bool stop_ = false;
};


/* stage timing of the reference entry points (QOI_GPU_TRACE=1): where a single-image call spends its time */
struct Trace {
	bool on;
	std::chrono::steady_clock::time_point t0, last;
	char line[512];
	int n = 0;
	explicit Trace(const char *what) {
		static const bool enabled = getenv("QOI_GPU_TRACE") != nullptr;
		on = enabled;
		if (on) {
			t0 = last = std::chrono::steady_clock::now();
			n = snprintf(line, sizeof line, "qoi_gpu trace %s:", what);
		}
	}
	void mark(const char *stage) {
		if (!on) return;
		const auto now = std::chrono::steady_clock::now();
		n += snprintf(line + n, sizeof line - (size_t)n, " %s %.3f", stage, std::chrono::duration<double, std::milli>(now - last).count());
		if (n > (int)sizeof line - 64) n = (int)sizeof line - 64;
		last = now;
	}
	~Trace() {
		if (on) fprintf(stderr, "%s | total %.3f ms\n", line, std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count());
	}
};

/* First touch of a freshly malloc'ed result buffer: the kernel zeroes every page on its first write (a 64 MiB image: 16 k
   faults, ~8 ms on one thread).  The pool touches the pages in parallel while the DMA and the kernels of the same call run,
   so that the copy-out afterwards lands in mapped pages.  Only whole pages inside [p, p + bytes) are written (one zero byte
   each; the buffer's content is undefined at this point anyway). */
static void pretouch(void *p, size_t bytes) {
	const size_t page = 4096;
	const uintptr_t lo = ((uintptr_t)p + page - 1) & ~(uintptr_t)(page - 1), hi = ((uintptr_t)p + bytes) & ~(uintptr_t)(page - 1);
	if (hi <= lo) return;
	const size_t npages = (hi - lo) / page;
	const size_t per = 512; /* 2 MiB per task */
	const int tasks = (int)((npages + per - 1) / per);
	HostPool::get().run(tasks, [&](int i) {
		volatile unsigned char *q = (volatile unsigned char *)(lo + (size_t)i * per * page);
		const size_t n = (size_t)i * per + per <= npages ? per : npages - (size_t)i * per;
		for (size_t k = 0; k < n; k++) q[k * page] = 0;
	});
}

#define QOI_STAGE_CHUNK ((size_t)4 << 20)
#define QOI_STAGE_MAX_CHUNKS 64

static int stage_reserve(qoi_gpu_ctx *ctx, size_t bytes) {
	if (ctx->h_stage_bytes >= bytes) return QOI_GPU_OK;
	if (ctx->h_stage) cudaFreeHost(ctx->h_stage);
	ctx->h_stage = nullptr;
	ctx->h_stage_bytes = 0;
	const size_t want = (bytes + ((size_t)1 << 20)) & ~(((size_t)1 << 20) - 1);
	CK(ctx, cudaMallocHost(&ctx->h_stage, want));
	ctx->h_stage_bytes = want;
	return QOI_GPU_OK;
}

static size_t stage_chunk_bytes(size_t bytes) {
	size_t c = QOI_STAGE_CHUNK;
	while ((bytes + c - 1) / c > QOI_STAGE_MAX_CHUNKS) c *= 2;
	return c;
}

/* pageable host -> device, through ctx->h_stage + stage_off; the DMAs are queued on s (the workers enqueue their own chunk) */
static int staged_h2d(qoi_gpu_ctx *ctx, void *d_dst, const void *h_src, size_t bytes, size_t stage_off, cudaStream_t s) {
	if (!bytes) return QOI_GPU_OK;
	const size_t c = stage_chunk_bytes(bytes);
	const int n = (int)((bytes + c - 1) / c);
	std::atomic<int> failed{0};
	const int dev = ctx->device;
	HostPool::get().run(n, [&](int i) {
		const size_t off = (size_t)i * c, len = bytes - off < c ? bytes - off : c;
		memcpy(ctx->h_stage + stage_off + off, (const unsigned char *)h_src + off, len);
		cudaSetDevice(dev); /* the current device is per thread */
		if (cudaMemcpyAsync((unsigned char *)d_dst + off, ctx->h_stage + stage_off + off, len, cudaMemcpyHostToDevice, s) != cudaSuccess) failed = 1;
	});
	return failed ? QOI_GPU_E_CUDA : QOI_GPU_OK;
}

/* device -> pageable host: chunked DMAs into the staging buffer, each followed by an event; the workers wait for their
   chunk's event and copy it out.  Synchronous: returns when h_dst is complete. */
static int staged_d2h(qoi_gpu_ctx *ctx, void *h_dst, const void *d_src, size_t bytes, size_t stage_off, cudaStream_t s) {
	if (!bytes) return QOI_GPU_OK;
	const size_t c = stage_chunk_bytes(bytes);
	const int n = (int)((bytes + c - 1) / c);
	while ((int)ctx->stage_events.size() < n) {
		cudaEvent_t e;
		CK(ctx, cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
		ctx->stage_events.push_back(e);
	}
	for (int i = 0; i < n; i++) {
		const size_t off = (size_t)i * c, len = bytes - off < c ? bytes - off : c;
		CK(ctx, cudaMemcpyAsync(ctx->h_stage + stage_off + off, (const unsigned char *)d_src + off, len, cudaMemcpyDeviceToHost, s));
		CK(ctx, cudaEventRecord(ctx->stage_events[(size_t)i], s));
	}
	std::atomic<int> failed{0};
	HostPool::get().run(n, [&](int i) {
		const size_t off = (size_t)i * c, len = bytes - off < c ? bytes - off : c;
		if (cudaEventSynchronize(ctx->stage_events[(size_t)i]) != cudaSuccess) failed = 1;
		memcpy((unsigned char *)h_dst + off, ctx->h_stage + stage_off + off, len);
	});
	return failed ? QOI_GPU_E_CUDA : QOI_GPU_OK;
}

/* ------------------------------------------------------------------ the reference's own entry points */
static std::mutex g_mu;
static qoi_gpu_ctx *g_ctx = nullptr;
static bool g_ctx_failed = false;
static qoi_gpu_plan *g_plan = nullptr;
static qoi_desc g_plan_desc;
static int g_plan_dec_ch = 0;
static size_t g_last_len = 0; /* bytes of the last stream qoi_encode produced with g_plan */

static qoi_gpu_ctx *global_ctx() {
	if (!g_ctx && !g_ctx_failed) {
		int dev = 0;
		const char *e = getenv("QOI_GPU_DEVICE");
		if (e) dev = atoi(e);
		if (qoi_gpu_ctx_create(&g_ctx, dev) != QOI_GPU_OK) {
			g_ctx_failed = true;
			fprintf(stderr, "qoi_gpu: no usable CUDA device -- qoi_encode/qoi_decode have no CPU path and will return NULL\n");
		}
	}
	return g_ctx;
}

static qoi_gpu_plan *global_plan(qoi_gpu_ctx *ctx, const qoi_desc *d, int dec_ch) {
	if (g_plan && memcmp(&g_plan_desc, d, sizeof *d) == 0 && g_plan_dec_ch == dec_ch) return g_plan;
	qoi_gpu_plan_destroy(g_plan);
	g_plan = nullptr;
	const uint64_t zero = 0;
	if (qoi_gpu_plan_create(ctx, &g_plan, 1, d, &zero, &zero, &zero, dec_ch) != QOI_GPU_OK) return nullptr;
	memset(&g_plan_desc, 0, sizeof g_plan_desc);
	g_last_len = 0;
	g_plan_desc = *d;
	g_plan_dec_ch = dec_ch;
	return g_plan;
}

/* reference qoi.h:356-486 */
void *qoi_encode(const void *data, const qoi_desc *desc, int *out_len) {
	if (!data || !out_len || !desc) return NULL;
	QoiDescPOD pod = {desc->width, desc->height, desc->channels, desc->colorspace};
	if (!qoi_geometry_ok(pod)) return NULL; /* qoi.h:364-372 */
	qoi_desc d;
	memset(&d, 0, sizeof d);
	d.width = desc->width; d.height = desc->height; d.channels = desc->channels; d.colorspace = desc->colorspace;
	std::lock_guard<std::mutex> lock(g_mu);
	qoi_gpu_ctx *ctx = global_ctx();
	if (!ctx) return NULL;
	DeviceGuard guard(ctx->device);
	qoi_gpu_plan *p = global_plan(ctx, &d, d.channels);
	if (!p || plan_mirrors(p)) return NULL;
	cudaStream_t s = ctx->stream;
	Trace tr("qoi_encode");
	const size_t nbytes = (size_t)d.width * d.height * d.channels;
	const size_t max_size = qoi_max_stream(pod);
	if (stage_reserve(ctx, max_size > nbytes ? max_size : nbytes)) return NULL;
	/* the reference allocates the worst case as well (qoi.h:374-379); untouched pages cost nothing */
	void *out = malloc(max_size);
	if (!out) return NULL;
	tr.mark("setup");
	bool ok = staged_h2d(ctx, p->d_pixels, data, nbytes, 0, s) == QOI_GPU_OK;
	tr.mark("h2d_copy");
	ok = ok && qoi_gpu_encode(p, p->d_pixels, p->d_streams, p->d_lens, s) == QOI_GPU_OK &&
	     cudaMemcpyAsync(p->h_lens_pinned, p->d_lens, sizeof(int32_t), cudaMemcpyDeviceToHost, s) == cudaSuccess;
	/* while the DMA and the kernel run: map the pages the stream will most likely need (as long as the last one of this plan) */
	if (ok) pretouch(out, g_last_len && g_last_len < max_size ? g_last_len : nbytes / 2 + nbytes / 4);
	tr.mark("pretouch");
	ok = ok && cudaStreamSynchronize(s) == cudaSuccess;
	tr.mark("sync");
	const int len = ok ? p->h_lens_pinned[0] : 0;
	ok = ok && staged_d2h(ctx, out, p->d_streams, (size_t)len, 0, s) == QOI_GPU_OK; /* the staging buffer is free again: the kernel has consumed the pixels */
	tr.mark("d2h_copy");
	if (!ok) {
		free(out);
		return NULL;
	}
	g_last_len = (size_t)len;
	*out_len = len; /* written only on success, qoi.h:484 */
	return out;
}

/* reference qoi.h:488-590 */
void *qoi_decode(const void *data, int size, qoi_desc *desc, int channels) {
	if (!data || !desc || (channels != 0 && channels != 3 && channels != 4) || size < QOI_HEADER_BYTES + QOI_END_BYTES) return NULL; /* qoi.h:497-503 */
	const unsigned char *b = (const unsigned char *)data;
	const unsigned magic = ((unsigned)b[0] << 24) | ((unsigned)b[1] << 16) | ((unsigned)b[2] << 8) | b[3];
	desc->width = ((unsigned)b[4] << 24) | ((unsigned)b[5] << 16) | ((unsigned)b[6] << 8) | b[7]; /* filled before validation, qoi.h:508-511 */
	desc->height = ((unsigned)b[8] << 24) | ((unsigned)b[9] << 16) | ((unsigned)b[10] << 8) | b[11];
	desc->channels = b[12];
	desc->colorspace = b[13];
	QoiDescPOD pod = {desc->width, desc->height, desc->channels, desc->colorspace};
	if (magic != 0x716F6966u || !qoi_geometry_ok(pod)) return NULL; /* qoi.h:513-521 */
	const int out_ch = channels ? channels : desc->channels;
	qoi_desc d;
	memset(&d, 0, sizeof d);
	d.width = desc->width; d.height = desc->height; d.channels = desc->channels; d.colorspace = desc->colorspace;
	std::lock_guard<std::mutex> lock(g_mu);
	qoi_gpu_ctx *ctx = global_ctx();
	if (!ctx) return NULL;
	DeviceGuard guard(ctx->device);
	qoi_gpu_plan *p = global_plan(ctx, &d, out_ch);
	if (!p || plan_mirrors(p)) return NULL;
	cudaStream_t s = ctx->stream;
	/* a stream longer than the worst case only carries bytes no op can consume before the image is full */
	size_t up = (size_t)size;
	const size_t cap = qoi_max_foreign_stream(pod);
	int eff_size = size;
	if (up > cap) { up = cap; eff_size = (int)cap; }
	const size_t out_bytes = (size_t)d.width * d.height * (size_t)out_ch;
	if (stage_reserve(ctx, out_bytes > up ? out_bytes : up)) return NULL;
	Trace tr("qoi_decode");
	void *out = malloc(out_bytes);
	if (!out) return NULL;
	p->h_lens_pinned[0] = eff_size;
	tr.mark("setup");
	bool ok = staged_h2d(ctx, p->d_streams, data, up, 0, s) == QOI_GPU_OK;
	tr.mark("h2d_copy");
	ok = ok && cudaMemcpyAsync(p->d_lens, p->h_lens_pinned, sizeof(int32_t), cudaMemcpyHostToDevice, s) == cudaSuccess &&
	     qoi_gpu_decode(p, p->d_streams, p->d_lens, p->d_decoded, 0, nullptr, s) == QOI_GPU_OK;
	tr.mark("launch");
	if (ok) pretouch(out, out_bytes); /* while the DMA and the kernels run */
	tr.mark("pretouch");
	ok = ok && staged_d2h(ctx, out, p->d_decoded, out_bytes, 0, s) == QOI_GPU_OK; /* stream-ordered behind the kernels, which have read the staged stream */
	tr.mark("d2h_copy");
	if (!ok) {
		free(out);
		return NULL;
	}
	return out;
}

/* the two stdio conveniences, so that the library alone offers all four reference entry points
   (a program that defines QOI_IMPLEMENTATION gets its own copies from include/qoi.h; reference qoi.h:595-646) */
int qoi_write(const char *filename, const void *data, const qoi_desc *desc) {
	FILE *fp = fopen(filename, "wb");
	if (!fp) return 0;
	int n = 0;
	void *stream = qoi_encode(data, desc, &n);
	if (!stream) {
		fclose(fp);
		return 0;
	}
	fwrite(stream, 1, (size_t)n, fp);
	fflush(fp);
	const int bad = ferror(fp);
	fclose(fp);
	free(stream);
	return bad ? 0 : n;
}

void *qoi_read(const char *filename, qoi_desc *desc, int channels) {
	FILE *fp = fopen(filename, "rb");
	if (!fp) return NULL;
	long len;
	if (fseek(fp, 0, SEEK_END) != 0 || (len = ftell(fp)) <= 0 || fseek(fp, 0, SEEK_SET) != 0) {
		fclose(fp);
		return NULL;
	}
	void *stream = malloc((size_t)len);
	if (!stream) {
		fclose(fp);
		return NULL;
	}
	const size_t got = fread(stream, 1, (size_t)len, fp);
	fclose(fp);
	void *px = (got == (size_t)len) ? qoi_decode(stream, (int)len, desc, channels) : NULL;
	free(stream);
	return px;
}

} /* extern "C" */
