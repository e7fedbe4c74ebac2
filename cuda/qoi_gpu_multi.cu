/*
 * qoi_gpu_multi.cu -- C ABI of the multi-device layer (libqoi_gpu_nccl.so): include/qoi_gpu_multi.h.
 *
 * One process drives N devices.  The codec itself is libqoi_gpu.so, used through its public C API (one context and one
 * plan per device); this file only partitions the batch and moves data between devices with NCCL: grouped
 * ncclSend / ncclRecv for the pixel scatter and the stream / pixel gathers (NCCL has no scatterv / gatherv), ncclAllGather
 * for the stream lengths.  No codec arithmetic here, no CPU path.
 */
#include <cuda_runtime.h>
#include <nccl.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <new>
#include <vector>

#include "../include/qoi_gpu_multi.h"

struct Rank {
	int device = 0;
	qoi_gpu_ctx *ctx = nullptr;
	cudaStream_t stream = nullptr;
	ncclComm_t comm = nullptr;
	cudaEvent_t ev[4] = {nullptr, nullptr, nullptr, nullptr};
};

struct qoi_gpu_mctx {
	std::vector<Rank> ranks;
	char err[256] = {0};
};

struct Shard {
	int first = 0, count = 0;
	qoi_gpu_plan *plan = nullptr;
	std::vector<uint64_t> pixel_off, stream_off, decoded_off;
	uint64_t pixel_bytes = 0, stream_bytes = 0, decoded_bytes = 0;
	unsigned char *d_pixels = nullptr, *d_streams = nullptr, *d_decoded = nullptr;
	int32_t *d_lens = nullptr;     /* padded to max_count */
	int32_t *d_all_lens = nullptr; /* [G][max_count] */
};

struct qoi_gpu_mplan {
	qoi_gpu_mctx *m = nullptr;
	std::vector<qoi_desc> descs;
	std::vector<Shard> shards;
	std::vector<uint64_t> root_pixel_off, root_decoded_off; /* packed batch on the root */
	int channels = 0, dec_channels = 0, max_count = 0;
	int32_t *h_all_lens = nullptr; /* pinned [G][max_count] */
};

#define MCK(m, call)                                                                              \
	do {                                                                                          \
		cudaError_t e_ = (call);                                                                  \
		if (e_ != cudaSuccess) {                                                                  \
			snprintf((m)->err, sizeof((m)->err), "%s: %s", #call, cudaGetErrorString(e_));        \
			return QOI_GPU_E_CUDA;                                                                \
		}                                                                                         \
	} while (0)
#define NCK(m, call)                                                                              \
	do {                                                                                          \
		ncclResult_t r_ = (call);                                                                 \
		if (r_ != ncclSuccess) {                                                                  \
			snprintf((m)->err, sizeof((m)->err), "%s: %s", #call, ncclGetErrorString(r_));        \
			return QOI_GPU_E_CUDA;                                                                \
		}                                                                                         \
	} while (0)

extern "C" {

int qoi_gpu_mctx_create(qoi_gpu_mctx **out, const int *devices, int n) {
	if (!out || !devices || n < 1) return QOI_GPU_E_ARG;
	*out = nullptr;
	if (n > qoi_gpu_device_count()) return QOI_GPU_E_CUDA;
	qoi_gpu_mctx *m = new (std::nothrow) qoi_gpu_mctx;
	if (!m) return QOI_GPU_E_NOMEM;
	m->ranks.resize((size_t)n);
	int prev = 0;
	cudaGetDevice(&prev);
	int rc = QOI_GPU_OK;
	std::vector<ncclComm_t> comms((size_t)n);
	if (ncclCommInitAll(comms.data(), n, devices) != ncclSuccess) rc = QOI_GPU_E_CUDA;
	for (int r = 0; r < n && !rc; r++) {
		Rank &k = m->ranks[(size_t)r];
		k.device = devices[r];
		k.comm = comms[(size_t)r];
		if (cudaSetDevice(k.device) != cudaSuccess || cudaStreamCreateWithFlags(&k.stream, cudaStreamNonBlocking) != cudaSuccess) rc = QOI_GPU_E_CUDA;
		for (int e = 0; e < 4 && !rc; e++)
			if (cudaEventCreate(&k.ev[e]) != cudaSuccess) rc = QOI_GPU_E_CUDA;
		if (!rc) rc = qoi_gpu_ctx_create(&k.ctx, k.device);
	}
	cudaSetDevice(prev);
	if (rc) {
		qoi_gpu_mctx_destroy(m);
		return rc;
	}
	*out = m;
	return QOI_GPU_OK;
}

void qoi_gpu_mctx_destroy(qoi_gpu_mctx *m) {
	if (!m) return;
	int prev = 0;
	cudaGetDevice(&prev);
	for (Rank &k : m->ranks) {
		cudaSetDevice(k.device);
		if (k.ctx) qoi_gpu_ctx_destroy(k.ctx);
		for (cudaEvent_t e : k.ev)
			if (e) cudaEventDestroy(e);
		if (k.stream) cudaStreamDestroy(k.stream);
		if (k.comm) ncclCommDestroy(k.comm);
	}
	cudaSetDevice(prev);
	delete m;
}

int qoi_gpu_mctx_size(const qoi_gpu_mctx *m) { return m ? (int)m->ranks.size() : 0; }
const char *qoi_gpu_mctx_last_error(const qoi_gpu_mctx *m) { return m ? m->err : ""; }

int qoi_gpu_mplan_create(qoi_gpu_mctx *m, qoi_gpu_mplan **out, int n, const qoi_desc *descs, int decoded_channels) {
	if (!m || !out || n < 1 || !descs) return QOI_GPU_E_ARG;
	*out = nullptr;
	qoi_gpu_mplan *p = new (std::nothrow) qoi_gpu_mplan;
	if (!p) return QOI_GPU_E_NOMEM;
	p->m = m;
	p->descs.assign(descs, descs + n);
	p->channels = descs[0].channels;
	p->dec_channels = decoded_channels ? decoded_channels : p->channels;
	const int G = (int)m->ranks.size();
	const int per = (n + G - 1) / G; /* contiguous blocks: gathers stay contiguous */
	p->max_count = per;
	p->shards.resize((size_t)G);
	p->root_pixel_off.resize((size_t)n);
	p->root_decoded_off.resize((size_t)n);
	uint64_t po = 0, dof = 0;
	for (int i = 0; i < n; i++) {
		p->root_pixel_off[(size_t)i] = po;
		p->root_decoded_off[(size_t)i] = dof;
		po += (uint64_t)descs[i].width * descs[i].height * descs[i].channels;
		dof += (uint64_t)descs[i].width * descs[i].height * (uint64_t)p->dec_channels;
	}
	int prev = 0;
	cudaGetDevice(&prev);
	int rc = QOI_GPU_OK;
	for (int r = 0; r < G && !rc; r++) {
		Shard &s = p->shards[(size_t)r];
		s.first = r * per < n ? r * per : n;
		s.count = s.first + per <= n ? per : n - s.first;
		if (cudaSetDevice(m->ranks[(size_t)r].device) != cudaSuccess) { rc = QOI_GPU_E_CUDA; break; }
		cudaError_t e = cudaMalloc(&s.d_lens, sizeof(int32_t) * (size_t)per);
		if (e == cudaSuccess) e = cudaMemset(s.d_lens, 0, sizeof(int32_t) * (size_t)per);
		if (e == cudaSuccess) e = cudaMalloc(&s.d_all_lens, sizeof(int32_t) * (size_t)per * (size_t)G);
		if (e != cudaSuccess) { rc = QOI_GPU_E_NOMEM; break; }
		if (s.count == 0) continue; /* more devices than images: this rank only takes part in the collectives */
		s.pixel_off.resize((size_t)s.count);
		s.stream_off.resize((size_t)s.count);
		s.decoded_off.resize((size_t)s.count);
		uint64_t a = 0, b = 0, c = 0;
		for (int k = 0; k < s.count; k++) {
			const qoi_desc &d = descs[s.first + k];
			s.pixel_off[(size_t)k] = a;
			s.stream_off[(size_t)k] = b;
			s.decoded_off[(size_t)k] = c;
			a += (uint64_t)d.width * d.height * d.channels;
			b += ((uint64_t)d.width * d.height * 5u + 22u + 255u) & ~(uint64_t)255; /* a slot also holds any foreign stream of the image */
			c += (uint64_t)d.width * d.height * (uint64_t)p->dec_channels;
		}
		s.pixel_bytes = a;
		s.stream_bytes = b;
		s.decoded_bytes = c;
		rc = qoi_gpu_plan_create(m->ranks[(size_t)r].ctx, &s.plan, s.count, descs + s.first, s.pixel_off.data(), s.stream_off.data(), s.decoded_off.data(),
		                         decoded_channels);
		if (rc) break;
		e = cudaMalloc(&s.d_pixels, a + 16);
		if (e == cudaSuccess) e = cudaMalloc(&s.d_streams, b + 16);
		if (e == cudaSuccess) e = cudaMalloc(&s.d_decoded, c + 16);
		if (e != cudaSuccess) rc = QOI_GPU_E_NOMEM;
	}
	if (!rc && cudaMallocHost(&p->h_all_lens, sizeof(int32_t) * (size_t)per * (size_t)G) != cudaSuccess) rc = QOI_GPU_E_NOMEM;
	cudaSetDevice(prev);
	if (rc) {
		snprintf(m->err, sizeof m->err, "qoi_gpu_mplan_create: allocation or plan failed (rc %d)", rc);
		qoi_gpu_mplan_destroy(p);
		return rc;
	}
	*out = p;
	return QOI_GPU_OK;
}

void qoi_gpu_mplan_destroy(qoi_gpu_mplan *p) {
	if (!p) return;
	int prev = 0;
	cudaGetDevice(&prev);
	for (size_t r = 0; r < p->shards.size(); r++) {
		Shard &s = p->shards[r];
		cudaSetDevice(p->m->ranks[r].device);
		if (s.plan) qoi_gpu_plan_destroy(s.plan);
		cudaFree(s.d_pixels);
		cudaFree(s.d_streams);
		cudaFree(s.d_decoded);
		cudaFree(s.d_lens);
		cudaFree(s.d_all_lens);
	}
	if (p->h_all_lens) cudaFreeHost(p->h_all_lens);
	cudaSetDevice(prev);
	delete p;
}

int qoi_gpu_mplan_shard(const qoi_gpu_mplan *p, int rank, qoi_gpu_shard *out) {
	if (!p || !out || rank < 0 || rank >= (int)p->shards.size()) return QOI_GPU_E_ARG;
	const Shard &s = p->shards[(size_t)rank];
	out->device = p->m->ranks[(size_t)rank].device;
	out->first_image = s.first;
	out->n_images = s.count;
	out->d_pixels = s.d_pixels;
	out->d_streams = s.d_streams;
	out->d_decoded = s.d_decoded;
	out->d_lens = s.d_lens;
	out->pixel_off = s.pixel_off.data();
	out->stream_off = s.stream_off.data();
	out->decoded_off = s.decoded_off.data();
	out->pixel_bytes = s.pixel_bytes;
	out->stream_bytes = s.stream_bytes;
	out->decoded_bytes = s.decoded_bytes;
	return QOI_GPU_OK;
}

/* max over ranks of the time between events a and b (ms); the streams must be idle */
static float phase_ms(qoi_gpu_mctx *m, int a, int b) {
	float best = 0.f;
	for (Rank &k : m->ranks) {
		float t = 0.f;
		cudaSetDevice(k.device);
		if (cudaEventElapsedTime(&t, k.ev[a], k.ev[b]) == cudaSuccess && t > best) best = t;
	}
	return best;
}

static int sync_all(qoi_gpu_mctx *m) {
	for (Rank &k : m->ranks) {
		MCK(m, cudaSetDevice(k.device));
		MCK(m, cudaStreamSynchronize(k.stream));
	}
	return QOI_GPU_OK;
}

static int record_all(qoi_gpu_mctx *m, int e) {
	for (Rank &k : m->ranks) {
		MCK(m, cudaSetDevice(k.device));
		MCK(m, cudaEventRecord(k.ev[e], k.stream));
	}
	return QOI_GPU_OK;
}

int qoi_gpu_mencode(qoi_gpu_mplan *p, const void *d_root_pixels, void *d_root_streams, int32_t *h_lens, uint64_t *h_root_off, qoi_gpu_mtimes *times) {
	if (!p || !h_lens) return QOI_GPU_E_ARG;
	qoi_gpu_mctx *m = p->m;
	const int G = (int)m->ranks.size(), n = (int)p->descs.size();
	int prev = 0;
	cudaGetDevice(&prev);
	qoi_gpu_mtimes T = {0.f, 0.f, 0.f, 0.f};
	int rc = QOI_GPU_OK;
	do {
		/* ---- scatter: root -> ranks, one message per rank (blocks are contiguous in the packed root buffer) ---- */
		if ((rc = record_all(m, 0))) break;
		if (d_root_pixels) {
			NCK(m, ncclGroupStart());
			for (int r = 1; r < G; r++) {
				Shard &s = p->shards[(size_t)r];
				if (!s.count) continue;
				const unsigned char *src = (const unsigned char *)d_root_pixels + p->root_pixel_off[(size_t)s.first];
				NCK(m, ncclSend(src, (size_t)s.pixel_bytes, ncclUint8, r, m->ranks[0].comm, m->ranks[0].stream));
				NCK(m, ncclRecv(s.d_pixels, (size_t)s.pixel_bytes, ncclUint8, 0, m->ranks[(size_t)r].comm, m->ranks[(size_t)r].stream));
			}
			NCK(m, ncclGroupEnd());
			MCK(m, cudaSetDevice(m->ranks[0].device));
			if (p->shards[0].count)
				MCK(m, cudaMemcpyAsync(p->shards[0].d_pixels, d_root_pixels, (size_t)p->shards[0].pixel_bytes, cudaMemcpyDeviceToDevice, m->ranks[0].stream));
		}
		if ((rc = record_all(m, 1))) break;
		/* ---- encode: every rank its block, concurrently ---- */
		for (int r = 0; r < G && !rc; r++) {
			Shard &s = p->shards[(size_t)r];
			if (!s.count) continue;
			rc = qoi_gpu_encode(s.plan, s.d_pixels, s.d_streams, s.d_lens, m->ranks[(size_t)r].stream);
		}
		if (rc) break;
		if ((rc = record_all(m, 2))) break;
		/* ---- all-gather of the stream lengths (padded blocks of max_count int32) ---- */
		NCK(m, ncclGroupStart());
		for (int r = 0; r < G; r++)
			NCK(m, ncclAllGather(p->shards[(size_t)r].d_lens, p->shards[(size_t)r].d_all_lens, (size_t)p->max_count, ncclInt32, m->ranks[(size_t)r].comm,
			                     m->ranks[(size_t)r].stream));
		NCK(m, ncclGroupEnd());
		MCK(m, cudaSetDevice(m->ranks[0].device));
		MCK(m, cudaMemcpyAsync(p->h_all_lens, p->shards[0].d_all_lens, sizeof(int32_t) * (size_t)p->max_count * (size_t)G, cudaMemcpyDeviceToHost, m->ranks[0].stream));
		if ((rc = record_all(m, 3))) break;
		if ((rc = sync_all(m))) break;
		T.scatter_ms = d_root_pixels ? phase_ms(m, 0, 1) : 0.f;
		T.kernel_ms = phase_ms(m, 1, 2);
		T.lens_ms = phase_ms(m, 2, 3);
		uint64_t off = 0;
		for (int i = 0; i < n; i++) {
			const int r = i / p->max_count, k = i - r * p->max_count;
			h_lens[i] = p->h_all_lens[(size_t)r * (size_t)p->max_count + (size_t)k];
			if (h_root_off) h_root_off[i] = off;
			off += (uint64_t)h_lens[i];
		}
		/* ---- gather: ranks -> root, every stream to the exclusive scan of the lengths ---- */
		if (d_root_streams) {
			if ((rc = record_all(m, 0))) break;
			NCK(m, ncclGroupStart());
			uint64_t o = 0;
			for (int i = 0; i < n; i++) {
				const int r = i / p->max_count, k = i - r * p->max_count;
				const Shard &s = p->shards[(size_t)r];
				unsigned char *dst = (unsigned char *)d_root_streams + o;
				if (r != 0) {
					NCK(m, ncclSend(s.d_streams + s.stream_off[(size_t)k], (size_t)h_lens[i], ncclUint8, 0, m->ranks[(size_t)r].comm, m->ranks[(size_t)r].stream));
					NCK(m, ncclRecv(dst, (size_t)h_lens[i], ncclUint8, r, m->ranks[0].comm, m->ranks[0].stream));
				}
				o += (uint64_t)h_lens[i];
			}
			NCK(m, ncclGroupEnd());
			MCK(m, cudaSetDevice(m->ranks[0].device));
			o = 0;
			for (int i = 0; i < n && i < p->shards[0].count; i++) { /* the root's own block: device-to-device copies */
				MCK(m, cudaMemcpyAsync((unsigned char *)d_root_streams + o, p->shards[0].d_streams + p->shards[0].stream_off[(size_t)i], (size_t)h_lens[i],
				                       cudaMemcpyDeviceToDevice, m->ranks[0].stream));
				o += (uint64_t)h_lens[i];
			}
			if ((rc = record_all(m, 1))) break;
			if ((rc = sync_all(m))) break;
			T.gather_ms = phase_ms(m, 0, 1);
		}
	} while (0);
	cudaSetDevice(prev);
	if (times) *times = T;
	return rc;
}

int qoi_gpu_mdecode(qoi_gpu_mplan *p, int use_index, void *d_root_decoded, qoi_gpu_mtimes *times) {
	if (!p) return QOI_GPU_E_ARG;
	qoi_gpu_mctx *m = p->m;
	const int G = (int)m->ranks.size();
	int prev = 0;
	cudaGetDevice(&prev);
	qoi_gpu_mtimes T = {0.f, 0.f, 0.f, 0.f};
	int rc = QOI_GPU_OK;
	do {
		if ((rc = record_all(m, 1))) break;
		for (int r = 0; r < G && !rc; r++) {
			Shard &s = p->shards[(size_t)r];
			if (!s.count) continue;
			rc = qoi_gpu_decode(s.plan, s.d_streams, s.d_lens, s.d_decoded, use_index, nullptr, m->ranks[(size_t)r].stream);
		}
		if (rc) break;
		if ((rc = record_all(m, 2))) break;
		if (d_root_decoded) { /* blocks are contiguous on both sides: one message per rank */
			NCK(m, ncclGroupStart());
			for (int r = 1; r < G; r++) {
				Shard &s = p->shards[(size_t)r];
				if (!s.count) continue;
				NCK(m, ncclSend(s.d_decoded, (size_t)s.decoded_bytes, ncclUint8, 0, m->ranks[(size_t)r].comm, m->ranks[(size_t)r].stream));
				NCK(m, ncclRecv((unsigned char *)d_root_decoded + p->root_decoded_off[(size_t)s.first], (size_t)s.decoded_bytes, ncclUint8, r, m->ranks[0].comm,
				                m->ranks[0].stream));
			}
			NCK(m, ncclGroupEnd());
			MCK(m, cudaSetDevice(m->ranks[0].device));
			if (p->shards[0].count)
				MCK(m, cudaMemcpyAsync(d_root_decoded, p->shards[0].d_decoded, (size_t)p->shards[0].decoded_bytes, cudaMemcpyDeviceToDevice, m->ranks[0].stream));
		}
		if ((rc = record_all(m, 3))) break;
		if ((rc = sync_all(m))) break;
		T.kernel_ms = phase_ms(m, 1, 2);
		T.gather_ms = d_root_decoded ? phase_ms(m, 2, 3) : 0.f;
	} while (0);
	cudaSetDevice(prev);
	if (times) *times = T;
	return rc;
}

} /* extern "C" */
