/*
 * qoi_gpu_cpu_double.c -- TEST DOUBLE, never shipped, never linked into the product.
 *
 * dropin/qoiconv_dir.c is host code around qoi_gpu_encode_host / qoi_gpu_decode_host.  This container has no
 * GPU, so its batching, file handling and threading are exercised here by LD_PRELOADing this object in front
 * of libqoi_gpu.so: the handful of include/qoi_gpu.h entry points the tool uses are re-implemented on the CPU
 * checker (oracle/liboracle.so: ora_encode / ora_decode, the restatement of qoi.h).  Only tests/ builds and loads
 * it (tests/test_zz_dropin_tools.py); on the GPU box the same tool runs against the real library.
 */
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "qoi_gpu.h"

void *ora_encode(const void *data, const qoi_desc *desc, int *out_len);
void *ora_decode(const void *data, int size, qoi_desc *desc, int channels);

struct qoi_gpu_ctx { int dummy; };
struct qoi_gpu_plan {
	int n, dec_channels;
	qoi_desc *descs;
	uint64_t *pix, *str, *dec;
};

int qoi_gpu_device_count(void) { return 1; }
int qoi_gpu_ctx_create(qoi_gpu_ctx **ctx, int device) {
	(void)device;
	*ctx = (qoi_gpu_ctx *)calloc(1, sizeof(qoi_gpu_ctx));
	return *ctx ? QOI_GPU_OK : QOI_GPU_E_NOMEM;
}
void qoi_gpu_ctx_destroy(qoi_gpu_ctx *ctx) { free(ctx); }
const char *qoi_gpu_last_error(const qoi_gpu_ctx *ctx) { (void)ctx; return "cpu test double"; }
size_t qoi_gpu_max_stream_size(const qoi_desc *d) { return (size_t)d->width * d->height * (d->channels + 1u) + 22u; }

int qoi_gpu_plan_create(qoi_gpu_ctx *ctx, qoi_gpu_plan **plan, int n, const qoi_desc *descs, const uint64_t *pixel_off,
                        const uint64_t *stream_off, const uint64_t *decoded_off, int decoded_channels) {
	(void)ctx;
	if (n <= 0 || !descs || !pixel_off || !stream_off || (decoded_channels != 3 && decoded_channels != 4)) return QOI_GPU_E_ARG;
	qoi_gpu_plan *p = (qoi_gpu_plan *)calloc(1, sizeof *p);
	p->n = n;
	p->dec_channels = decoded_channels;
	p->descs = (qoi_desc *)malloc(sizeof(qoi_desc) * (size_t)n);
	p->pix = (uint64_t *)malloc(8u * (size_t)n);
	p->str = (uint64_t *)malloc(8u * (size_t)n);
	p->dec = (uint64_t *)malloc(8u * (size_t)n);
	uint64_t run = 0;
	for (int i = 0; i < n; i++) {
		if (descs[i].channels != descs[0].channels) return QOI_GPU_E_ARG; /* one channel count per plan, like the real library */
		p->descs[i] = descs[i];
		p->pix[i] = pixel_off[i];
		p->str[i] = stream_off[i];
		p->dec[i] = decoded_off ? decoded_off[i] : run;
		run += (uint64_t)descs[i].width * descs[i].height * (uint64_t)decoded_channels;
	}
	*plan = p;
	return QOI_GPU_OK;
}
void qoi_gpu_plan_destroy(qoi_gpu_plan *p) {
	if (!p) return;
	free(p->descs); free(p->pix); free(p->str); free(p->dec); free(p);
}

int qoi_gpu_encode_host(qoi_gpu_plan *p, const void *h_pixels, void *h_streams, int32_t *h_lens) {
	for (int i = 0; i < p->n; i++) {
		int len = 0;
		void *e = ora_encode((const unsigned char *)h_pixels + p->pix[i], &p->descs[i], &len);
		if (!e) return QOI_GPU_E_ARG;
		memcpy((unsigned char *)h_streams + p->str[i], e, (size_t)len);
		h_lens[i] = len;
		free(e);
	}
	return QOI_GPU_OK;
}

int qoi_gpu_decode_host(qoi_gpu_plan *p, const void *h_streams, const int32_t *h_lens, void *h_decoded, int use_index) {
	(void)use_index;
	for (int i = 0; i < p->n; i++) {
		qoi_desc d;
		void *px = ora_decode((const unsigned char *)h_streams + p->str[i], h_lens[i], &d, p->dec_channels);
		/* the device compares the stream header with the plan (cuda/qoi_fdec.cuh stream_ok) */
		if (!px || d.width != p->descs[i].width || d.height != p->descs[i].height || d.channels != p->descs[i].channels ||
		    d.colorspace != p->descs[i].colorspace) {
			free(px);
			return QOI_GPU_E_STREAM;
		}
		memcpy((unsigned char *)h_decoded + p->dec[i], px, (size_t)d.width * d.height * (size_t)p->dec_channels);
		free(px);
	}
	return QOI_GPU_OK;
}

void *qoi_gpu_host_alloc(size_t bytes) { return malloc(bytes ? bytes : 1); }
void qoi_gpu_host_free(void *p) { free(p); }
