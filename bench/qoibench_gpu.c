/*
 * qoibench_gpu.c -- C host harness for the batch API (include/qoi_gpu.h), in the spirit of the
 * reference's qoibench.c: for a list of synthetic configurations it encodes and decodes a batch on
 * one or more GPUs, verifies the round trip (qoibench.c:410-417), and prints the reference's table
 * columns (qoibench.c:340-357: decode ms, encode ms, decode mpps, encode mpps, size kb, rate) plus
 * GB/s and the fraction of an HBM peak.  Timing protocol of BENCHMARK_FN (qoibench.c:364-376): one
 * discarded warm-up run, mean of N runs, device-resident buffers, cudaEvents on the launching stream.
 *
 * Host code is plain C99 + the CUDA runtime C API; one context, plan and pthread per GPU; images are
 * sharded by contiguous blocks, nothing crosses GPUs.  No CPU codec is involved (the CPU baseline
 * lives in bench.py).   usage: qoibench_gpu <runs> [--gpus N] [--batch B] [--peak GBps] [--sweep]
 * --sweep: BASELINE.json configs[4], sizes 64x64 .. 16384x16384, channels 3 and 4, batch sized to ~0.5 Gpx per GPU.
 */
#include <cuda_runtime_api.h>
#include <pthread.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "qoi_gpu.h"

void qoi_synth_batch(uint8_t *dst, size_t stride, int n, uint32_t w, uint32_t h, int channels, int dist, uint32_t first_index);

typedef struct {
	int w, h, ch, dist, batch;
	const char *name;
} config_t;

typedef struct {
	int device, rank, ranks, runs;
	config_t cfg;
	double enc_ms, dec_ms; /* out */
	uint64_t enc_bytes;    /* out */
	int ok;                /* out */
} job_t;

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { fprintf(stderr, "%s: %s\n", #x, cudaGetErrorString(e_)); j->ok = 0; return NULL; } } while (0)

static void *run_job(void *arg) {
	job_t *j = (job_t *)arg;
	const config_t c = j->cfg;
	const int per = (c.batch + j->ranks - 1) / j->ranks;
	const int lo = j->rank * per < c.batch ? j->rank * per : c.batch;
	const int n = (lo + per <= c.batch ? per : c.batch - lo);
	j->ok = 1;
	j->enc_ms = j->dec_ms = 0;
	j->enc_bytes = 0;
	if (n <= 0) return NULL;
	CK(cudaSetDevice(j->device));
	qoi_gpu_ctx *ctx = NULL;
	if (qoi_gpu_ctx_create(&ctx, j->device) != QOI_GPU_OK) { j->ok = 0; return NULL; }
	qoi_desc *descs = (qoi_desc *)calloc((size_t)n, sizeof(qoi_desc));
	uint64_t *poff = (uint64_t *)calloc((size_t)n, 8), *soff = (uint64_t *)calloc((size_t)n, 8);
	const size_t img = (size_t)c.w * c.h * c.ch, dimg = (size_t)c.w * c.h * 4;
	size_t slot = 0;
	for (int i = 0; i < n; i++) {
		descs[i].width = (unsigned)c.w; descs[i].height = (unsigned)c.h; descs[i].channels = (unsigned char)c.ch; descs[i].colorspace = QOI_SRGB;
		slot = (qoi_gpu_max_stream_size(&descs[i]) + 255) & ~(size_t)255;
		poff[i] = (uint64_t)i * img;
		soff[i] = (uint64_t)i * slot;
	}
	qoi_gpu_plan *plan = NULL;
	if (qoi_gpu_plan_create(ctx, &plan, n, descs, poff, soff, NULL, 4) != QOI_GPU_OK) { j->ok = 0; return NULL; }
	uint8_t *h_pix = (uint8_t *)malloc(img * n);
	qoi_synth_batch(h_pix, img, n, (uint32_t)c.w, (uint32_t)c.h, c.ch, c.dist, (uint32_t)lo);
	uint8_t *d_pix, *d_str, *d_dec;
	int32_t *d_len;
	CK(cudaMalloc((void **)&d_pix, img * n + 64));
	CK(cudaMalloc((void **)&d_str, slot * n + 64));
	CK(cudaMalloc((void **)&d_dec, dimg * n + 64));
	CK(cudaMalloc((void **)&d_len, sizeof(int32_t) * n));
	CK(cudaMemcpy(d_pix, h_pix, img * n, cudaMemcpyHostToDevice));
	cudaStream_t st;
	cudaEvent_t e0, e1, e2;
	CK(cudaStreamCreate(&st));
	CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1)); CK(cudaEventCreate(&e2));
	for (int r = 0; r <= j->runs; r++) { /* run 0 is the warm-up */
		CK(cudaEventRecord(e0, st));
		if (qoi_gpu_encode(plan, d_pix, d_str, d_len, st) != QOI_GPU_OK) { j->ok = 0; return NULL; }
		CK(cudaEventRecord(e1, st));
		if (qoi_gpu_decode(plan, d_str, d_len, d_dec, 1, NULL, st) != QOI_GPU_OK) { j->ok = 0; return NULL; }
		CK(cudaEventRecord(e2, st));
		CK(cudaStreamSynchronize(st));
		float a, b;
		CK(cudaEventElapsedTime(&a, e0, e1));
		CK(cudaEventElapsedTime(&b, e1, e2));
		if (r > 0) { j->enc_ms += a; j->dec_ms += b; }
	}
	j->enc_ms /= j->runs;
	j->dec_ms /= j->runs;
	/* verify: decoded pixels == input (RGB input is compared on its three channels) */
	uint8_t *h_dec = (uint8_t *)malloc(dimg * n);
	int32_t *h_len = (int32_t *)malloc(sizeof(int32_t) * n);
	CK(cudaMemcpy(h_dec, d_dec, dimg * n, cudaMemcpyDeviceToHost));
	CK(cudaMemcpy(h_len, d_len, sizeof(int32_t) * n, cudaMemcpyDeviceToHost));
	for (size_t p = 0; p < (size_t)c.w * c.h * n && j->ok; p++) {
		if (memcmp(h_dec + p * 4, h_pix + p * c.ch, (size_t)(c.ch == 4 ? 4 : 3)) != 0 || (c.ch == 3 && h_dec[p * 4 + 3] != 255)) j->ok = 0;
	}
	for (int i = 0; i < n; i++) j->enc_bytes += (uint64_t)h_len[i];
	free(h_dec); free(h_len); free(h_pix); free(descs); free(poff); free(soff);
	cudaFree(d_pix); cudaFree(d_str); cudaFree(d_dec); cudaFree(d_len);
	cudaEventDestroy(e0); cudaEventDestroy(e1); cudaEventDestroy(e2); cudaStreamDestroy(st);
	qoi_gpu_plan_destroy(plan);
	qoi_gpu_ctx_destroy(ctx);
	return NULL;
}

int main(int argc, char **argv) {
	int runs = argc > 1 ? atoi(argv[1]) : 5, gpus = 1, batch_override = 0, sweep = 0;
	double peak = 6490.5; /* MEASURED_PEAKS.json hbm_gbs on this pool */
	for (int i = 2; i < argc; i++) {
		if (!strcmp(argv[i], "--gpus") && i + 1 < argc) gpus = atoi(argv[++i]);
		else if (!strcmp(argv[i], "--batch") && i + 1 < argc) batch_override = atoi(argv[++i]);
		else if (!strcmp(argv[i], "--peak") && i + 1 < argc) peak = atof(argv[++i]);
		else if (!strcmp(argv[i], "--sweep")) sweep = 1;
	}
	if (runs < 1) runs = 1;
	int ndev = qoi_gpu_device_count();
	if (ndev < 1) { fprintf(stderr, "no CUDA device: this harness has no CPU path\n"); return 1; }
	if (gpus > ndev) gpus = ndev;
	config_t sweep_cfgs[18];
	char sweep_names[18][64];
	int n_sweep = 0;
	for (int sz = 64; sz <= 16384; sz *= 2) {
		for (int ch = 3; ch <= 4; ch++) {
			long long b = (1ll << 29) / ((long long)sz * sz); /* ~0.5 Gpx per GPU, at least one image */
			if (b < 1) b = 1;
			if (b > 16384) b = 16384;
			snprintf(sweep_names[n_sweep], sizeof sweep_names[n_sweep], "%dx%d %s photo", sz, sz, ch == 3 ? "RGB" : "RGBA");
			sweep_cfgs[n_sweep] = (config_t){sz, sz, ch, 4, (int)b, sweep_names[n_sweep]};
			n_sweep++;
		}
	}
	const config_t fixed_cfgs[] = {
	    {256, 256, 4, 4, 1, "256x256 RGBA photo x1 (configs[0] shape)"},
	    {4096, 4096, 4, 4, 1, "4096x4096 RGBA photo x1 (configs[1])"},
	    {1920, 1080, 3, 4, 1024, "1920x1080 RGB photo x1024 (configs[2])"},
	    {4096, 4096, 4, 4, 32, "4096x4096 RGBA photo x32"},
	    {4096, 4096, 4, 0, 32, "4096x4096 RGBA noise x32 (worst case)"},
	    {4096, 4096, 4, 7, 32, "4096x4096 RGBA const0 x32 (best case)"},
	    {8192, 8192, 4, 4, 8, "8192x8192 RGBA photo x8 (configs[3] image size)"},
	};
	const config_t *cfgs = sweep ? sweep_cfgs : fixed_cfgs;
	const size_t n_cfgs = sweep ? (size_t)n_sweep : sizeof fixed_cfgs / sizeof fixed_cfgs[0];
	printf("qoibench_gpu: %d run(s), %d GPU(s), peak %.1f GB/s\n", runs, gpus, peak);
	for (size_t k = 0; k < n_cfgs; k++) {
		config_t c = cfgs[k];
		if (batch_override && c.batch > 1) c.batch = batch_override;
		if (c.batch > 1) c.batch *= gpus; /* weak scaling: the per-GPU batch stays fixed */
		job_t jobs[16];
		pthread_t th[16];
		for (int g = 0; g < gpus; g++) {
			jobs[g].device = g; jobs[g].rank = g; jobs[g].ranks = gpus; jobs[g].runs = runs; jobs[g].cfg = c;
			pthread_create(&th[g], NULL, run_job, &jobs[g]);
		}
		double enc = 0, dec = 0;
		uint64_t bytes = 0;
		int ok = 1;
		for (int g = 0; g < gpus; g++) {
			pthread_join(th[g], NULL);
			if (jobs[g].enc_ms > enc) enc = jobs[g].enc_ms; /* max over GPUs */
			if (jobs[g].dec_ms > dec) dec = jobs[g].dec_ms;
			bytes += jobs[g].enc_bytes;
			ok = ok && jobs[g].ok;
		}
		const double px = (double)c.w * c.h * c.batch, raw = px * c.ch;
		printf("## %s, batch %d%s\n", c.name, c.batch, ok ? "" : "   *** FAILED ***");
		printf("          decode ms   encode ms   decode mpps   encode mpps   size kb    rate   enc GB/s (%%peak)   dec GB/s (%%peak)\n");
		printf("qoi_gpu:   %8.3f    %8.3f    %10.1f    %10.1f  %8llu   %4.1f%%   %8.1f (%4.1f%%)   %8.1f (%4.1f%%)\n", dec, enc,
		       dec > 0 ? px / (dec * 1e3) : 0, enc > 0 ? px / (enc * 1e3) : 0, (unsigned long long)(bytes / 1024 / (uint64_t)c.batch),
		       100.0 * (double)bytes / raw, (raw + (double)bytes) / (enc * 1e6), 100.0 * (raw + (double)bytes) / (enc * 1e6) / (peak * gpus),
		       ((double)bytes + px * 4) / (dec * 1e6), 100.0 * ((double)bytes + px * 4) / (dec * 1e6) / (peak * gpus));
		if (!ok) return 1;
	}
	return 0;
}
