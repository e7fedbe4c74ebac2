/*
 * striped_smoke.c -- the shortest possible GPU check of the experimental encoders (cuda/qoi_enc4.cuh): encode one batch
 * with the shipped kernel and with each QOI_GPU_ENC variant, require byte-identical streams and lengths, decode the
 * variant's output with its own restart index and require the input back, and time both with CUDA events.
 * C + cudart only (no Python start-up), so that it fits into a gpurun call of a few seconds:
 *     make -C bench striped_smoke && gpurun --timeout 60 -- 'timeout 50 bench/striped_smoke > gpurun_out/striped_smoke.txt 2>&1'
 */
#include <cuda_runtime_api.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "qoi_gpu.h"

void qoi_synth_batch(uint8_t *dst, size_t stride, int n, uint32_t w, uint32_t h, int channels, int dist, uint32_t first_index);

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e_)); return 2; } } while (0)

int main(int argc, char **argv) {
	const int n = argc > 1 ? atoi(argv[1]) : 8, W = 4096, H = 4096, CH = 4, runs = 5;
	/* QOI_GPU_ENC value, optionally "+half": the same encoder with QOI_GPU_DEC=half (indexed decode with two lanes per tile).
	   Names given on the command line replace the built-in list -- e.g. for the experimental library:
	   LD_PRELOAD=qoi_b200/libqoi_gpu_exp.so bench/striped_smoke 32 earlyw mid+half earlyw-mid+half */
	static const char *builtin[] = {"", "striped", "striped-earlyw", "striped-match", "slab", "slab-match", "slab-pipe", "slab-pipe-ballot", "slab4", "slab4-pipe",
	                                "slab-match+half", "slab4-pipe+half"};
	const char *variants[64];
	int nvar = 0;
	variants[nvar++] = "";
	if (argc > 2) for (int i = 2; i < argc && nvar < 64; i++) variants[nvar++] = argv[i];
	else for (size_t i = 1; i < sizeof builtin / sizeof builtin[0]; i++) variants[nvar++] = builtin[i];
	const size_t img = (size_t)W * H * CH;
	qoi_desc *descs = (qoi_desc *)calloc((size_t)n, sizeof(qoi_desc));
	uint64_t *poff = (uint64_t *)calloc((size_t)n, 8), *soff = (uint64_t *)calloc((size_t)n, 8);
	size_t slot = 0;
	for (int i = 0; i < n; i++) {
		descs[i] = (qoi_desc){W, H, CH, QOI_SRGB};
		slot = (qoi_gpu_max_stream_size(&descs[i]) + 255) & ~(size_t)255;
		poff[i] = (uint64_t)i * img;
		soff[i] = (uint64_t)i * slot;
	}
	uint8_t *h_pix = (uint8_t *)malloc(img * n), *h_ref = (uint8_t *)malloc(slot * n), *h_got = (uint8_t *)malloc(slot * n), *h_dec = (uint8_t *)malloc(img * n);
	int32_t *l_ref = (int32_t *)malloc(4u * n), *l_got = (int32_t *)malloc(4u * n);
	qoi_synth_batch(h_pix, img, n, W, H, CH, 4, 0);
	uint8_t *d_pix, *d_str, *d_dec;
	int32_t *d_len;
	CK(cudaMalloc((void **)&d_pix, img * n));
	CK(cudaMalloc((void **)&d_str, slot * n));
	CK(cudaMalloc((void **)&d_dec, img * n));
	CK(cudaMalloc((void **)&d_len, 4u * n));
	CK(cudaMemcpy(d_pix, h_pix, img * n, cudaMemcpyHostToDevice));
	cudaStream_t st;
	cudaEvent_t e0, e1, e2;
	CK(cudaStreamCreate(&st));
	CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1)); CK(cudaEventCreate(&e2));
	int bad = 0;
	for (int v = 0; v < nvar; v++) {
		char encname[64];
		snprintf(encname, sizeof encname, "%s", variants[v]);
		char *plus = strchr(encname, 0x2B);
		if (plus) *plus = 0;
		setenv("QOI_GPU_ENC", encname, 1); /* both read by qoi_gpu_ctx_create */
		setenv("QOI_GPU_DEC", plus ? plus + 1 : "", 1);
		qoi_gpu_ctx *ctx = NULL;
		qoi_gpu_plan *plan = NULL;
		if (qoi_gpu_ctx_create(&ctx, 0) != QOI_GPU_OK || qoi_gpu_plan_create(ctx, &plan, n, descs, poff, soff, NULL, 4) != QOI_GPU_OK) { printf("setup failed\n"); return 2; }
		CK(cudaMemset(d_str, 0xEE, slot * n));
		float enc = 0, dec = 0;
		for (int r = 0; r <= runs; r++) {
			CK(cudaEventRecord(e0, st));
			if (qoi_gpu_encode(plan, d_pix, d_str, d_len, st) != QOI_GPU_OK) { printf("encode failed: %s\n", qoi_gpu_last_error(ctx)); return 2; }
			CK(cudaEventRecord(e1, st));
			if (qoi_gpu_decode(plan, d_str, d_len, d_dec, 1, NULL, st) != QOI_GPU_OK) { printf("decode failed\n"); return 2; }
			CK(cudaEventRecord(e2, st));
			CK(cudaStreamSynchronize(st));
			float a, b;
			CK(cudaEventElapsedTime(&a, e0, e1));
			CK(cudaEventElapsedTime(&b, e1, e2));
			if (r) { enc += a; dec += b; }
		}
		uint8_t *hs = v == 0 ? h_ref : h_got;
		int32_t *hl = v == 0 ? l_ref : l_got;
		CK(cudaMemcpy(hs, d_str, slot * n, cudaMemcpyDeviceToHost));
		CK(cudaMemcpy(hl, d_len, 4u * n, cudaMemcpyDeviceToHost));
		CK(cudaMemcpy(h_dec, d_dec, img * n, cudaMemcpyDeviceToHost));
		int same = 1, rt = memcmp(h_dec, h_pix, img * n) == 0;
		if (v) {
			for (int i = 0; i < n && same; i++) same = l_got[i] == l_ref[i] && memcmp(h_got + soff[i], h_ref + soff[i], (size_t)l_ref[i]) == 0;
		}
		printf("%-17s encode %7.3f ms (%6.1f Gpx/s)  indexed decode %7.3f ms   streams %s   round trip %s\n", v ? variants[v] : "shipped", enc / runs,
		       (double)W * H * n / (enc / runs * 1e6), dec / runs, v ? (same ? "IDENTICAL to shipped" : "*** DIFFER ***") : "(reference for the others)",
		       rt ? "ok" : "*** MISMATCH ***");
		fflush(stdout);
		bad += !same || !rt;
		qoi_gpu_plan_destroy(plan);
		qoi_gpu_ctx_destroy(ctx);
	}
	printf(bad ? "FAILED\n" : "ALL OK\n");
	return bad ? 1 : 0;
}
