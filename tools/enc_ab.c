/*
 * enc_ab.c -- the shortest GPU check + timing of the codec kernels, for A/B runs of different builds of the library.
 * MEASUREMENT / TEST TOOL (not part of the product): C + cudart only, no Python start-up, so a gpurun call costs seconds.
 *
 *     bench/enc_ab <libqoi_gpu*.so> <n images> <w> <h> <channels> <dist> [<dist> ...]
 *
 * For every distribution: encode the batch (device resident), decode it with the restart index and as foreign streams,
 * CUDA-event times (mean of 5 after one warm-up), round trip checked on every image, and -- where oracle/_ref/libqoi_ref.so
 * or oracle/liboracle.so is present -- the first and last image's stream compared byte for byte with the CPU checker.
 * The library is loaded with dlopen so that several builds (e.g. different launch bounds) can be timed in one session.
 */
#ifndef _GNU_SOURCE
#define _GNU_SOURCE
#endif
#include <cuda_runtime_api.h>
#include <dlfcn.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "qoi_gpu.h"

void qoi_synth_batch(uint8_t *dst, size_t stride, int n, uint32_t w, uint32_t h, int channels, int dist, uint32_t first_index);

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e_)); return 2; } } while (0)

typedef int (*fn_ctx_create)(qoi_gpu_ctx **, int);
typedef void (*fn_ctx_destroy)(qoi_gpu_ctx *);
typedef int (*fn_plan_create)(qoi_gpu_ctx *, qoi_gpu_plan **, int, const qoi_desc *, const uint64_t *, const uint64_t *, const uint64_t *, int);
typedef void (*fn_plan_destroy)(qoi_gpu_plan *);
typedef int (*fn_encode)(qoi_gpu_plan *, const void *, void *, int32_t *, void *);
typedef int (*fn_decode)(qoi_gpu_plan *, const void *, const int32_t *, void *, int, int32_t *, void *);
typedef const char *(*fn_err)(const qoi_gpu_ctx *);
typedef int (*fn_fstats)(qoi_gpu_plan *, uint32_t *);
typedef void *(*fn_cpu_encode)(const void *, const qoi_desc *, int *);

int main(int argc, char **argv) {
	if (argc < 7) { printf("usage: enc_ab <lib> <n> <w> <h> <ch> <dist>...\n"); return 2; }
	void *lib = dlopen(argv[1], RTLD_NOW | RTLD_LOCAL);
	if (!lib) { printf("dlopen: %s\n", dlerror()); return 2; }
	fn_ctx_create ctx_create = (fn_ctx_create)dlsym(lib, "qoi_gpu_ctx_create");
	fn_ctx_destroy ctx_destroy = (fn_ctx_destroy)dlsym(lib, "qoi_gpu_ctx_destroy");
	fn_plan_create plan_create = (fn_plan_create)dlsym(lib, "qoi_gpu_plan_create");
	fn_plan_destroy plan_destroy = (fn_plan_destroy)dlsym(lib, "qoi_gpu_plan_destroy");
	fn_encode encode = (fn_encode)dlsym(lib, "qoi_gpu_encode");
	fn_decode decode = (fn_decode)dlsym(lib, "qoi_gpu_decode");
	fn_err last_error = (fn_err)dlsym(lib, "qoi_gpu_last_error");
	fn_fstats fstats = (fn_fstats)dlsym(lib, "qoi_gpu_plan_fstats");
	if (!ctx_create || !plan_create || !encode || !decode) { printf("missing symbols\n"); return 2; }
	/* CPU checker (test infrastructure): the compiled reference if it travelled with the repository, else the restatement */
	fn_cpu_encode cpu_encode = NULL;
	const char *cpu_name = "none";
	{
		void *r = dlopen("oracle/_ref/libqoi_ref.so", RTLD_NOW | RTLD_LOCAL);
		if (r) { cpu_encode = (fn_cpu_encode)dlsym(r, "qoi_encode"); cpu_name = "compiled reference"; }
		if (!cpu_encode) {
			r = dlopen("oracle/liboracle.so", RTLD_NOW | RTLD_LOCAL);
			if (r) { cpu_encode = (fn_cpu_encode)dlsym(r, "ora_encode"); cpu_name = "oracle port"; }
		}
	}
	const int n = atoi(argv[2]), W = atoi(argv[3]), H = atoi(argv[4]), CH = atoi(argv[5]), runs = 5;
	const int with_foreign = getenv("ENC_AB_FOREIGN") != NULL;
	/* decode to 4 channels like qoibench.c:446-450, or to the image's own count with ENC_AB_DEC_CH=3 (RGB input only) */
	const int DCH = (CH == 3 && getenv("ENC_AB_DEC_CH") && atoi(getenv("ENC_AB_DEC_CH")) == 3) ? 3 : 4;
	const size_t img = (size_t)W * H * CH, dimg = (size_t)W * H * DCH;
	qoi_desc *descs = (qoi_desc *)calloc((size_t)n, sizeof(qoi_desc));
	uint64_t *poff = (uint64_t *)calloc((size_t)n, 8), *soff = (uint64_t *)calloc((size_t)n, 8);
	size_t slot = 0;
	for (int i = 0; i < n; i++) {
		descs[i] = (qoi_desc){(unsigned)W, (unsigned)H, (unsigned char)CH, QOI_SRGB};
		slot = ((size_t)W * H * 5 + 22 + 255) & ~(size_t)255;
		poff[i] = (uint64_t)i * img;
		soff[i] = (uint64_t)i * slot;
	}
	uint8_t *h_pix = (uint8_t *)malloc(img * n), *h_got = (uint8_t *)malloc(slot * n), *h_dec = (uint8_t *)malloc(dimg * n);
	int32_t *l_got = (int32_t *)malloc(4u * n);
	uint8_t *d_pix, *d_str, *d_dec;
	int32_t *d_len;
	CK(cudaMalloc((void **)&d_pix, img * n));
	CK(cudaMalloc((void **)&d_str, slot * n));
	CK(cudaMalloc((void **)&d_dec, dimg * n));
	CK(cudaMalloc((void **)&d_len, 4u * n));
	cudaStream_t st;
	cudaEvent_t e0, e1, e2, e3;
	CK(cudaStreamCreate(&st));
	CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1)); CK(cudaEventCreate(&e2)); CK(cudaEventCreate(&e3));
	qoi_gpu_ctx *ctx = NULL;
	qoi_gpu_plan *plan = NULL;
	if (ctx_create(&ctx, 0) != QOI_GPU_OK || plan_create(ctx, &plan, n, descs, poff, soff, NULL, DCH) != QOI_GPU_OK) { printf("setup failed\n"); return 2; }
	printf("library %s, checker %s, %d x %dx%dx%d\n", argv[1], cpu_name, n, W, H, CH);
	int bad = 0;
	for (int a = 6; a < argc; a++) {
		const int dist = atoi(argv[a]);
		qoi_synth_batch(h_pix, img, n, (uint32_t)W, (uint32_t)H, CH, dist, getenv("ENC_AB_FIRST") ? (uint32_t)atoi(getenv("ENC_AB_FIRST")) : 0u);
		CK(cudaMemcpy(d_pix, h_pix, img * n, cudaMemcpyHostToDevice));
		CK(cudaMemset(d_str, 0xEE, slot * n));
		float enc = 0, dec = 0, fdec = 0;
		for (int r = 0; r <= runs; r++) {
			CK(cudaEventRecord(e0, st));
			if (encode(plan, d_pix, d_str, d_len, st) != QOI_GPU_OK) { printf("encode failed: %s\n", last_error(ctx)); return 2; }
			CK(cudaEventRecord(e1, st));
			if (decode(plan, d_str, d_len, d_dec, 1, NULL, st) != QOI_GPU_OK) { printf("decode failed\n"); return 2; }
			CK(cudaEventRecord(e2, st));
			CK(cudaStreamSynchronize(st));
			float x, y;
			CK(cudaEventElapsedTime(&x, e0, e1));
			CK(cudaEventElapsedTime(&y, e1, e2));
			if (r) { enc += x; dec += y; }
		}
		CK(cudaMemcpy(h_got, d_str, slot * n, cudaMemcpyDeviceToHost));
		CK(cudaMemcpy(l_got, d_len, 4u * n, cudaMemcpyDeviceToHost));
		CK(cudaMemcpy(h_dec, d_dec, dimg * n, cudaMemcpyDeviceToHost));
		int rt = 1, rtf = 1, same = 1;
		if (CH == DCH) rt = memcmp(h_dec, h_pix, img * n) == 0;
		else for (size_t k = 0; k < (size_t)W * H * n && rt; k++) rt = h_dec[4 * k] == h_pix[3 * k] && h_dec[4 * k + 1] == h_pix[3 * k + 1] && h_dec[4 * k + 2] == h_pix[3 * k + 2] && h_dec[4 * k + 3] == 255;
		if (with_foreign) {
			for (int r = 0; r <= 2; r++) {
				CK(cudaMemsetAsync(d_dec, 0x5A, dimg * n, st));
				CK(cudaEventRecord(e2, st));
				if (decode(plan, d_str, d_len, d_dec, 0, NULL, st) != QOI_GPU_OK) { printf("foreign decode failed\n"); return 2; }
				CK(cudaEventRecord(e3, st));
				CK(cudaStreamSynchronize(st));
				float z;
				CK(cudaEventElapsedTime(&z, e2, e3));
				if (r) fdec += z;
			}
			CK(cudaMemcpy(h_dec, d_dec, dimg * n, cudaMemcpyDeviceToHost));
			if (CH == DCH) rtf = memcmp(h_dec, h_pix, img * n) == 0;
			else for (size_t k = 0; k < (size_t)W * H * n && rtf; k++) rtf = h_dec[4 * k] == h_pix[3 * k] && h_dec[4 * k + 1] == h_pix[3 * k + 1] && h_dec[4 * k + 2] == h_pix[3 * k + 2] && h_dec[4 * k + 3] == 255;
		}
		double bytes = 0;
		for (int i = 0; i < n; i++) bytes += l_got[i];
		if (cpu_encode) {
			const int idx[2] = {0, n - 1};
			for (int k = 0; k < (n > 1 ? 2 : 1) && same; k++) {
				int len = 0;
				uint8_t *ref = (uint8_t *)cpu_encode(h_pix + poff[idx[k]], &descs[idx[k]], &len);
				same = ref && len == l_got[idx[k]] && memcmp(ref, h_got + soff[idx[k]], (size_t)len) == 0;
				free(ref);
			}
		}
		const double px = (double)W * H * n;
		printf("dist %d  e=%.3f B/px  encode %7.3f ms (%6.1f Gpx/s, %5.0f GB/s)  indexed decode %7.3f ms (%6.1f Gpx/s, %5.0f GB/s)", dist, bytes / px,
		       enc / runs, px / (enc / runs * 1e6), (px * CH + bytes) / (enc / runs * 1e6), dec / runs, px / (dec / runs * 1e6), (px * 4 + bytes) / (dec / runs * 1e6));
		if (with_foreign) printf("  foreign decode %8.3f ms (%6.1f Gpx/s) %s", fdec / 2, px / (fdec / 2 * 1e6), rtf ? "ok" : "*** MISMATCH ***");
		if (with_foreign && fstats && getenv("ENC_AB_FSTATS")) { /* segments walked / streams unsettled per state round */
			uint32_t fs[128];
			if (fstats(plan, fs) == QOI_GPU_OK) {
				printf("\n    rounds:");
				for (int r = 1; r < 20 && (fs[r] || fs[64 + r - 1]); r++) printf(" %d:%u/%u", r, fs[r], fs[64 + r - 1]);
				printf("  (segments walked / streams unsettled before)");
				for (int q = 120; q < 128; q++) if (fs[q]) printf("%s%u", q == 120 ? "  fall-back for images " : ",", fs[q] - 1u);
				printf("\n   ");
			}
		}
		printf("  streams %s  round trip %s\n", cpu_encode ? (same ? "== checker" : "*** DIFFER ***") : "(unchecked)", rt ? "ok" : "*** MISMATCH ***");
		fflush(stdout);
		bad += !same || !rt || !rtf;
	}
	plan_destroy(plan);
	ctx_destroy(ctx);
	printf(bad ? "FAILED\n" : "ALL OK\n");
	return bad ? 1 : 0;
}
