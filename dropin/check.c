/*
 * check.c -- a dependency-free caller written the way the reference's tools are
 * (qoibench.c:27-28, qoiconv.c:28-29): it defines QOI_IMPLEMENTATION, includes "qoi.h" and
 * then uses qoi_encode / qoi_decode / qoi_write / qoi_read and frees the results with free().
 * Built against include/qoi.h + libqoi_gpu.so it proves the header swap and the link line of
 * INTEGRATION.md without stb_image / libpng (absent from this image).
 * Protocol: the round-trip self check of qoibench.c:410-417 and its BENCHMARK_FN timing (:364-376).
 */
#define QOI_IMPLEMENTATION
#include "qoi.h"

#include <stdint.h>
#include <stdio.h>
#include <string.h>
#include <time.h>

static uint64_t ns(void) {
	struct timespec ts;
	clock_gettime(CLOCK_MONOTONIC, &ts);
	return (uint64_t)ts.tv_sec * 1000000000ull + (uint64_t)ts.tv_nsec;
}

int main(int argc, char **argv) {
	int w = argc > 1 ? atoi(argv[1]) : 1024, h = argc > 2 ? atoi(argv[2]) : 1024, runs = argc > 3 ? atoi(argv[3]) : 5;
	int channels = 4;
	size_t n = (size_t)w * h * channels;
	unsigned char *pixels = (unsigned char *)malloc(n);
	uint32_t s = 0x9E3779B9u ^ 4u;
	for (int y = 0; y < h; y++)
		for (int x = 0; x < w; x++) { /* the "photo" distribution of SURVEY.md 8(d) */
			s ^= s << 13; s ^= s >> 17; s ^= s << 5;
			unsigned char *p = pixels + ((size_t)y * w + x) * 4;
			p[0] = (unsigned char)(x * 255u / w + (s & 15));
			p[1] = (unsigned char)(y * 255u / h + ((s >> 4) & 15));
			p[2] = (unsigned char)((x + y) * 255u / (w + h) + ((s >> 8) & 15));
			p[3] = 255;
		}
	qoi_desc desc = {(unsigned)w, (unsigned)h, (unsigned char)channels, QOI_SRGB};
	int size = 0;
	void *enc = qoi_encode(pixels, &desc, &size);
	if (!enc) { fprintf(stderr, "qoi_encode returned NULL (no CUDA device?)\n"); return 1; }
	qoi_desc dc;
	void *dec = qoi_decode(enc, size, &dc, channels);
	if (!dec || memcmp(dec, pixels, n) != 0) { fprintf(stderr, "QOI roundtrip pixel mismatch\n"); return 1; }
	free(dec);
	uint64_t te = 0, td = 0;
	for (int i = 0; i <= runs; i++) { /* first run discarded, like BENCHMARK_FN */
		uint64_t t0 = ns();
		int sz;
		void *e = qoi_encode(pixels, &desc, &sz);
		uint64_t t1 = ns();
		void *d = qoi_decode(enc, size, &dc, 4);
		uint64_t t2 = ns();
		free(e); free(d);
		if (i > 0) { te += t1 - t0; td += t2 - t1; }
	}
	double px = (double)w * h;
	printf("          decode ms   encode ms   decode mpps   encode mpps   size kb    rate\n");
	printf("qoi_gpu:   %8.2f    %8.2f      %8.2f      %8.2f  %8d   %4.1f%%\n", td / 1e6 / runs, te / 1e6 / runs,
	       px / (td / 1e3 / runs), px / (te / 1e3 / runs), size / 1024, 100.0 * size / (double)n);
	if (!qoi_write("/tmp/qoi_dropin_check.qoi", pixels, &desc)) { fprintf(stderr, "qoi_write failed\n"); return 1; }
	void *rd = qoi_read("/tmp/qoi_dropin_check.qoi", &dc, 0);
	if (!rd || dc.width != (unsigned)w || memcmp(rd, pixels, n) != 0) { fprintf(stderr, "qoi_read mismatch\n"); return 1; }
	free(rd); free(enc); free(pixels);
	remove("/tmp/qoi_dropin_check.qoi");
	printf("dropin check ok\n");
	return 0;
}
