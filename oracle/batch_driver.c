/*
 * batch_driver.c -- times a CPU codec over a batch of images the way the
 * reference harness does (qoibench.c:446-450,471-481: the result buffer's
 * malloc and free are inside the timed call), optionally with OpenMP over the
 * images.  Compiled twice by oracle/Makefile:
 *   -DDRV_REF     -> calls qoi_encode/qoi_decode of the unmodified reference  (oracle/_ref/libqoi_ref.so)
 *   (default)     -> calls ora_encode/ora_decode of the restatement           (oracle/liboracle.so)
 * TEST INFRASTRUCTURE ONLY.
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>
#ifdef _OPENMP
#include <omp.h>
#endif

typedef struct {
	unsigned int width, height;
	unsigned char channels, colorspace;
} drv_desc;

#ifdef DRV_REF
void *qoi_encode(const void *data, const drv_desc *desc, int *out_len);
void *qoi_decode(const void *data, int size, drv_desc *desc, int channels);
#define ENC qoi_encode
#define DEC qoi_decode
#define SYM(x) ref_##x
#else
void *ora_encode(const void *data, const drv_desc *desc, int *out_len);
void *ora_decode(const void *data, int size, drv_desc *desc, int channels);
#define ENC ora_encode
#define DEC ora_decode
#define SYM(x) ora_##x
#endif

static double now_s(void) {
	struct timespec ts;
	clock_gettime(CLOCK_MONOTONIC, &ts); /* qoibench.c:67-77 */
	return (double)ts.tv_sec + 1e-9 * (double)ts.tv_nsec;
}

int SYM(host_threads)(void) {
#ifdef _OPENMP
	return omp_get_max_threads();
#else
	return 1;
#endif
}

/*
 * Encode n images (image i at pixels + i*pix_stride), `threads` OpenMP threads
 * (<=1: serial).  If keep != NULL the streams are copied to keep + i*keep_stride
 * and lens[i] is set; otherwise each result is freed right away.  Returns the
 * wall time in seconds of the whole pass, or -1 on a NULL result.
 */
double SYM(encode_batch)(const uint8_t *pixels, size_t pix_stride, int n, const drv_desc *desc, int threads,
                         uint8_t *keep, size_t keep_stride, int *lens) {
	int failed = 0;
	double t0 = now_s();
#pragma omp parallel for schedule(dynamic) num_threads(threads > 1 ? threads : 1)
	for (int i = 0; i < n; i++) {
		int len = 0;
		void *enc = ENC(pixels + (size_t)i * pix_stride, desc, &len);
		if (!enc) { failed = 1; continue; }
		if (keep) memcpy(keep + (size_t)i * keep_stride, enc, (size_t)len);
		if (lens) lens[i] = len;
		free(enc);
	}
	double t1 = now_s();
	return failed ? -1.0 : t1 - t0;
}

/* Decode n streams (stream i at streams + i*stream_stride, lens[i] bytes) to out_channels. */
double SYM(decode_batch)(const uint8_t *streams, size_t stream_stride, const int *lens, int n, int out_channels,
                         int threads, uint8_t *keep, size_t keep_stride) {
	int failed = 0;
	double t0 = now_s();
#pragma omp parallel for schedule(dynamic) num_threads(threads > 1 ? threads : 1)
	for (int i = 0; i < n; i++) {
		drv_desc d;
		void *px = DEC(streams + (size_t)i * stream_stride, lens[i], &d, out_channels);
		if (!px) { failed = 1; continue; }
		if (keep) {
			int ch = out_channels ? out_channels : d.channels;
			memcpy(keep + (size_t)i * keep_stride, px, (size_t)d.width * d.height * (size_t)ch);
		}
		free(px);
	}
	double t1 = now_s();
	return failed ? -1.0 : t1 - t0;
}
