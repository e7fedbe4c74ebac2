/*
 * qoi.h -- drop-in header for the B200-native QOI codec (libqoi_gpu.so).
 *
 * This header takes the place of the reference's single-header library for a
 * program that is built against libqoi_gpu.so: it declares the same four
 * entry points with the same C ABI (reference qoi.h:236-241 qoi_desc, :252
 * qoi_write, :265 qoi_read, :278 qoi_encode, :289 qoi_decode) and, when the
 * including file defines QOI_IMPLEMENTATION (as qoibench.c:27 / qoiconv.c:28
 * do), it supplies only the two stdio convenience wrappers.  The codec itself
 * -- qoi_encode / qoi_decode -- is NOT in this header: those two symbols are
 * resolved at link time from cuda/qoi_gpu.cu (sm_100a kernels).  There is no
 * CPU implementation behind them; without a CUDA device they return NULL.
 *
 * Contract kept from the reference:
 *   - results are heap blocks the caller releases with free();
 *   - failure is NULL (pointer functions) or 0 (qoi_write), nothing else;
 *   - *out_len is written only on success;
 *   - qoi_decode fills *desc from the header before it validates the header.
 *   - QOI_NO_STDIO removes qoi_read / qoi_write.
 */
#ifndef QOI_H
#define QOI_H

#ifdef __cplusplus
extern "C" {
#endif

#define QOI_SRGB   0
#define QOI_LINEAR 1

/* 12 bytes; layout fixed by the reference ABI (qoi.h:236-241) */
typedef struct {
	unsigned int width;
	unsigned int height;
	unsigned char channels;   /* 3 = RGB, 4 = RGBA */
	unsigned char colorspace; /* QOI_SRGB or QOI_LINEAR; informative only */
} qoi_desc;

/* pixels (width*height*channels bytes, row major) -> malloc'd .qoi stream, length in *out_len */
void *qoi_encode(const void *data, const qoi_desc *desc, int *out_len);

/* .qoi stream -> malloc'd pixels; channels: 0 = as in the header, 3 or 4 = forced output stride */
void *qoi_decode(const void *data, int size, qoi_desc *desc, int channels);

#ifndef QOI_NO_STDIO
/* encode and write to a file; returns bytes written, 0 on any failure */
int qoi_write(const char *filename, const void *data, const qoi_desc *desc);
/* read a file and decode it; NULL on any failure */
void *qoi_read(const char *filename, qoi_desc *desc, int channels);
#endif

#ifdef __cplusplus
}
#endif
#endif /* QOI_H */

/* ------------------------------------------------------------------------- */
#ifdef QOI_IMPLEMENTATION
#ifndef QOI_B200_IMPL_ONCE
#define QOI_B200_IMPL_ONCE
#include <stdlib.h>

/* Results of qoi_encode/qoi_decode come from malloc() inside libqoi_gpu.so, so
 * a custom QOI_MALLOC cannot be honoured across the library boundary; QOI_FREE
 * must therefore be free()-compatible (it is, by default). */
#ifndef QOI_FREE
#define QOI_FREE(p) free(p)
#endif

#ifndef QOI_NO_STDIO
#include <stdio.h>

/* behaviour of reference qoi.h:595-617 */
int qoi_write(const char *filename, const void *data, const qoi_desc *desc) {
	int n = 0, bad;
	void *stream;
	FILE *fp = fopen(filename, "wb");
	if (fp == NULL) return 0;
	stream = qoi_encode(data, desc, &n);
	if (stream == NULL) {
		fclose(fp);
		return 0;
	}
	fwrite(stream, 1, (size_t)n, fp);
	fflush(fp);
	bad = ferror(fp);
	fclose(fp);
	QOI_FREE(stream);
	return bad ? 0 : n;
}

/* behaviour of reference qoi.h:619-646 */
void *qoi_read(const char *filename, qoi_desc *desc, int channels) {
	long len;
	size_t got;
	void *stream, *px = NULL;
	FILE *fp = fopen(filename, "rb");
	if (fp == NULL) return NULL;
	if (fseek(fp, 0, SEEK_END) != 0 || (len = ftell(fp)) <= 0 || fseek(fp, 0, SEEK_SET) != 0) {
		fclose(fp);
		return NULL;
	}
	stream = malloc((size_t)len);
	if (stream == NULL) {
		fclose(fp);
		return NULL;
	}
	got = fread(stream, 1, (size_t)len, fp);
	fclose(fp);
	if (got == (size_t)len) px = qoi_decode(stream, (int)len, desc, channels);
	free(stream);
	return px;
}
#endif /* QOI_NO_STDIO */
#endif /* QOI_B200_IMPL_ONCE */
#endif /* QOI_IMPLEMENTATION */
