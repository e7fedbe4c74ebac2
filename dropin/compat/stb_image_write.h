/*
 * stb_image_write.h -- STAND-IN, not Sean Barrett's library: the two PNG writers the reference's
 * tools call (qoibench.c:466 stbi_write_png_to_func, qoiconv.c:73 stbi_write_png) on top of
 * dropin/compat/pngmini.h (zlib deflate, adaptive row filters). See pngmini.h for why it exists.
 * Conventions kept: comp 1..4 = grey, grey+alpha, RGB, RGBA; stride_in_bytes 0 means w*comp;
 * non-zero return on success; the callback may be invoked several times.
 */
#ifndef STBIW_STANDIN_INCLUDE_STB_IMAGE_WRITE_H
#define STBIW_STANDIN_INCLUDE_STB_IMAGE_WRITE_H

#include <stdio.h>
#include <stdlib.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef void stbi_write_func(void *context, void *data, int size);

extern int stbi_write_png_compression_level;

int stbi_write_png(char const *filename, int w, int h, int comp, const void *data, int stride_in_bytes);
int stbi_write_png_to_func(stbi_write_func *func, void *context, int w, int h, int comp, const void *data, int stride_in_bytes);

#ifdef __cplusplus
}
#endif

#endif /* STBIW_STANDIN_INCLUDE_STB_IMAGE_WRITE_H */

#if defined(STB_IMAGE_WRITE_IMPLEMENTATION) && !defined(STBIW_STANDIN_IMPLEMENTED)
#define STBIW_STANDIN_IMPLEMENTED

#include "pngmini.h"

int stbi_write_png_compression_level = 8;

typedef struct { stbi_write_func *func; void *context; } stbiw__standin_sink;

static void stbiw__standin_emit(void *ctx, const void *data, size_t len) {
	stbiw__standin_sink *s = (stbiw__standin_sink *)ctx;
	while (len) { /* the callback takes an int */
		int n = len > 0x40000000u ? 0x40000000 : (int)len;
		s->func(s->context, (void *)data, n);
		data = (const unsigned char *)data + n;
		len -= (size_t)n;
	}
}

int stbi_write_png_to_func(stbi_write_func *func, void *context, int w, int h, int comp, const void *data, int stride_in_bytes) {
	if (!func || !data || w <= 0 || h <= 0 || comp < 1 || comp > 4) return 0;
	size_t stride = stride_in_bytes ? (size_t)stride_in_bytes : (size_t)w * comp;
	const unsigned char **rows = (const unsigned char **)malloc(sizeof(*rows) * (size_t)h);
	if (!rows) return 0;
	for (int y = 0; y < h; y++) rows[y] = (const unsigned char *)data + stride * (size_t)y;
	stbiw__standin_sink sink = {func, context};
	int level = stbi_write_png_compression_level;
	int ok = pngmini_encode(stbiw__standin_emit, &sink, rows, (unsigned)w, (unsigned)h, comp, level < 0 ? 0 : level > 9 ? 9 : level);
	free(rows);
	return ok;
}

static void stbiw__standin_fwrite(void *context, void *data, int size) {
	FILE **f = (FILE **)context;
	if (*f && fwrite(data, 1, (size_t)size, *f) != (size_t)size) { fclose(*f); *f = NULL; }
}

int stbi_write_png(char const *filename, int w, int h, int comp, const void *data, int stride_in_bytes) {
	FILE *f = fopen(filename, "wb");
	if (!f) return 0;
	int ok = stbi_write_png_to_func(stbiw__standin_fwrite, &f, w, h, comp, data, stride_in_bytes);
	if (!f) return 0; /* a write failed; the callback closed the file */
	if (fclose(f) != 0) return 0;
	return ok;
}

#endif /* STB_IMAGE_WRITE_IMPLEMENTATION */
