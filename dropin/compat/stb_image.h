/*
 * stb_image.h -- STAND-IN, not Sean Barrett's library. It provides the handful of stb_image entry
 * points the reference's tools call (qoibench.c:387,395,441; qoiconv.c:46,56) for PNG files only
 * (both tools define STBI_ONLY_PNG), on top of dropin/compat/pngmini.h (zlib). See pngmini.h for why
 * it exists. Same conventions as the real library for these calls: 8 bits per channel, `comp` reports
 * the channels of the file (palette -> 3 or 4, tRNS adds alpha, 16-bit is reduced to 8), `req_comp`
 * forces the channel count of the result, results are freed with stbi_image_free() / free().
 */
#ifndef STBI_STANDIN_INCLUDE_STB_IMAGE_H
#define STBI_STANDIN_INCLUDE_STB_IMAGE_H

#include <stdio.h>
#include <stdlib.h>
#include <string.h>

typedef unsigned char stbi_uc;

#ifdef __cplusplus
extern "C" {
#endif

stbi_uc *stbi_load(char const *filename, int *x, int *y, int *channels_in_file, int desired_channels);
stbi_uc *stbi_load_from_memory(stbi_uc const *buffer, int len, int *x, int *y, int *channels_in_file, int desired_channels);
int stbi_info(char const *filename, int *x, int *y, int *comp);
int stbi_info_from_memory(stbi_uc const *buffer, int len, int *x, int *y, int *comp);
void stbi_image_free(void *retval_from_stbi_load);
const char *stbi_failure_reason(void);

#ifdef __cplusplus
}
#endif

#endif /* STBI_STANDIN_INCLUDE_STB_IMAGE_H */

#if defined(STB_IMAGE_IMPLEMENTATION) && !defined(STBI_STANDIN_IMPLEMENTED)
#define STBI_STANDIN_IMPLEMENTED

#include "pngmini.h"

static const char *stbi__standin_reason = "";

const char *stbi_failure_reason(void) { return stbi__standin_reason; }

void stbi_image_free(void *p) { free(p); }

static unsigned char *stbi__standin_slurp(const char *filename, size_t *size, size_t limit) {
	FILE *f = fopen(filename, "rb");
	if (!f) { stbi__standin_reason = "can't fopen"; return NULL; }
	unsigned char *buf = NULL;
	size_t n = 0;
	if (limit == 0) {
		fseek(f, 0, SEEK_END);
		long end = ftell(f);
		fseek(f, 0, SEEK_SET);
		limit = end > 0 ? (size_t)end : 0;
	}
	buf = (unsigned char *)malloc(limit ? limit : 1);
	if (buf) n = fread(buf, 1, limit, f);
	fclose(f);
	if (!buf) { stbi__standin_reason = "outofmem"; return NULL; }
	*size = n;
	return buf;
}

int stbi_info_from_memory(stbi_uc const *buffer, int len, int *x, int *y, int *comp) {
	pngmini_info in;
	if (len < 0 || !pngmini_decode(buffer, (size_t)len, &in, PNGMINI_ALL, NULL)) { stbi__standin_reason = "not a PNG"; return 0; }
	if (x) *x = (int)in.w;
	if (y) *y = (int)in.h;
	if (comp) *comp = pngmini_channels(&in, PNGMINI_ALL);
	return 1;
}

int stbi_info(char const *filename, int *x, int *y, int *comp) {
	size_t size = 0;
	/* IHDR, PLTE and tRNS precede the first IDAT; 64 KiB covers them together with any ancillary chunks in practice */
	unsigned char *buf = stbi__standin_slurp(filename, &size, 65536);
	if (!buf) return 0;
	int ok = size <= 0x7fffffff && stbi_info_from_memory(buf, (int)size, x, y, comp);
	if (!ok && size == 65536) { /* an unusually long run of ancillary chunks: look at the whole file */
		free(buf);
		buf = stbi__standin_slurp(filename, &size, 0);
		ok = buf && size <= 0x7fffffff && stbi_info_from_memory(buf, (int)size, x, y, comp);
	}
	free(buf);
	return ok;
}

stbi_uc *stbi_load_from_memory(stbi_uc const *buffer, int len, int *x, int *y, int *channels_in_file, int desired_channels) {
	pngmini_info in;
	unsigned char *px = NULL;
	if (desired_channels < 0 || desired_channels > 4) { stbi__standin_reason = "bad req_comp"; return NULL; }
	if (len < 0 || !pngmini_decode(buffer, (size_t)len, &in, PNGMINI_ALL, &px)) { stbi__standin_reason = "bad PNG"; return NULL; }
	int nc = pngmini_channels(&in, PNGMINI_ALL);
	if (desired_channels) px = pngmini_convert(px, nc, desired_channels, in.w, in.h);
	if (!px) { stbi__standin_reason = "outofmem"; return NULL; }
	if (x) *x = (int)in.w;
	if (y) *y = (int)in.h;
	if (channels_in_file) *channels_in_file = nc;
	return px;
}

stbi_uc *stbi_load(char const *filename, int *x, int *y, int *channels_in_file, int desired_channels) {
	size_t size = 0;
	unsigned char *buf = stbi__standin_slurp(filename, &size, 0);
	if (!buf) return NULL;
	stbi_uc *px = size <= 0x7fffffff ? stbi_load_from_memory(buf, (int)size, x, y, channels_in_file, desired_channels) : NULL;
	free(buf);
	return px;
}

#endif /* STB_IMAGE_IMPLEMENTATION */
