/*
 * png.h -- STAND-IN, not libpng: the subset of the libpng 1.6 API that the reference's benchmark
 * uses for its "libpng" column (qoibench.c:106-251), header-only on top of dropin/compat/pngmini.h
 * (zlib), so that the UNCHANGED qoibench.c compiles without libpng (link with -lz instead of
 * -lpng). See pngmini.h for why it exists. It pulls bytes through the caller's read function in the
 * order libpng does (signature, chunk header, chunk body + CRC; png_read_info stops at the first
 * IDAT header, png_read_image consumes the IDAT run, png_read_end the rest up to IEND) and reports
 * failures through longjmp(png_jmpbuf(png)) after calling the error / warning callbacks.
 *
 * Transform calls follow libpng's meaning: png_set_strip_16 (always on: results are 8-bit),
 * png_set_packing (one byte per sample, values unscaled), png_set_expand / png_set_tRNS_to_alpha /
 * png_set_palette_to_rgb (palette -> RGB, grey scaled to 8 bits, tRNS -> alpha),
 * png_set_gray_to_rgb, png_set_filler(..., PNG_FILLER_AFTER).
 */
#ifndef PNG_STANDIN_H
#define PNG_STANDIN_H

#include <setjmp.h>
#include <stddef.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "pngmini.h"

#define PNG_LIBPNG_VER_STRING "1.6.0-standin"

#define PNG_COLOR_MASK_PALETTE 1
#define PNG_COLOR_MASK_COLOR 2
#define PNG_COLOR_MASK_ALPHA 4
#define PNG_COLOR_TYPE_GRAY 0
#define PNG_COLOR_TYPE_PALETTE (PNG_COLOR_MASK_COLOR | PNG_COLOR_MASK_PALETTE)
#define PNG_COLOR_TYPE_RGB (PNG_COLOR_MASK_COLOR)
#define PNG_COLOR_TYPE_RGB_ALPHA (PNG_COLOR_MASK_COLOR | PNG_COLOR_MASK_ALPHA)
#define PNG_COLOR_TYPE_GRAY_ALPHA (PNG_COLOR_MASK_ALPHA)
#define PNG_COLOR_TYPE_RGBA PNG_COLOR_TYPE_RGB_ALPHA
#define PNG_COLOR_TYPE_GA PNG_COLOR_TYPE_GRAY_ALPHA
#define PNG_INTERLACE_NONE 0
#define PNG_INTERLACE_ADAM7 1
#define PNG_COMPRESSION_TYPE_BASE 0
#define PNG_COMPRESSION_TYPE_DEFAULT PNG_COMPRESSION_TYPE_BASE
#define PNG_FILTER_TYPE_BASE 0
#define PNG_FILTER_TYPE_DEFAULT PNG_FILTER_TYPE_BASE
#define PNG_TRANSFORM_IDENTITY 0
#define PNG_INFO_PLTE 0x0008U
#define PNG_INFO_tRNS 0x0010U
#define PNG_FILLER_BEFORE 0
#define PNG_FILLER_AFTER 1

typedef unsigned char png_byte;
typedef png_byte *png_bytep;
typedef png_byte **png_bytepp;
typedef const png_byte *png_const_bytep;
typedef unsigned int png_uint_32;
typedef size_t png_size_t;
typedef char *png_charp;
typedef const char *png_const_charp;
typedef void *png_voidp;

struct png_standin_struct;
struct png_standin_info;
typedef struct png_standin_struct png_struct;
typedef struct png_standin_info png_info;
typedef png_struct *png_structp;
typedef png_struct **png_structpp;
typedef png_info *png_infop;
typedef png_info **png_infopp;

typedef void (*png_error_ptr)(png_structp, png_const_charp);
typedef void (*png_rw_ptr)(png_structp, png_bytep, png_size_t);
typedef void (*png_flush_ptr)(png_structp);

struct png_standin_info {
	pngmini_info hdr;
	png_uint_32 valid;
	png_bytepp rows; /* png_set_rows */
};

struct png_standin_struct {
	jmp_buf jb;
	int is_read;
	png_voidp error_ptr, io_ptr;
	png_error_ptr error_fn, warn_fn;
	png_rw_ptr rw_fn;
	int sig_bytes;
	/* read state */
	png_uint_32 pending_len;   /* body length of the chunk whose header has been read */
	unsigned char pending_type[4];
	int have_pending, info_read, ended;
	struct png_standin_info *cur_info;
	int packing, expand, gray_to_rgb, filler, filler_byte, filler_after;
	/* write state */
	png_uint_32 w, h;
	int depth, ctype;
};

#define png_jmpbuf(png_ptr) ((png_ptr)->jb)

static void png_standin_fail(png_structp png, const char *msg) {
	if (png->error_fn) png->error_fn(png, msg);
	longjmp(png->jb, 1);
}

static png_structp png_standin_create(int is_read, png_voidp error_ptr, png_error_ptr error_fn, png_error_ptr warn_fn) {
	png_structp png = (png_structp)calloc(1, sizeof(png_struct));
	if (!png) return NULL;
	png->is_read = is_read;
	png->error_ptr = error_ptr;
	png->error_fn = error_fn;
	png->warn_fn = warn_fn;
	return png;
}

static png_structp png_create_write_struct(png_const_charp ver, png_voidp error_ptr, png_error_ptr error_fn, png_error_ptr warn_fn) {
	(void)ver;
	return png_standin_create(0, error_ptr, error_fn, warn_fn);
}

static png_structp png_create_read_struct(png_const_charp ver, png_voidp error_ptr, png_error_ptr error_fn, png_error_ptr warn_fn) {
	(void)ver;
	return png_standin_create(1, error_ptr, error_fn, warn_fn);
}

static png_infop png_create_info_struct(png_structp png) {
	(void)png;
	return (png_infop)calloc(1, sizeof(png_info));
}

static void png_destroy_write_struct(png_structpp png, png_infopp info) {
	if (info && *info) { free(*info); *info = NULL; }
	if (png && *png) { free(*png); *png = NULL; }
}

static void png_destroy_read_struct(png_structpp png, png_infopp info, png_infopp end_info) {
	if (end_info && *end_info) { free(*end_info); *end_info = NULL; }
	png_destroy_write_struct(png, info);
}

static png_voidp png_get_io_ptr(png_structp png) { return png->io_ptr; }
static png_voidp png_get_error_ptr(png_structp png) { return png->error_ptr; }

static void png_set_write_fn(png_structp png, png_voidp io_ptr, png_rw_ptr write_fn, png_flush_ptr flush_fn) {
	(void)flush_fn;
	png->io_ptr = io_ptr;
	png->rw_fn = write_fn;
}

static void png_set_read_fn(png_structp png, png_voidp io_ptr, png_rw_ptr read_fn) {
	png->io_ptr = io_ptr;
	png->rw_fn = read_fn;
}

static void png_set_sig_bytes(png_structp png, int n) { png->sig_bytes = n; }

/* ------------------------------------------------------------------ write */

static void png_set_IHDR(png_structp png, png_infop info, png_uint_32 w, png_uint_32 h, int depth, int ctype, int interlace,
                         int compression, int filter) {
	(void)info;
	if (depth != 8 || interlace != PNG_INTERLACE_NONE || compression != 0 || filter != 0 ||
	    (ctype != PNG_COLOR_TYPE_GRAY && ctype != PNG_COLOR_TYPE_GRAY_ALPHA && ctype != PNG_COLOR_TYPE_RGB && ctype != PNG_COLOR_TYPE_RGBA))
		png_standin_fail(png, "png_set_IHDR: only 8-bit non-interlaced grey / grey+alpha / RGB / RGBA is supported");
	png->w = w; png->h = h; png->depth = depth; png->ctype = ctype;
}

static void png_set_rows(png_structp png, png_infop info, png_bytepp rows) {
	(void)png;
	info->rows = rows;
}

static void png_standin_write_sink(void *ctx, const void *data, size_t len) {
	png_structp png = (png_structp)ctx;
	png->rw_fn(png, (png_bytep)data, len);
}

static void png_write_png(png_structp png, png_infop info, int transforms, png_voidp params) {
	(void)params;
	if (transforms != PNG_TRANSFORM_IDENTITY) png_standin_fail(png, "png_write_png: only PNG_TRANSFORM_IDENTITY");
	if (!png->rw_fn || !info->rows || png->w == 0) png_standin_fail(png, "png_write_png: IHDR, rows and a write function are required");
	int comp = png->ctype == PNG_COLOR_TYPE_GRAY ? 1 : png->ctype == PNG_COLOR_TYPE_GRAY_ALPHA ? 2 : png->ctype == PNG_COLOR_TYPE_RGB ? 3 : 4;
	if (!pngmini_encode(png_standin_write_sink, png, (const unsigned char *const *)info->rows, png->w, png->h, comp, Z_DEFAULT_COMPRESSION))
		png_standin_fail(png, "png_write_png: out of memory");
}

/* ------------------------------------------------------------------ read */

static void png_standin_pull(png_structp png, unsigned char *dst, size_t n) {
	if (!png->rw_fn) png_standin_fail(png, "no read function");
	png->rw_fn(png, dst, n);
}

/* read one chunk header into pending_* */
static void png_standin_next_header(png_structp png) {
	unsigned char hd[8];
	png_standin_pull(png, hd, 8);
	png->pending_len = pngmini_be32(hd);
	memcpy(png->pending_type, hd + 4, 4);
	png->have_pending = 1;
	if (png->pending_len > 0x7fffffffu) png_standin_fail(png, "chunk length out of range");
}

/* body + CRC of the pending chunk; returns a malloc'd body (caller frees) */
static unsigned char *png_standin_body(png_structp png) {
	size_t len = png->pending_len;
	unsigned char *body = (unsigned char *)malloc(len + 4);
	if (!body) png_standin_fail(png, "out of memory");
	png_standin_pull(png, body, len + 4);
	uLong c = crc32(crc32(0L, Z_NULL, 0), png->pending_type, 4);
	if (len) c = crc32(c, body, (uInt)len);
	png->have_pending = 0;
	if ((unsigned)c != pngmini_be32(body + len)) {
		free(body);
		png_standin_fail(png, "CRC error");
	}
	return body;
}

static void png_read_info(png_structp png, png_infop info) {
	if (png->sig_bytes < 8) {
		unsigned char sig[8];
		size_t need = (size_t)(8 - png->sig_bytes);
		png_standin_pull(png, sig, need);
		if (memcmp(sig, pngmini_sig + png->sig_bytes, need) != 0) png_standin_fail(png, "Not a PNG file");
	}
	int seen_ihdr = 0;
	for (;;) {
		png_standin_next_header(png);
		if (memcmp(png->pending_type, "IDAT", 4) == 0) break;
		unsigned len = png->pending_len;
		unsigned char type[4];
		memcpy(type, png->pending_type, 4);
		unsigned char *body = png_standin_body(png);
		int ok = 1;
		if (!seen_ihdr) {
			ok = memcmp(type, "IHDR", 4) == 0 && len == 13 && pngmini_parse_ihdr(body, &info->hdr);
			seen_ihdr = 1;
		} else if (memcmp(type, "PLTE", 4) == 0) {
			ok = pngmini_parse_plte(body, len, &info->hdr);
			if (ok) info->valid |= PNG_INFO_PLTE;
		} else if (memcmp(type, "tRNS", 4) == 0) {
			ok = pngmini_parse_trns(body, len, &info->hdr);
			if (ok) info->valid |= PNG_INFO_tRNS;
		} else if (memcmp(type, "IEND", 4) == 0) {
			ok = 0;
		} else if (!(type[0] & 32)) {
			ok = 0; /* unknown critical chunk */
		} else if (png->warn_fn && memcmp(type, "iCCP", 4) == 0) {
			/* libpng is chatty about colour profiles; callers install a warning handler to mute it */
		}
		free(body);
		if (!ok) png_standin_fail(png, "invalid chunk before IDAT");
	}
	if (!seen_ihdr) png_standin_fail(png, "IDAT before IHDR");
	png->info_read = 1;
	png->cur_info = info; /* libpng keeps the header fields inside png_struct; png_read_image has no info argument */
}

static png_uint_32 png_get_IHDR(png_structp png, png_infop info, png_uint_32 *w, png_uint_32 *h, int *depth, int *ctype, int *interlace,
                                int *compression, int *filter) {
	(void)png;
	if (w) *w = info->hdr.w;
	if (h) *h = info->hdr.h;
	if (depth) *depth = info->hdr.depth;
	if (ctype) *ctype = info->hdr.ctype;
	if (interlace) *interlace = info->hdr.interlace;
	if (compression) *compression = 0;
	if (filter) *filter = 0;
	return 1;
}

static png_uint_32 png_get_valid(png_structp png, png_infop info, png_uint_32 flag) {
	(void)png;
	return info->valid & flag;
}

static void png_set_strip_16(png_structp png) { (void)png; }
static void png_set_packing(png_structp png) { png->packing = 1; }
static void png_set_expand(png_structp png) { png->expand = 1; }
static void png_set_palette_to_rgb(png_structp png) { png->expand = 1; }
static void png_set_tRNS_to_alpha(png_structp png) { png->expand = 1; }
static void png_set_expand_gray_1_2_4_to_8(png_structp png) { png->expand = 1; }
static void png_set_gray_to_rgb(png_structp png) { png->gray_to_rgb = 1; }
static void png_set_filler(png_structp png, png_uint_32 filler, int where) {
	png->filler = 1;
	png->filler_byte = (int)(filler & 255);
	png->filler_after = where == PNG_FILLER_AFTER;
}
static void png_read_update_info(png_structp png, png_infop info) { (void)png; (void)info; }

/* channels of a row png_read_image writes under the transforms set so far */
static int png_standin_out_channels(png_structp png, png_infop info, int *base_nc, int *flags) {
	int f = png->expand ? PNGMINI_ALL : 0;
	int nc = pngmini_channels(&info->hdr, f), out = nc;
	if (png->gray_to_rgb && out <= 2) out += 2;
	if (png->filler && (out == 1 || out == 3)) out += 1;
	*base_nc = nc;
	*flags = f;
	return out;
}

static png_size_t png_get_rowbytes(png_structp png, png_infop info) {
	int nc, f;
	int out = png_standin_out_channels(png, info, &nc, &f);
	if (info->hdr.depth < 8 && !png->packing && !png->expand) return pngmini_rowbytes(info->hdr.w, info->hdr.depth, 1);
	return (png_size_t)info->hdr.w * (png_size_t)out;
}

static void png_read_image(png_structp png, png_bytepp rows) {
	png_infop info = png->cur_info;
	if (!png->info_read || !info) png_standin_fail(png, "png_read_image before png_read_info");
	if (info->hdr.depth < 8 && !png->packing && !png->expand)
		png_standin_fail(png, "packed sub-byte rows are not supported: call png_set_packing or png_set_expand");
	/* the run of IDAT chunks */
	unsigned char *idat = NULL;
	size_t idat_len = 0, cap = 0;
	while (png->have_pending && memcmp(png->pending_type, "IDAT", 4) == 0) {
		size_t len = png->pending_len;
		unsigned char *body = png_standin_body(png);
		if (idat_len + len > cap) {
			cap = (idat_len + len) * 2;
			unsigned char *n = (unsigned char *)realloc(idat, cap);
			if (!n) { free(idat); free(body); png_standin_fail(png, "out of memory"); }
			idat = n;
		}
		memcpy(idat + idat_len, body, len);
		idat_len += len;
		free(body);
		png_standin_next_header(png);
	}
	int nc, flags;
	int out = png_standin_out_channels(png, info, &nc, &flags);
	unsigned char *px = pngmini_decode_idat(&info->hdr, idat, idat_len, flags);
	free(idat);
	if (!px) png_standin_fail(png, "corrupt image data");
	unsigned w = info->hdr.w, h = info->hdr.h;
	for (unsigned y = 0; y < h; y++) {
		const unsigned char *s = px + (size_t)y * w * nc;
		unsigned char *d = rows[y];
		if (out == nc) { memcpy(d, s, (size_t)w * nc); continue; }
		for (unsigned x = 0; x < w; x++, s += nc) {
			unsigned char c[4];
			int n = 0;
			if (nc <= 2 && png->gray_to_rgb) { c[0] = c[1] = c[2] = s[0]; n = 3; if (nc == 2) c[n++] = s[1]; }
			else { for (; n < nc; n++) c[n] = s[n]; }
			if (png->filler && (n == 1 || n == 3)) {
				if (png->filler_after) c[n++] = (unsigned char)png->filler_byte;
				else { for (int k = n; k > 0; k--) c[k] = c[k - 1]; c[0] = (unsigned char)png->filler_byte; n++; }
			}
			memcpy(d, c, (size_t)n);
			d += n;
		}
	}
	free(px);
}

static void png_read_end(png_structp png, png_infop info) {
	(void)info;
	while (!png->ended) {
		if (!png->have_pending) png_standin_next_header(png);
		int is_end = memcmp(png->pending_type, "IEND", 4) == 0;
		int critical = !(png->pending_type[0] & 32) && !is_end && memcmp(png->pending_type, "IDAT", 4) != 0;
		free(png_standin_body(png));
		if (critical) png_standin_fail(png, "unknown critical chunk");
		if (is_end) png->ended = 1;
	}
}

/* Silence "defined but not used" for the entry points a given caller does not touch. */
#if defined(__GNUC__)
__attribute__((unused))
#endif
static void png_standin_unused(void) {
	(void)png_create_write_struct; (void)png_create_read_struct; (void)png_create_info_struct; (void)png_destroy_write_struct;
	(void)png_destroy_read_struct; (void)png_get_io_ptr; (void)png_get_error_ptr; (void)png_set_write_fn; (void)png_set_read_fn;
	(void)png_set_sig_bytes; (void)png_set_IHDR; (void)png_set_rows; (void)png_write_png; (void)png_get_IHDR; (void)png_get_valid;
	(void)png_set_strip_16; (void)png_set_packing; (void)png_set_expand; (void)png_set_palette_to_rgb; (void)png_set_tRNS_to_alpha;
	(void)png_set_expand_gray_1_2_4_to_8; (void)png_set_gray_to_rgb; (void)png_set_filler; (void)png_read_update_info;
	(void)png_get_rowbytes; (void)png_read_info; (void)png_read_image; (void)png_read_end; (void)png_standin_unused;
	(void)pngmini_decode; (void)pngmini_convert;
}

#endif /* PNG_STANDIN_H */
