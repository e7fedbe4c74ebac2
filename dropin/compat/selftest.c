/*
 * selftest.c -- drives the three stand-ins of dropin/compat/ the way the reference's tools do and
 * dumps what they produce, so that tests/test_zz_dropin_tools.py can compare it with independently
 * computed pixels. PNG I/O only; no QOI code is involved.
 *
 *   selftest decode in.png out.raw      -> stbi_info + stbi_load(req = 3 if the file has 3 channels else 4)
 *                                          (qoibench.c:387-395), raw pixels to out.raw; prints "w h file_channels used_channels"
 *   selftest pngdecode in.png out.raw   -> the libpng call sequence of qoibench.c:183-251 (always RGBA); prints "w h"
 *   selftest encode in.raw w h ch stb.png libpng.png
 *                                       -> stbi_write_png (qoiconv.c:73) and the libpng write sequence of qoibench.c:115-162
 */
#include <png.h>

#define STB_IMAGE_IMPLEMENTATION
#define STBI_ONLY_PNG
#define STBI_NO_LINEAR
#include "stb_image.h"

#define STB_IMAGE_WRITE_IMPLEMENTATION
#include "stb_image_write.h"

typedef struct { size_t pos, size; unsigned char *data; } mem_t;

static void rd(png_structp png, png_bytep out, png_size_t n) {
	mem_t *m = (mem_t *)png_get_io_ptr(png);
	if (m->pos + n > m->size) { fprintf(stderr, "read past the end\n"); exit(3); }
	memcpy(out, m->data + m->pos, n);
	m->pos += n;
}

static void wr(png_structp png, png_bytep data, png_size_t n) {
	FILE *f = (FILE *)png_get_io_ptr(png);
	if (fwrite(data, 1, n, f) != n) exit(3);
}

static void warn(png_structp png, png_const_charp msg) { (void)png; (void)msg; }

static unsigned char *slurp(const char *path, size_t *n) {
	FILE *f = fopen(path, "rb");
	if (!f) { perror(path); exit(2); }
	fseek(f, 0, SEEK_END);
	long sz = ftell(f);
	fseek(f, 0, SEEK_SET);
	unsigned char *b = (unsigned char *)malloc(sz > 0 ? (size_t)sz : 1);
	*n = fread(b, 1, (size_t)sz, f);
	fclose(f);
	return b;
}

static void dump(const char *path, const void *p, size_t n) {
	FILE *f = fopen(path, "wb");
	if (!f || fwrite(p, 1, n, f) != n) { perror(path); exit(2); }
	fclose(f);
}

int main(int argc, char **argv) {
	if (argc == 4 && strcmp(argv[1], "decode") == 0) {
		int w, h, ch;
		if (!stbi_info(argv[2], &w, &h, &ch)) { fprintf(stderr, "stbi_info failed: %s\n", stbi_failure_reason()); return 1; }
		int use = ch != 3 ? 4 : 3;
		int w2, h2;
		unsigned char *px = stbi_load(argv[2], &w2, &h2, NULL, use);
		if (!px || w2 != w || h2 != h) { fprintf(stderr, "stbi_load failed: %s\n", stbi_failure_reason()); return 1; }
		dump(argv[3], px, (size_t)w * h * use);
		printf("%d %d %d %d\n", w, h, ch, use);
		stbi_image_free(px);
		return 0;
	}
	if (argc == 4 && strcmp(argv[1], "pngdecode") == 0) {
		mem_t m = {0, 0, NULL};
		m.data = slurp(argv[2], &m.size);
		png_structp png = png_create_read_struct(PNG_LIBPNG_VER_STRING, NULL, NULL, warn);
		png_infop info = png_create_info_struct(png);
		if (setjmp(png_jmpbuf(png))) { fprintf(stderr, "png error\n"); return 1; }
		png_set_read_fn(png, &m, rd);
		png_set_sig_bytes(png, 0);
		png_read_info(png, info);
		png_uint_32 w, h;
		int depth, ctype, il;
		png_get_IHDR(png, info, &w, &h, &depth, &ctype, &il, NULL, NULL);
		png_set_strip_16(png);
		if (depth < 8) png_set_packing(png);
		if (ctype & PNG_COLOR_MASK_PALETTE) png_set_expand(png);
		if (!(ctype & PNG_COLOR_MASK_COLOR)) png_set_gray_to_rgb(png);
		if (png_get_valid(png, info, PNG_INFO_tRNS)) png_set_tRNS_to_alpha(png);
		if (!(ctype & PNG_COLOR_MASK_ALPHA)) png_set_filler(png, 255, PNG_FILLER_AFTER);
		png_read_update_info(png, info);
		unsigned char *out = (unsigned char *)malloc((size_t)w * h * 4);
		png_bytep *rows = (png_bytep *)malloc(sizeof(png_bytep) * h);
		for (png_uint_32 y = 0; y < h; y++) rows[y] = out + (size_t)y * w * 4;
		png_read_image(png, rows);
		png_read_end(png, info);
		png_destroy_read_struct(&png, &info, NULL);
		dump(argv[3], out, (size_t)w * h * 4);
		printf("%u %u\n", w, h);
		return 0;
	}
	if (argc == 8 && strcmp(argv[1], "encode") == 0) {
		size_t n;
		unsigned char *px = slurp(argv[2], &n);
		int w = atoi(argv[3]), h = atoi(argv[4]), ch = atoi(argv[5]);
		if (n != (size_t)w * h * ch) { fprintf(stderr, "size mismatch\n"); return 1; }
		if (!stbi_write_png(argv[6], w, h, ch, px, 0)) { fprintf(stderr, "stbi_write_png failed\n"); return 1; }
		FILE *f = fopen(argv[7], "wb");
		if (!f) { perror(argv[7]); return 2; }
		png_structp png = png_create_write_struct(PNG_LIBPNG_VER_STRING, NULL, NULL, NULL);
		png_infop info = png_create_info_struct(png);
		if (setjmp(png_jmpbuf(png))) { fprintf(stderr, "png error\n"); return 1; }
		png_set_IHDR(png, info, (png_uint_32)w, (png_uint_32)h, 8, ch == 3 ? PNG_COLOR_TYPE_RGB : PNG_COLOR_TYPE_RGBA, PNG_INTERLACE_NONE,
		             PNG_COMPRESSION_TYPE_DEFAULT, PNG_FILTER_TYPE_DEFAULT);
		png_bytep *rows = (png_bytep *)malloc(sizeof(png_bytep) * (size_t)h);
		for (int y = 0; y < h; y++) rows[y] = px + (size_t)y * w * ch;
		png_set_rows(png, info, rows);
		png_set_write_fn(png, f, wr, NULL);
		png_write_png(png, info, PNG_TRANSFORM_IDENTITY, NULL);
		png_destroy_write_struct(&png, &info);
		fclose(f);
		return 0;
	}
	fprintf(stderr, "usage: selftest decode|pngdecode in.png out.raw | encode in.raw w h ch stb.png libpng.png\n");
	return 64;
}
