/*
 * pngmini.h -- a small PNG codec over zlib, header-only (all functions static).
 *
 * Why it exists: the reference's tools (qoibench.c:17-25, qoiconv.c:23-26) include <png.h>,
 * "stb_image.h" and "stb_image_write.h" for their PNG input/output. None of the three is in this
 * image and there is no network, so dropin/compat/ carries stand-ins with the same entry points
 * (stb_image.h, stb_image_write.h, png.h beside this file), all built on this codec. They exist so
 * that the UNCHANGED tool sources compile and link against include/qoi.h + libqoi_gpu.so; they are
 * not QOI code and not on the hot path. With the real stb / libpng headers on the include path the
 * tools build the same way and these files are not needed.
 *
 * Coverage: PNG 1.2 critical chunks (IHDR, PLTE, IDAT, IEND) + tRNS; colour types 0, 2, 3, 4, 6; bit
 * depths 1, 2, 4, 8, 16; Adam7 interlace; CRC and zlib checks. Output is always 8 bits per sample.
 * Encoder: 8-bit grey / grey+alpha / RGB / RGBA, per-row adaptive filter, zlib deflate.
 */
#ifndef PNGMINI_H
#define PNGMINI_H

#include <stdlib.h>
#include <string.h>
#include <zlib.h>

typedef struct {
	unsigned w, h;
	int depth, ctype, interlace;
	int npal, ntrns;           /* palette entries / tRNS entries (palette images) */
	int has_trns;              /* tRNS chunk seen */
	unsigned char pal[256 * 3];
	unsigned char trns[256];   /* palette alpha */
	unsigned short key[3];     /* tRNS colour key of grey / RGB images, at file depth */
} pngmini_info;

#define PNGMINI_SCALE    1   /* scale 1/2/4-bit grey to 0..255 (otherwise raw sample values)   */
#define PNGMINI_PALETTE  2   /* expand palette indices to RGB (RGBA if tRNS and PNGMINI_TRNS)  */
#define PNGMINI_TRNS     4   /* turn a tRNS chunk into an alpha channel                        */
#define PNGMINI_ALL      7

static unsigned pngmini_be32(const unsigned char *p) {
	return ((unsigned)p[0] << 24) | ((unsigned)p[1] << 16) | ((unsigned)p[2] << 8) | p[3];
}

static void pngmini_put32(unsigned char *p, unsigned v) {
	p[0] = (unsigned char)(v >> 24); p[1] = (unsigned char)(v >> 16); p[2] = (unsigned char)(v >> 8); p[3] = (unsigned char)v;
}

static int pngmini_samples(int ctype) {
	switch (ctype) { case 0: return 1; case 2: return 3; case 3: return 1; case 4: return 2; case 6: return 4; }
	return 0;
}

static const unsigned char pngmini_sig[8] = {137, 80, 78, 71, 13, 10, 26, 10};

/* IHDR payload (13 bytes) -> info; 0 on an invalid combination */
static int pngmini_parse_ihdr(const unsigned char *d, pngmini_info *in) {
	memset(in, 0, sizeof *in);
	in->w = pngmini_be32(d); in->h = pngmini_be32(d + 4);
	in->depth = d[8]; in->ctype = d[9]; in->interlace = d[12];
	if (in->w == 0 || in->h == 0 || in->w > 0x7fffffffu || in->h > 0x7fffffffu) return 0;
	if (d[10] != 0 || d[11] != 0 || in->interlace > 1) return 0;
	switch (in->ctype) {
	case 0: if (in->depth != 1 && in->depth != 2 && in->depth != 4 && in->depth != 8 && in->depth != 16) return 0; break;
	case 3: if (in->depth != 1 && in->depth != 2 && in->depth != 4 && in->depth != 8) return 0; break;
	case 2: case 4: case 6: if (in->depth != 8 && in->depth != 16) return 0; break;
	default: return 0;
	}
	return 1;
}

/* tRNS / PLTE payloads */
static int pngmini_parse_plte(const unsigned char *d, unsigned len, pngmini_info *in) {
	if (len % 3 != 0 || len > 768) return 0;
	memcpy(in->pal, d, len);
	in->npal = (int)(len / 3);
	return 1;
}

static int pngmini_parse_trns(const unsigned char *d, unsigned len, pngmini_info *in) {
	if (in->ctype == 3) {
		if (len > 256) return 0;
		memset(in->trns, 255, sizeof in->trns);
		memcpy(in->trns, d, len);
		in->ntrns = (int)len;
	} else if (in->ctype == 0) {
		if (len != 2) return 0;
		in->key[0] = (unsigned short)((d[0] << 8) | d[1]);
	} else if (in->ctype == 2) {
		if (len != 6) return 0;
		for (int i = 0; i < 3; i++) in->key[i] = (unsigned short)((d[2 * i] << 8) | d[2 * i + 1]);
	} else {
		return 0;
	}
	in->has_trns = 1;
	return 1;
}

/* number of 8-bit channels pngmini_expand produces under `flags` */
static int pngmini_channels(const pngmini_info *in, int flags) {
	int nc;
	if (in->ctype == 3) nc = (flags & PNGMINI_PALETTE) ? 3 : 1;
	else nc = pngmini_samples(in->ctype);
	if (in->has_trns && (flags & PNGMINI_TRNS) && (in->ctype == 0 || in->ctype == 2 || (in->ctype == 3 && (flags & PNGMINI_PALETTE)))) nc++;
	return nc;
}

static size_t pngmini_rowbytes(unsigned w, int depth, int ns) {
	return ((size_t)w * (size_t)(depth * ns) + 7) / 8;
}

static const unsigned char pngmini_a7x0[7] = {0, 4, 0, 2, 0, 1, 0}, pngmini_a7y0[7] = {0, 0, 4, 0, 2, 0, 1};
static const unsigned char pngmini_a7dx[7] = {8, 8, 4, 4, 2, 2, 1}, pngmini_a7dy[7] = {8, 8, 8, 4, 4, 2, 2};

/* size of the inflated IDAT payload */
static size_t pngmini_inflated_size(const pngmini_info *in) {
	int ns = pngmini_samples(in->ctype);
	if (!in->interlace) return (pngmini_rowbytes(in->w, in->depth, ns) + 1) * in->h;
	size_t n = 0;
	for (int p = 0; p < 7; p++) {
		unsigned pw = (in->w - pngmini_a7x0[p] + pngmini_a7dx[p] - 1) / pngmini_a7dx[p];
		unsigned ph = (in->h - pngmini_a7y0[p] + pngmini_a7dy[p] - 1) / pngmini_a7dy[p];
		if (in->w <= pngmini_a7x0[p] || in->h <= pngmini_a7y0[p]) continue;
		n += (pngmini_rowbytes(pw, in->depth, ns) + 1) * ph;
	}
	return n;
}

static int pngmini_paeth(int a, int b, int c) {
	int p = a + b - c, pa = abs(p - a), pb = abs(p - b), pc = abs(p - c);
	return (pa <= pb && pa <= pc) ? a : (pb <= pc ? b : c);
}

/* undo the row filters of one (sub)image in place; rows are 1 + rb bytes. 0 on a bad filter type */
static int pngmini_unfilter(unsigned char *d, size_t rb, unsigned rows, int bpp) {
	const unsigned char *prev = NULL;
	for (unsigned y = 0; y < rows; y++) {
		unsigned char *r = d + (rb + 1) * (size_t)y;
		int f = r[0];
		unsigned char *c = r + 1;
		switch (f) {
		case 0: break;
		case 1: for (size_t i = (size_t)bpp; i < rb; i++) c[i] = (unsigned char)(c[i] + c[i - bpp]); break;
		case 2: if (prev) for (size_t i = 0; i < rb; i++) c[i] = (unsigned char)(c[i] + prev[i]); break;
		case 3:
			for (size_t i = 0; i < rb; i++) {
				int a = i >= (size_t)bpp ? c[i - bpp] : 0, b = prev ? prev[i] : 0;
				c[i] = (unsigned char)(c[i] + ((a + b) >> 1));
			}
			break;
		case 4:
			for (size_t i = 0; i < rb; i++) {
				int a = i >= (size_t)bpp ? c[i - bpp] : 0, b = prev ? prev[i] : 0, cc = (prev && i >= (size_t)bpp) ? prev[i - bpp] : 0;
				c[i] = (unsigned char)(c[i] + pngmini_paeth(a, b, cc));
			}
			break;
		default: return 0;
		}
		prev = c;
	}
	return 1;
}

/* one unfiltered row of `pw` pixels -> 8-bit pixels of `nc` channels written at out + k*step*nc, k = 0..pw-1 */
static void pngmini_expand_row(const pngmini_info *in, int flags, const unsigned char *row, unsigned pw, unsigned char *out,
                               size_t step, int nc) {
	int ns = pngmini_samples(in->ctype), depth = in->depth;
	int want_alpha = in->has_trns && (flags & PNGMINI_TRNS);
	for (unsigned x = 0; x < pw; x++) {
		unsigned s[4] = {0, 0, 0, 0}; /* samples at file depth */
		if (depth == 8) for (int k = 0; k < ns; k++) s[k] = row[(size_t)x * ns + k];
		else if (depth == 16) for (int k = 0; k < ns; k++) s[k] = ((unsigned)row[((size_t)x * ns + k) * 2] << 8) | row[((size_t)x * ns + k) * 2 + 1];
		else { /* 1, 2, 4 bits: a single sample per pixel */
			size_t bit = (size_t)x * depth;
			s[0] = (row[bit >> 3] >> (8 - depth - (bit & 7))) & ((1u << depth) - 1);
		}
		unsigned char *o = out + (size_t)x * step * nc;
		if (in->ctype == 3) {
			unsigned idx = s[0];
			if (flags & PNGMINI_PALETTE) {
				if ((int)idx < in->npal) { o[0] = in->pal[idx * 3]; o[1] = in->pal[idx * 3 + 1]; o[2] = in->pal[idx * 3 + 2]; }
				else { o[0] = o[1] = o[2] = 0; }
				if (nc == 4) o[3] = (int)idx < in->ntrns ? in->trns[idx] : 255;
			} else {
				o[0] = (unsigned char)idx;
			}
			continue;
		}
		int k;
		for (k = 0; k < ns; k++) {
			unsigned v = s[k];
			if (depth == 16) v >>= 8;
			else if (depth < 8 && (flags & PNGMINI_SCALE)) v = v * 255u / ((1u << depth) - 1);
			o[k] = (unsigned char)v;
		}
		if (want_alpha && (in->ctype == 0 || in->ctype == 2)) {
			int match = 1;
			for (int j = 0; j < ns; j++) if (s[j] != in->key[j]) match = 0;
			o[k] = match ? 0 : 255;
		}
	}
}

/*
 * Inflate + unfilter + de-interlace + expand. `idat` is the concatenation of all IDAT payloads.
 * Returns a malloc'd w*h*nc buffer (nc = pngmini_channels(in, flags)) or NULL.
 */
static unsigned char *pngmini_decode_idat(const pngmini_info *in, const unsigned char *idat, size_t idat_len, int flags) {
	int ns = pngmini_samples(in->ctype), nc = pngmini_channels(in, flags);
	int bpp = (in->depth * ns + 7) / 8;
	size_t raw_len = pngmini_inflated_size(in);
	if (in->ctype == 3 && in->npal == 0) return NULL;
	if ((size_t)in->w * in->h > ((size_t)1 << 31) / 4) return NULL;
	unsigned char *raw = (unsigned char *)malloc(raw_len ? raw_len : 1);
	unsigned char *out = (unsigned char *)malloc((size_t)in->w * in->h * nc);
	if (!raw || !out) { free(raw); free(out); return NULL; }
	z_stream z;
	memset(&z, 0, sizeof z);
	int ok = inflateInit(&z) == Z_OK;
	if (ok) {
		size_t ipos = 0, opos = 0;
		int rc = Z_OK;
		while (rc == Z_OK) { /* feed in < 4 GiB pieces (uInt fields) */
			size_t ni = idat_len - ipos, no = raw_len - opos;
			z.next_in = (Bytef *)(idat + ipos); z.avail_in = (uInt)(ni > 0x40000000u ? 0x40000000u : ni);
			z.next_out = raw + opos; z.avail_out = (uInt)(no > 0x40000000u ? 0x40000000u : no);
			uInt gi = z.avail_in, go = z.avail_out;
			rc = inflate(&z, Z_NO_FLUSH);
			ipos += gi - z.avail_in; opos += go - z.avail_out;
			if (rc == Z_OK && gi == z.avail_in && go == z.avail_out) break; /* no progress: truncated or too long */
		}
		ok = (opos == raw_len) && (rc == Z_STREAM_END || rc == Z_OK);
		inflateEnd(&z);
	}
	if (ok) {
		if (!in->interlace) {
			size_t rb = pngmini_rowbytes(in->w, in->depth, ns);
			ok = pngmini_unfilter(raw, rb, in->h, bpp);
			for (unsigned y = 0; ok && y < in->h; y++)
				pngmini_expand_row(in, flags, raw + (rb + 1) * (size_t)y + 1, in->w, out + (size_t)y * in->w * nc, 1, nc);
		} else {
			unsigned char *p = raw;
			for (int pass = 0; ok && pass < 7; pass++) {
				if (in->w <= pngmini_a7x0[pass] || in->h <= pngmini_a7y0[pass]) continue;
				unsigned pw = (in->w - pngmini_a7x0[pass] + pngmini_a7dx[pass] - 1) / pngmini_a7dx[pass];
				unsigned ph = (in->h - pngmini_a7y0[pass] + pngmini_a7dy[pass] - 1) / pngmini_a7dy[pass];
				size_t rb = pngmini_rowbytes(pw, in->depth, ns);
				ok = pngmini_unfilter(p, rb, ph, bpp);
				for (unsigned y = 0; ok && y < ph; y++) {
					size_t oy = pngmini_a7y0[pass] + (size_t)y * pngmini_a7dy[pass];
					pngmini_expand_row(in, flags, p + (rb + 1) * (size_t)y + 1, pw, out + (oy * in->w + pngmini_a7x0[pass]) * nc,
					                   pngmini_a7dx[pass], nc);
				}
				p += (rb + 1) * (size_t)ph;
			}
		}
	}
	free(raw);
	if (!ok) { free(out); return NULL; }
	return out;
}

/*
 * Whole file in memory -> info (+ pixels if `pixels` is not NULL). 1 on success. With pixels == NULL only
 * the chunks before the first IDAT are looked at (what stbi_info needs).
 */
static int pngmini_decode(const unsigned char *data, size_t size, pngmini_info *in, int flags, unsigned char **pixels) {
	if (size < 8 + 25 || memcmp(data, pngmini_sig, 8) != 0) return 0;
	size_t pos = 8, idat_total = 0;
	int seen_ihdr = 0, seen_iend = 0;
	unsigned char *idat = NULL;
	for (int sweep = 0; sweep < 2; sweep++) { /* sweep 0: headers and IDAT size, sweep 1: gather IDAT */
		pos = 8;
		size_t got = 0;
		while (pos + 12 <= size) {
			unsigned len = pngmini_be32(data + pos);
			const unsigned char *type = data + pos + 4, *body = data + pos + 8;
			if (len > 0x7fffffffu || (size_t)len + 12 > size - pos) { free(idat); return 0; }
			if (sweep == 0) {
				if (crc32(crc32(0L, Z_NULL, 0), type, (uInt)len + 4) != pngmini_be32(body + len)) return 0;
				if (!seen_ihdr) {
					if (memcmp(type, "IHDR", 4) != 0 || len != 13 || !pngmini_parse_ihdr(body, in)) return 0;
					seen_ihdr = 1;
				} else if (memcmp(type, "PLTE", 4) == 0) {
					if (!pngmini_parse_plte(body, len, in)) return 0;
				} else if (memcmp(type, "tRNS", 4) == 0) {
					if (!pngmini_parse_trns(body, len, in)) return 0;
				} else if (memcmp(type, "IDAT", 4) == 0) {
					if (!pixels) return 1;
					idat_total += len;
				} else if (memcmp(type, "IEND", 4) == 0) {
					seen_iend = 1;
				} else if (!(type[0] & 32)) {
					return 0; /* unknown critical chunk */
				}
			} else if (memcmp(type, "IDAT", 4) == 0) {
				memcpy(idat + got, body, len);
				got += len;
			}
			pos += (size_t)len + 12;
			if (memcmp(type, "IEND", 4) == 0) break;
		}
		if (sweep == 0) {
			if (!seen_ihdr) return 0;
			if (!pixels) return 1;
			if (!seen_iend || idat_total == 0) return 0;
			idat = (unsigned char *)malloc(idat_total);
			if (!idat) return 0;
		}
	}
	*pixels = pngmini_decode_idat(in, idat, idat_total, flags);
	free(idat);
	return *pixels != NULL;
}

/* channel-count conversion with stb_image's conventions (grey = (77 r + 150 g + 29 b) >> 8, new alpha = 255) */
static unsigned char *pngmini_convert(unsigned char *src, int nc, int want, unsigned w, unsigned h) {
	if (nc == want) return src;
	size_t n = (size_t)w * h;
	unsigned char *dst = (unsigned char *)malloc(n * want);
	if (!dst) { free(src); return NULL; }
	for (size_t i = 0; i < n; i++) {
		const unsigned char *s = src + i * nc;
		unsigned char *d = dst + i * want;
		unsigned char r, g, b, a = 255;
		if (nc <= 2) { r = g = b = s[0]; if (nc == 2) a = s[1]; }
		else { r = s[0]; g = s[1]; b = s[2]; if (nc == 4) a = s[3]; }
		if (want <= 2) { d[0] = nc <= 2 ? r : (unsigned char)((r * 77 + g * 150 + b * 29) >> 8); if (want == 2) d[1] = a; }
		else { d[0] = r; d[1] = g; d[2] = b; if (want == 4) d[3] = a; }
	}
	free(src);
	return dst;
}

/* ------------------------------------------------------------------ encoder */

typedef void (*pngmini_sink)(void *ctx, const void *data, size_t len);

static void pngmini_emit_chunk(pngmini_sink sink, void *ctx, const char *type, const unsigned char *body, size_t len) {
	unsigned char hd[8], tail[4];
	pngmini_put32(hd, (unsigned)len);
	memcpy(hd + 4, type, 4);
	uLong c = crc32(crc32(0L, Z_NULL, 0), hd + 4, 4);
	if (len) c = crc32(c, body, (uInt)len);
	pngmini_put32(tail, (unsigned)c);
	sink(ctx, hd, 8);
	if (len) sink(ctx, body, len);
	sink(ctx, tail, 4);
}

/* 8-bit image with comp = 1 (grey), 2 (grey+alpha), 3 (RGB) or 4 (RGBA) channels -> PNG through `sink`. 1 on success */
static int pngmini_encode(pngmini_sink sink, void *ctx, const unsigned char *const *rows, unsigned w, unsigned h, int comp, int level) {
	static const unsigned char ctype_of[5] = {0, 0, 4, 2, 6};
	if (comp < 1 || comp > 4 || w == 0 || h == 0) return 0;
	size_t rb = (size_t)w * comp;
	unsigned char *filt = (unsigned char *)malloc((rb + 1) * h), *cand = (unsigned char *)malloc(rb);
	if (!filt || !cand) { free(filt); free(cand); return 0; }
	for (unsigned y = 0; y < h; y++) {
		const unsigned char *c = rows[y], *prev = y ? rows[y - 1] : NULL;
		unsigned char *dst = filt + (rb + 1) * (size_t)y;
		long best = -1;
		for (int f = 0; f < 5; f++) { /* smallest sum of |signed residual| wins */
			long sum = 0;
			for (size_t i = 0; i < rb; i++) {
				int a = i >= (size_t)comp ? c[i - comp] : 0, b = prev ? prev[i] : 0, cc = (prev && i >= (size_t)comp) ? prev[i - comp] : 0;
				int pred = f == 0 ? 0 : f == 1 ? a : f == 2 ? b : f == 3 ? ((a + b) >> 1) : pngmini_paeth(a, b, cc);
				cand[i] = (unsigned char)(c[i] - pred);
				sum += abs((signed char)cand[i]);
			}
			if (best < 0 || sum < best) { best = sum; dst[0] = (unsigned char)f; memcpy(dst + 1, cand, rb); }
		}
	}
	free(cand);
	uLong bound = compressBound((uLong)((rb + 1) * h));
	unsigned char *zbuf = (unsigned char *)malloc(bound);
	uLongf zlen = bound;
	int ok = zbuf && compress2(zbuf, &zlen, filt, (uLong)((rb + 1) * h), level) == Z_OK;
	free(filt);
	if (ok) {
		unsigned char ihdr[13];
		pngmini_put32(ihdr, w); pngmini_put32(ihdr + 4, h);
		ihdr[8] = 8; ihdr[9] = ctype_of[comp]; ihdr[10] = ihdr[11] = ihdr[12] = 0;
		sink(ctx, pngmini_sig, 8);
		pngmini_emit_chunk(sink, ctx, "IHDR", ihdr, 13);
		for (size_t o = 0; o < zlen; o += 1u << 20) /* IDAT in 1 MiB chunks */
			pngmini_emit_chunk(sink, ctx, "IDAT", zbuf + o, zlen - o < (1u << 20) ? zlen - o : (1u << 20));
		pngmini_emit_chunk(sink, ctx, "IEND", NULL, 0);
	}
	free(zbuf);
	return ok;
}

#endif /* PNGMINI_H */
