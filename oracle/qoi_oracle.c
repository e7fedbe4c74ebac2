/*
 * qoi_oracle.c -- CPU restatement of the QOI encode / decode algorithm.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing in the product path (cuda/, qoi_b200/,
 * include/) may call, link or load this file.  It is used by tests/, by
 * __graft_entry__.smoke() and by bench.py's cpu_baseline / --impl reference
 * legs as the checker and the CPU baseline.
 *
 * Parity is PINNED: tests/test_oracle.py checks this restatement byte-for-byte
 * against (a) the known-answer vectors of SURVEY.md Appendix A.4 (committed as
 * tests/golden/kat.json, produced by the reference) and (b) the unmodified
 * reference compiled into oracle/_ref/libqoi_ref.so, on thousands of seeded
 * images and hand-made foreign streams.
 *
 * The code below is written from the format rules, not transcribed from the
 * reference; each function cites the reference lines whose behaviour it
 * restates (paths relative to /root/reference).
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define ORA_HEADER 14
#define ORA_TAIL 8
#define ORA_PIXEL_LIMIT 400000000u /* qoi.h:332 */

typedef struct {
	uint32_t width, height;
	uint8_t channels, colorspace;
} ora_desc; /* same 12-byte layout as qoi_desc, qoi.h:236-241 */

typedef struct {
	uint8_t r, g, b, a;
} ora_px;

static inline int ora_same(ora_px x, ora_px y) {
	return x.r == y.r && x.g == y.g && x.b == y.b && x.a == y.a;
}

/* qoi.h:322 (QOI_COLOR_HASH) with the & 63 of qoi.h:430,577 folded in */
static inline unsigned ora_slot(ora_px c) {
	return (c.r * 3u + c.g * 5u + c.b * 7u + c.a * 11u) & 63u;
}

static void ora_put_be32(uint8_t *dst, uint32_t v) { /* qoi.h:341-346 */
	dst[0] = (uint8_t)(v >> 24);
	dst[1] = (uint8_t)(v >> 16);
	dst[2] = (uint8_t)(v >> 8);
	dst[3] = (uint8_t)v;
}

static uint32_t ora_get_be32(const uint8_t *src) { /* qoi.h:348-354 */
	return ((uint32_t)src[0] << 24) | ((uint32_t)src[1] << 16) | ((uint32_t)src[2] << 8) | src[3];
}

/* the geometry / argument checks shared by both directions: qoi.h:366-369, 514-518 */
static int ora_geometry_ok(uint32_t w, uint32_t h, unsigned ch, unsigned cs) {
	if (w == 0 || h == 0) return 0;
	if (ch < 3 || ch > 4) return 0;
	if (cs > 1) return 0;
	if (h >= ORA_PIXEL_LIMIT / w) return 0;
	return 1;
}

size_t ora_max_encoded_size(const ora_desc *d) { /* qoi.h:374-376 */
	return (size_t)d->width * d->height * (d->channels + 1u) + ORA_HEADER + ORA_TAIL;
}

/*
 * Encoder.  Restates qoi.h:356-486.  Returns a malloc'd stream (free()) or NULL.
 * Differences in form, not behaviour: the output is produced by a small
 * emitter; the run counter is flushed by one helper.
 */
void *ora_encode(const void *pixels, const ora_desc *d, int *out_len) {
	if (!pixels || !d || !out_len) return NULL; /* qoi.h:365 */
	if (!ora_geometry_ok(d->width, d->height, d->channels, d->colorspace)) return NULL;

	const uint8_t *src = (const uint8_t *)pixels;
	const unsigned ch = d->channels;
	const size_t npx = (size_t)d->width * d->height;
	uint8_t *out = (uint8_t *)malloc(ora_max_encoded_size(d));
	if (!out) return NULL;

	size_t w = 0;
	out[w++] = 'q'; out[w++] = 'o'; out[w++] = 'i'; out[w++] = 'f'; /* qoi.h:323-325,384 */
	ora_put_be32(out + w, d->width);  w += 4;
	ora_put_be32(out + w, d->height); w += 4;
	out[w++] = d->channels;
	out[w++] = d->colorspace;

	ora_px table[64];
	memset(table, 0, sizeof table);            /* qoi.h:393 */
	ora_px before = {0, 0, 0, 255};            /* qoi.h:396-399 */
	unsigned pending = 0;                      /* run length not yet written */

	for (size_t i = 0; i < npx; i++) {
		ora_px cur = before;                   /* alpha persists for 3-channel input, qoi.h:400,411-413 */
		cur.r = src[i * ch + 0];
		cur.g = src[i * ch + 1];
		cur.b = src[i * ch + 2];
		if (ch == 4) cur.a = src[i * ch + 3];

		if (ora_same(cur, before)) {           /* qoi.h:415-421 */
			pending++;
			if (pending == 62 || i + 1 == npx) {
				out[w++] = (uint8_t)(0xC0u | (pending - 1));
				pending = 0;
			}
			continue;                          /* table untouched, prev unchanged */
		}
		if (pending) {                         /* qoi.h:425-428 */
			out[w++] = (uint8_t)(0xC0u | (pending - 1));
			pending = 0;
		}
		unsigned s = ora_slot(cur);
		if (ora_same(table[s], cur)) {         /* qoi.h:432-434 */
			out[w++] = (uint8_t)s;
		} else {
			table[s] = cur;                    /* qoi.h:436 */
			if (cur.a != before.a) {           /* qoi.h:468-474 */
				out[w++] = 0xFF;
				out[w++] = cur.r; out[w++] = cur.g; out[w++] = cur.b; out[w++] = cur.a;
			} else {
				int8_t dr = (int8_t)(cur.r - before.r);   /* qoi.h:439-441, wraps mod 256 */
				int8_t dg = (int8_t)(cur.g - before.g);
				int8_t db = (int8_t)(cur.b - before.b);
				int8_t dr_g = (int8_t)(dr - dg);          /* qoi.h:443-444 */
				int8_t db_g = (int8_t)(db - dg);
				if (dr >= -2 && dr <= 1 && dg >= -2 && dg <= 1 && db >= -2 && db <= 1) {
					out[w++] = (uint8_t)(0x40 | ((dr + 2) << 4) | ((dg + 2) << 2) | (db + 2)); /* qoi.h:451 */
				} else if (dg >= -32 && dg <= 31 && dr_g >= -8 && dr_g <= 7 && db_g >= -8 && db_g <= 7) {
					out[w++] = (uint8_t)(0x80 | (dg + 32));                /* qoi.h:458-459 */
					out[w++] = (uint8_t)(((dr_g + 8) << 4) | (db_g + 8));
				} else {
					out[w++] = 0xFE;                                       /* qoi.h:462-465 */
					out[w++] = cur.r; out[w++] = cur.g; out[w++] = cur.b;
				}
			}
		}
		before = cur;                          /* qoi.h:477 */
	}

	memset(out + w, 0, 7); w += 7;             /* qoi.h:339,480-482 */
	out[w++] = 1;
	*out_len = (int)w;
	return out;
}

/*
 * Decoder.  Restates qoi.h:488-590, including: desc filled before header
 * validation (508-511 precede 513-521); channels==0 means "as in header";
 * chunk parsing stops 8 bytes before the end and the remaining pixels repeat
 * the last colour (544); every parsed op re-inserts the current colour (577).
 */
void *ora_decode(const void *data, int size, ora_desc *d, int channels) {
	if (!data || !d) return NULL;
	if (channels != 0 && channels != 3 && channels != 4) return NULL;
	if (size < ORA_HEADER + ORA_TAIL) return NULL;

	const uint8_t *in = (const uint8_t *)data;
	uint32_t magic = ora_get_be32(in);
	d->width = ora_get_be32(in + 4);
	d->height = ora_get_be32(in + 8);
	d->channels = in[12];
	d->colorspace = in[13];
	if (magic != 0x716F6966u) return NULL; /* "qoif" */
	if (!ora_geometry_ok(d->width, d->height, d->channels, d->colorspace)) return NULL;

	const unsigned och = channels ? (unsigned)channels : d->channels;
	const size_t npx = (size_t)d->width * d->height;
	uint8_t *dst = (uint8_t *)malloc(npx * och);
	if (!dst) return NULL;

	ora_px table[64];
	memset(table, 0, sizeof table);
	ora_px cur = {0, 0, 0, 255};
	const long stop = (long)size - ORA_TAIL; /* qoi.h:539 */
	long rd = ORA_HEADER;
	unsigned repeat = 0;

	for (size_t i = 0; i < npx; i++) {
		if (repeat) {
			repeat--;
		} else if (rd < stop) {
			unsigned tag = in[rd++];
			if (tag == 0xFE) {
				cur.r = in[rd]; cur.g = in[rd + 1]; cur.b = in[rd + 2]; rd += 3;
			} else if (tag == 0xFF) {
				cur.r = in[rd]; cur.g = in[rd + 1]; cur.b = in[rd + 2]; cur.a = in[rd + 3]; rd += 4;
			} else {
				switch (tag >> 6) {
				case 0: cur = table[tag]; break;
				case 1:
					cur.r = (uint8_t)(cur.r + ((tag >> 4) & 3) - 2);
					cur.g = (uint8_t)(cur.g + ((tag >> 2) & 3) - 2);
					cur.b = (uint8_t)(cur.b + (tag & 3) - 2);
					break;
				case 2: {
					unsigned second = in[rd++];
					int dg = (int)(tag & 63) - 32;
					cur.r = (uint8_t)(cur.r + dg - 8 + (int)(second >> 4));
					cur.g = (uint8_t)(cur.g + dg);
					cur.b = (uint8_t)(cur.b + dg - 8 + (int)(second & 15));
					break;
				}
				default: repeat = tag & 63; break;
				}
			}
			table[ora_slot(cur)] = cur;
		}
		uint8_t *o = dst + i * och;
		o[0] = cur.r; o[1] = cur.g; o[2] = cur.b;
		if (och == 4) o[3] = cur.a;
	}
	return dst;
}

void ora_free(void *p) { free(p); }
