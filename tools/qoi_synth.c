/*
 * qoi_synth.c -- deterministic synthetic image generator.
 *
 * Inputs for the tests and for bench.py (both arms read the same buffers).
 * Not part of the product library and not part of the oracle.
 *
 * One xorshift32 draw per pixel, in raster order; seed =
 * 0x9E3779B9 ^ dist ^ (image_index * 2654435761u), forced to 1 if it is 0.
 *
 *   0 noise_rgba    r, r>>8, r>>16, r>>24             worst case (RGBA ops)
 *   1 noise_opaque  same, a = 255                      RGB ops
 *   2 gradient      x*255/w, y*255/h, (x+y)*255/(w+h)  RUN + DIFF
 *   3 grad2         gradient + 2-bit noise per channel INDEX/DIFF/LUMA
 *   4 photo         gradient + 4-bit noise per channel LUMA/RGB  (headline)
 *   5 tiles32       32x32 flat tiles, two alpha values RUN + RGBA/INDEX
 *   6 palette16     16 colours picked by noise         INDEX
 *   7 const0        all bytes zero                     RUN
 */
#include <stdint.h>
#include <stddef.h>

static inline uint32_t xs32(uint32_t *s) {
	uint32_t x = *s;
	x ^= x << 13;
	x ^= x >> 17;
	x ^= x << 5;
	*s = x;
	return x;
}

void qoi_synth_image(uint8_t *dst, uint32_t w, uint32_t h, int channels, int dist, uint32_t image_index) {
	uint32_t s = 0x9E3779B9u ^ (uint32_t)dist ^ (image_index * 2654435761u);
	if (s == 0) s = 1;
	for (uint32_t y = 0; y < h; y++) {
		for (uint32_t x = 0; x < w; x++) {
			uint32_t r = xs32(&s);
			uint8_t R, G, B, A = 255;
			uint32_t gx = x * 255u / w, gy = y * 255u / h, gxy = (x + y) * 255u / (w + h);
			switch (dist) {
			case 0: R = (uint8_t)r; G = (uint8_t)(r >> 8); B = (uint8_t)(r >> 16); A = (uint8_t)(r >> 24); break;
			case 1: R = (uint8_t)r; G = (uint8_t)(r >> 8); B = (uint8_t)(r >> 16); break;
			case 2: R = (uint8_t)gx; G = (uint8_t)gy; B = (uint8_t)gxy; break;
			case 3: R = (uint8_t)(gx + (r & 3)); G = (uint8_t)(gy + ((r >> 2) & 3)); B = (uint8_t)(gxy + ((r >> 4) & 3)); break;
			case 4: R = (uint8_t)(gx + (r & 15)); G = (uint8_t)(gy + ((r >> 4) & 15)); B = (uint8_t)(gxy + ((r >> 8) & 15)); break;
			case 5: {
				uint32_t k = ((x >> 5) + (y >> 5) * 7u) & 31u;
				R = (uint8_t)(k * 8); G = (uint8_t)(255 - k * 5); B = (uint8_t)(k * 3); A = (k & 1) ? 255 : 128;
				break;
			}
			case 6: {
				uint32_t k = r & 15u;
				R = (uint8_t)(k * 16); G = (uint8_t)(k * 9); B = (uint8_t)(255 - k * 13);
				break;
			}
			default: R = G = B = A = 0; break;
			}
			dst[0] = R; dst[1] = G; dst[2] = B;
			if (channels == 4) dst[3] = A;
			dst += channels;
		}
	}
}

/* A batch: image i uses image_index = first_index + i; images are laid out back to back with `stride` bytes each. */
void qoi_synth_batch(uint8_t *dst, size_t stride, int n, uint32_t w, uint32_t h, int channels, int dist, uint32_t first_index) {
#pragma omp parallel for schedule(dynamic)
	for (int i = 0; i < n; i++) {
		qoi_synth_image(dst + (size_t)i * stride, w, h, channels, dist, first_index + (uint32_t)i);
	}
}
