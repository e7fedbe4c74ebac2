/*
 * qoiconv_dir.c -- qoiconv for directories on the batch API.
 *
 * The reference's converter (qoiconv.c:46-85) handles one file per process: stbi_load -> qoi_write, or
 * qoi_read -> stbi_write_png.  This tool does the same conversions for every file of a directory, but
 * feeds the GPU whole batches through qoi_gpu_encode_host / qoi_gpu_decode_host (include/qoi_gpu.h) and
 * overlaps the three stages on a ring of two batch slots:
 *
 *     reader thread     PNG decode (OpenMP over the files of a batch) or .qoi file reads, straight into
 *                       the slot's pinned buffers
 *     main thread       plan + one host-buffer call per batch (H2D, kernels and D2H are pipelined inside it)
 *     writer thread     .qoi file writes / PNG encode (OpenMP) of the previous batch
 *
 *   qoiconv_dir <in_dir> <out_dir> [--budget-mpx N] [--verify]
 *       *.png in in_dir  ->  out_dir/<name>.qoi     (channels as qoiconv.c:51-54: 3 stays 3, everything else 4)
 *       *.qoi in in_dir  ->  out_dir/<name>.png     (channels of the header, qoiconv.c:60-63)
 *   --budget-mpx  pixels per batch in millions (default 256)
 *   --verify      after encoding, decode the batch on the GPU again and compare with the input pixels
 *                 (the round-trip check of qoibench.c:410-417)
 *
 * There is no CPU codec in here: without a CUDA device qoi_gpu_ctx_create fails and the tool exits with 2.
 * PNG I/O comes from stb_image / stb_image_write (the stand-ins of dropin/compat/ in this image).
 */
#define _POSIX_C_SOURCE 200809L
#include <dirent.h>
#include <pthread.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

#define STB_IMAGE_IMPLEMENTATION
#define STBI_ONLY_PNG
#define STBI_NO_LINEAR
#include "stb_image.h"
#define STB_IMAGE_WRITE_IMPLEMENTATION
#include "stb_image_write.h"

#include "qoi_gpu.h"

typedef struct {
	char *in_path, *out_path;
	unsigned w, h;
	int channels; /* of the pixels: 3 or 4 */
	int colorspace;
	int size;     /* .qoi input: file size */
	int failed;
} item_t;

typedef struct {
	int first, count;          /* items[first .. first+count) */
	int channels, to_qoi;
	uint64_t *pix_off, *str_off;
	int32_t *lens;
	size_t pix_bytes, str_bytes;
	uint8_t *pixels, *streams; /* pinned */
	size_t pix_cap, str_cap;
	int state;                 /* 0 free, 1 loaded, 2 converted */
	int last;                  /* no more batches after this one */
} slot_t;

static item_t *items;
static slot_t slots[2];
static pthread_mutex_t mu = PTHREAD_MUTEX_INITIALIZER;
static pthread_cond_t cv = PTHREAD_COND_INITIALIZER;
static int n_batches, opt_verify;
static int *batch_first, *batch_count;
static long failures;

static double now_s(void) {
	struct timespec ts;
	clock_gettime(CLOCK_MONOTONIC, &ts);
	return (double)ts.tv_sec + 1e-9 * (double)ts.tv_nsec;
}

static int ends_with(const char *s, const char *e) {
	size_t a = strlen(s), b = strlen(e);
	return a >= b && strcmp(s + a - b, e) == 0;
}

static void wait_state(slot_t *s, int want) {
	pthread_mutex_lock(&mu);
	while (s->state != want) pthread_cond_wait(&cv, &mu);
	pthread_mutex_unlock(&mu);
}

static void set_state(slot_t *s, int st) {
	pthread_mutex_lock(&mu);
	s->state = st;
	pthread_cond_broadcast(&cv);
	pthread_mutex_unlock(&mu);
}

static int ensure(uint8_t **buf, size_t *cap, size_t need) {
	if (need <= *cap) return 1;
	if (*buf) qoi_gpu_host_free(*buf);
	*cap = need + need / 4 + 4096;
	*buf = (uint8_t *)qoi_gpu_host_alloc(*cap);
	if (!*buf) *cap = 0;
	return *buf != NULL;
}

/* geometry of batch b in slot s: offsets 256-byte aligned; stream slots hold the worst case (qoi.h:374-376) */
static int layout_batch(slot_t *s, int b) {
	s->first = batch_first[b];
	s->count = batch_count[b];
	s->channels = items[s->first].channels;
	s->to_qoi = items[s->first].size == 0;
	s->last = b == n_batches - 1;
	s->pix_off = (uint64_t *)realloc(s->pix_off, sizeof(uint64_t) * (size_t)s->count);
	s->str_off = (uint64_t *)realloc(s->str_off, sizeof(uint64_t) * (size_t)s->count);
	s->lens = (int32_t *)realloc(s->lens, sizeof(int32_t) * (size_t)s->count);
	size_t po = 0, so = 0;
	for (int i = 0; i < s->count; i++) {
		const item_t *it = &items[s->first + i];
		qoi_desc d = {it->w, it->h, (unsigned char)it->channels, QOI_SRGB};
		s->pix_off[i] = po;
		s->str_off[i] = so;
		po += ((size_t)it->w * it->h * (size_t)it->channels + 255) & ~(size_t)255;
		/* a file may hold RGBA ops under a 3-channel header (qoi.h:552-557 decodes them): 5 bytes per pixel at most */
		so += ((s->to_qoi ? qoi_gpu_max_stream_size(&d) : (size_t)it->w * it->h * 5u + 22u) + 255) & ~(size_t)255;
	}
	s->pix_bytes = po;
	s->str_bytes = so;
	return ensure(&s->pixels, &s->pix_cap, po + 64) && ensure(&s->streams, &s->str_cap, so + 64);
}

static void *reader_main(void *arg) {
	(void)arg;
	for (int b = 0; b < n_batches; b++) {
		slot_t *s = &slots[b & 1];
		wait_state(s, 0);
		if (!layout_batch(s, b)) {
			fprintf(stderr, "qoiconv_dir: out of pinned memory for batch %d\n", b);
			exit(2);
		}
#pragma omp parallel for schedule(dynamic)
		for (int i = 0; i < s->count; i++) {
			item_t *it = &items[s->first + i];
			if (s->to_qoi) { /* qoiconv.c:56 */
				int w, h;
				unsigned char *px = stbi_load(it->in_path, &w, &h, NULL, it->channels);
				if (!px || (unsigned)w != it->w || (unsigned)h != it->h) it->failed = 1;
				else memcpy(s->pixels + s->pix_off[i], px, (size_t)w * h * (size_t)it->channels);
				stbi_image_free(px);
			} else { /* the file bytes are the stream */
				FILE *f = fopen(it->in_path, "rb");
				if (!f || fread(s->streams + s->str_off[i], 1, (size_t)it->size, f) != (size_t)it->size) it->failed = 1;
				if (f) fclose(f);
				s->lens[i] = it->size;
			}
		}
		set_state(s, 1);
	}
	return NULL;
}

static void *writer_main(void *arg) {
	(void)arg;
	for (int b = 0; b < n_batches; b++) {
		slot_t *s = &slots[b & 1];
		wait_state(s, 2);
		long bad = 0;
#pragma omp parallel for schedule(dynamic) reduction(+ : bad)
		for (int i = 0; i < s->count; i++) {
			item_t *it = &items[s->first + i];
			int ok = !it->failed;
			if (ok && s->to_qoi) { /* what qoi_write does with the encoder's result (qoi.h:595-617) */
				FILE *f = fopen(it->out_path, "wb");
				ok = f && s->lens[i] > 0 && fwrite(s->streams + s->str_off[i], 1, (size_t)s->lens[i], f) == (size_t)s->lens[i];
				if (f && fclose(f) != 0) ok = 0;
			} else if (ok) { /* qoiconv.c:73 */
				ok = stbi_write_png(it->out_path, (int)it->w, (int)it->h, it->channels, s->pixels + s->pix_off[i], 0) != 0;
			}
			if (!ok) {
				fprintf(stderr, "qoiconv_dir: failed: %s\n", it->in_path);
				bad++;
			}
		}
		failures += bad;
		set_state(s, 0);
	}
	return NULL;
}

static int cmp_items(const void *a, const void *b) {
	const item_t *x = (const item_t *)a, *y = (const item_t *)b;
	int kx = (x->size ? 8 : 0) + x->channels, ky = (y->size ? 8 : 0) + y->channels; /* direction, then channel count */
	if (kx != ky) return kx - ky;
	return strcmp(x->in_path, y->in_path);
}

int main(int argc, char **argv) {
	if (argc < 3) {
		puts("Usage: qoiconv_dir <in_dir> <out_dir> [--budget-mpx N] [--verify]");
		puts("  converts every *.png of in_dir to out_dir/*.qoi and every *.qoi to out_dir/*.png on the GPU");
		return 1;
	}
	double budget = 256e6;
	for (int i = 3; i < argc; i++) {
		if (strcmp(argv[i], "--verify") == 0) opt_verify = 1;
		else if (strcmp(argv[i], "--budget-mpx") == 0 && i + 1 < argc) budget = atof(argv[++i]) * 1e6;
		else { fprintf(stderr, "Unknown option %s\n", argv[i]); return 1; }
	}
	DIR *dir = opendir(argv[1]);
	if (!dir) { fprintf(stderr, "Couldn't open directory %s\n", argv[1]); return 1; }
	int n = 0, cap = 0;
	struct dirent *e;
	while ((e = readdir(dir)) != NULL) {
		const int png = ends_with(e->d_name, ".png"), qoi = ends_with(e->d_name, ".qoi");
		if (!png && !qoi) continue;
		if (n == cap) {
			cap = cap ? cap * 2 : 64;
			items = (item_t *)realloc(items, sizeof(item_t) * (size_t)cap);
		}
		item_t *it = &items[n];
		memset(it, 0, sizeof *it);
		size_t li = strlen(argv[1]) + strlen(e->d_name) + 2, lo = strlen(argv[2]) + strlen(e->d_name) + 2;
		it->in_path = (char *)malloc(li);
		it->out_path = (char *)malloc(lo);
		snprintf(it->in_path, li, "%s/%s", argv[1], e->d_name);
		snprintf(it->out_path, lo, "%s/%s", argv[2], e->d_name);
		memcpy(it->out_path + strlen(it->out_path) - 3, png ? "qoi" : "png", 3);
		int ok = 0;
		if (png) { /* qoiconv.c:46-54 */
			int w, h, ch;
			if (stbi_info(it->in_path, &w, &h, &ch)) {
				it->w = (unsigned)w; it->h = (unsigned)h; it->channels = ch == 3 ? 3 : 4;
				ok = 1;
			}
		} else { /* header fields as qoi.h:507-521 reads them; the device re-checks them against the plan */
			unsigned char hd[14];
			FILE *f = fopen(it->in_path, "rb");
			if (f && fread(hd, 1, 14, f) == 14 && memcmp(hd, "qoif", 4) == 0) {
				it->w = ((unsigned)hd[4] << 24) | ((unsigned)hd[5] << 16) | ((unsigned)hd[6] << 8) | hd[7];
				it->h = ((unsigned)hd[8] << 24) | ((unsigned)hd[9] << 16) | ((unsigned)hd[10] << 8) | hd[11];
				it->channels = hd[12];
				fseek(f, 0, SEEK_END);
				long sz = ftell(f);
				it->colorspace = hd[13];
				ok = (it->channels == 3 || it->channels == 4) && hd[13] <= 1 && it->w && it->h && it->h < 400000000u / it->w &&
				     sz >= 22 && (uint64_t)sz <= (uint64_t)it->w * it->h * 5u + 22u;
				it->size = (int)sz;
			}
			if (f) fclose(f);
		}
		if (!ok) {
			fprintf(stderr, "qoiconv_dir: skipping %s (header not usable)\n", it->in_path);
			failures++;
			continue;
		}
		n++;
	}
	closedir(dir);
	if (n == 0) { printf("No images found in %s\n", argv[1]); return failures ? 1 : 0; }
	qsort(items, (size_t)n, sizeof(item_t), cmp_items);
	/* batches: same direction and channel count, at most `budget` pixels (one oversized image is its own batch) */
	batch_first = (int *)malloc(sizeof(int) * (size_t)n);
	batch_count = (int *)malloc(sizeof(int) * (size_t)n);
	double acc = 0;
	for (int i = 0; i < n; i++) {
		const double px = (double)items[i].w * items[i].h;
		const int same = n_batches && items[i].channels == items[i - 1].channels && !items[i].size == !items[i - 1].size;
		if (!same || acc + px > budget) {
			batch_first[n_batches] = i;
			batch_count[n_batches++] = 0;
			acc = 0;
		}
		batch_count[n_batches - 1]++;
		acc += px;
	}

	qoi_gpu_ctx *ctx = NULL;
	if (qoi_gpu_ctx_create(&ctx, 0) != QOI_GPU_OK) {
		fprintf(stderr, "qoiconv_dir: no usable CUDA device (there is no CPU path)\n");
		return 2;
	}
	const double t0 = now_s();
	pthread_t reader, writer;
	pthread_create(&reader, NULL, reader_main, NULL);
	pthread_create(&writer, NULL, writer_main, NULL);
	double gpu_s = 0, total_px = 0;
	uint8_t *check = NULL;
	size_t check_cap = 0;
	for (int b = 0; b < n_batches; b++) {
		slot_t *s = &slots[b & 1];
		wait_state(s, 1);
		qoi_desc *descs = (qoi_desc *)malloc(sizeof(qoi_desc) * (size_t)s->count);
		for (int i = 0; i < s->count; i++) {
			const item_t *it = &items[s->first + i];
			descs[i] = (qoi_desc){it->w, it->h, (unsigned char)it->channels, (unsigned char)it->colorspace}; /* qoiconv.c:80: QOI_SRGB when encoding */
			total_px += (double)it->w * it->h;
		}
		qoi_gpu_plan *plan = NULL;
		int rc = qoi_gpu_plan_create(ctx, &plan, s->count, descs, s->pix_off, s->str_off, s->pix_off, s->channels);
		const double g0 = now_s();
		if (rc == QOI_GPU_OK) {
			if (s->to_qoi) {
				rc = qoi_gpu_encode_host(plan, s->pixels, s->streams, s->lens);
				if (rc == QOI_GPU_OK && opt_verify) { /* qoibench.c:410-417 on the whole batch */
					if (!ensure(&check, &check_cap, s->pix_bytes + 64)) rc = QOI_GPU_E_NOMEM;
					else rc = qoi_gpu_decode_host(plan, s->streams, s->lens, check, 1);
					for (int i = 0; rc == QOI_GPU_OK && i < s->count; i++) {
						const item_t *it = &items[s->first + i];
						if (!it->failed && memcmp(check + s->pix_off[i], s->pixels + s->pix_off[i], (size_t)it->w * it->h * (size_t)it->channels) != 0) {
							fprintf(stderr, "qoiconv_dir: QOI roundtrip pixel mismatch for %s\n", it->in_path);
							items[s->first + i].failed = 1;
						}
					}
				}
			} else {
				rc = qoi_gpu_decode_host(plan, s->streams, s->lens, s->pixels, 0); /* files are foreign streams */
			}
		}
		gpu_s += now_s() - g0;
		if (rc != QOI_GPU_OK) {
			fprintf(stderr, "qoiconv_dir: batch %d failed (%d): %s\n", b, rc, qoi_gpu_last_error(ctx));
			for (int i = 0; i < s->count; i++) items[s->first + i].failed = 1;
		}
		if (plan) qoi_gpu_plan_destroy(plan);
		free(descs);
		set_state(s, 2);
	}
	pthread_join(reader, NULL);
	pthread_join(writer, NULL);
	const double dt = now_s() - t0;
	printf("qoiconv_dir: %d files in %d batches, %.1f Mpx, %.3f s wall (%.1f Mpx/s), %.3f s inside the GPU calls, %ld failed\n", n, n_batches,
	       total_px / 1e6, dt, total_px / 1e6 / dt, gpu_s, failures);
	if (check) qoi_gpu_host_free(check);
	for (int k = 0; k < 2; k++) {
		if (slots[k].pixels) qoi_gpu_host_free(slots[k].pixels);
		if (slots[k].streams) qoi_gpu_host_free(slots[k].streams);
	}
	qoi_gpu_ctx_destroy(ctx);
	return failures ? 1 : 0;
}
