/*
 * emu_main.cpp -- runs the REAL kernel source (cuda/qoi_enc.cuh, cuda/qoi_dec.cuh) on the CPU
 * through tests/emu/cuda_emu.h and checks every byte against the oracle.
 *
 * TEST INFRASTRUCTURE ONLY (driven by tests/test_emu.py).  Not a product path: it is a separate
 * executable, built with g++ -DQOI_EMU, and is never loaded by qoi_b200 or libqoi_gpu.so.
 *
 *   emu_main enc <w> <h> <ch> <n> <grid> <seed> <misalign> <pixel file>
 *        pixel file: n images back to back (w*h*ch bytes each).  Encodes the batch twice
 *        (two epochs on the same state), compares with ora_encode, then decodes the result with
 *        and without the restart index to 3 and 4 channels and compares with ora_decode.
 *   emu_main dec <w> <h> <ch> <n> <seed> <stream file>
 *        stream file: n records [int32 length][bytes]; foreign streams, decoded without index.
 */
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <algorithm>
#include <string>
#include <vector>

#define QOI_EMU 1
#include "qoi_dec.cuh"
#include "qoi_enc.cuh"
#include "qoi_fdec.cuh"
#include "qoi_plan.h"

extern "C" {
void *ora_encode(const void *data, const QoiDescPOD *desc, int *out_len);
void *ora_decode(const void *data, int size, QoiDescPOD *desc, int channels);
}

static std::vector<unsigned char> read_file(const char *path) {
	FILE *f = fopen(path, "rb");
	if (!f) { fprintf(stderr, "cannot open %s\n", path); exit(2); }
	fseek(f, 0, SEEK_END);
	long n = ftell(f);
	fseek(f, 0, SEEK_SET);
	std::vector<unsigned char> v((size_t)n);
	if (n && fread(v.data(), 1, (size_t)n, f) != (size_t)n) { fprintf(stderr, "short read\n"); exit(2); }
	fclose(f);
	return v;
}

template <int CH, bool EXPAND>
static void run_encode(const QoiEncParams &P, unsigned grid, unsigned seed) {
	using L = qoienc::Layout<CH, EXPAND>;
	qemu::launch(grid, QOI_ENC_THREADS, (size_t)L::TOTAL, [&]() { qoienc::qoi_encode_kernel<CH, EXPAND>(P); }, seed);
}

static void run_decode(const QoiDecParams &P, unsigned seed) {
	unsigned grid = (P.n_jobs + QOI_DEC_THREADS - 1) / QOI_DEC_THREADS;
	size_t smem = QOI_DEC_SMEM_BYTES(QOI_DEC_THREADS);
	if (P.out_channels == 4) qemu::launch(grid, QOI_DEC_THREADS, smem, [&]() { qoidec::qoi_decode_kernel<4, QOI_DEC_THREADS>(P); }, seed);
	else qemu::launch(grid, QOI_DEC_THREADS, smem, [&]() { qoidec::qoi_decode_kernel<3, QOI_DEC_THREADS>(P); }, seed);
}

/* the foreign-stream pipeline exactly as qoi_gpu_decode enqueues it: build the restart index, then the tile decoder */
static void run_foreign(const QoiDecParams &D, const QoiPlanGeom &g, unsigned seed, int rounds = QOI_FROUNDS) {
	static_assert(QOI_FSEG_BYTES_HOST == QOI_FSEG_BYTES && QOI_FCHUNK_HOST == QOI_FCHUNK, "segment / chunk size");
	QoiFLayout FL = qoi_fstate_layout(g);
	std::vector<unsigned char> store(FL.total + 256, 0x5A);
	unsigned char *base = (unsigned char *)(((uintptr_t)store.data() + 255) & ~(uintptr_t)255);
	memset(base, 0, FL.total);
	memcpy(base + FL.off[0], FL.seg_base.data(), FL.seg_base.size() * 4);
	memcpy(base + FL.off[1], FL.chunk_base.data(), FL.chunk_base.size() * 4);
	/* the index goes where the plan keeps it; a caller without one (mode `dec`) gets a scratch copy */
	QoiStateLayout SL = qoi_state_layout(g.n_tiles);
	std::vector<unsigned char> idx_store;
	QoiTileState st = D.st;
	if (!st.ent) {
		idx_store.assign(SL.total + 256, 0x3C);
		qoi_state_bind(st, (unsigned char *)(((uintptr_t)idx_store.data() + 255) & ~(uintptr_t)255), SL);
	}
	qoifdec::QoiFParams F;
	memset(&F, 0, sizeof F);
	F.streams = D.streams; F.lens = D.lens; F.status = D.status; F.imgs = D.imgs;
	F.n_images = D.n_images; F.n_segs = FL.n_segs; F.n_chunks = FL.n_chunks;
	F.st = st;
	F.epoch = 1;
	qoifdec::fstate_bind(F.f, base, FL.off);
	const unsigned gs = (FL.n_segs + QOI_FDEC_THREADS - 1) / QOI_FDEC_THREADS;
	qemu::launch(gs, QOI_FDEC_THREADS, 0, [&]() { qoifdec::fparse_kernel(F); }, seed);
	qemu::launch(D.n_images, QOI_FCHAIN_THREADS, QOI_FCHAIN_SMEM, [&]() { qoifdec::fchain_kernel(F); }, seed);
	qemu::launch(gs, QOI_FDEC_THREADS, 0, [&]() { qoifdec::fzero_kernel(F); }, seed);
	for (int r = 0; r < rounds; r++) {
		F.round = (uint32_t)r;
		const int last = r + 1 == rounds;
		const unsigned gst = (FL.n_segs + QOI_FSTATE_THREADS - 1) / QOI_FSTATE_THREADS;
		qemu::launch(r < 3 ? gst : (gst < 5u ? gst : 5u), QOI_FSTATE_THREADS, QOI_FDEC_SMEM, [&]() { qoifdec::fstate_kernel(F); }, seed);
		qemu::launch(r < 3 ? FL.n_chunks : (FL.n_chunks < 7u ? FL.n_chunks : 7u), QOI_FSCAN_THREADS, QOI_FMETA_SMEM, [&]() { qoifdec::fscan_a_kernel(F); }, seed);
		qemu::launch(D.n_images, QOI_FSCANB_THREADS, QOI_FSCANB_SMEM, [&]() { qoifdec::fscan_b_kernel(F); }, seed);
		qemu::launch(r < 3 ? FL.n_chunks : (FL.n_chunks < 7u ? FL.n_chunks : 7u), QOI_FSCAN_THREADS, QOI_FVERIFY_SMEM, [&]() { qoifdec::fverify_kernel(F, last); }, seed);
		if (getenv("QOI_EMU_FTRACE2")) {
			for (uint32_t i = 0; i < D.n_images; i++) {
				fprintf(stderr, "round %d image %u flagged:", r, i);
				for (uint32_t j = 0; j < F.f.n_live[i]; j++) if (F.f.flag[F.f.seg_base[i] + j]) fprintf(stderr, " %u(%u)", j, F.f.flag[F.f.seg_base[i] + j]);
				fprintf(stderr, "\n");
			}
		}
		if (getenv("QOI_EMU_FTRACE3")) { /* the segments listed for the next round, with what they depend on */
			const uint32_t nl = F.f.n_list[r + 1];
			fprintf(stderr, "round %d: first_bad key %u, %u listed:", r, F.f.first_bad[((uint32_t)r & 1u) * D.n_images], nl);
			std::vector<uint32_t> l(F.f.list, F.f.list + nl);
			std::sort(l.begin(), l.end());
			for (uint32_t k = 0; k < nl && k < 24; k++)
				fprintf(stderr, " %u[dep %x rbw %llx gpx %08x]", l[k], F.f.dep[l[k]], (unsigned long long)F.f.rbw[l[k]], F.f.g_px[l[k]]);
			fprintf(stderr, "\n");
		}
		if (getenv("QOI_EMU_FWIN")) { /* "lo-hi": every segment of the window after this round's check */
			unsigned lo = 0, hi = 0;
			sscanf(getenv("QOI_EMU_FWIN"), "%u-%u", &lo, &hi);
			const uint32_t nl = F.f.n_list[r + 1];
			fprintf(stderr, "round %d window:\n", r);
			for (uint32_t g = lo; g <= hi && g < F.f.n_live[0]; g++) {
				bool listed = false;
				for (uint32_t k = 0; k < nl; k++) listed = listed || F.f.list[k] == g;
				fprintf(stderr, "  seg %u%s flag %u dep %02x rbw %016llx taint %016llx wr %016llx g_px %08x o_px %08x claim %x\n", g, listed ? " LISTED" : "", F.f.flag[g], F.f.dep[g],
				        (unsigned long long)F.f.rbw[g], (unsigned long long)F.f.o_taint[g], (unsigned long long)F.f.o_wr[g], F.f.g_px[g], F.f.o_px[g], F.f.claim[g]);
			}
		}
		if (getenv("QOI_EMU_FTRACE")) {
			for (uint32_t i = 0; i < D.n_images; i++)
				fprintf(stderr, "round %d image %u: %u live segments, %u to walk again, bad_dep %u\n", r, i, F.f.n_live[i], F.f.n_list[r + 1],
				        F.f.bad_cnt[((uint32_t)r & 1u) * D.n_images + i]);
		}
	}
	if (getenv("QOI_EMU_FSTATS"))
		for (uint32_t i = 0; i < D.n_images; i++)
			fprintf(stderr, "image %u: %u live segments, fall-back from segment %u after %d rounds\n", i, F.f.n_live[i], qoifdec::stream_first_bad(F, i), rounds);
	qemu::launch(FL.n_chunks, QOI_FSCAN_THREADS, QOI_FMETA_SMEM, [&]() { qoifdec::ffinal_kernel(F); }, seed);
	qemu::launch(D.n_images, 32, QOI_FFALL_SMEM, [&]() { qoifdec::ffall_kernel(F); }, seed);
	qemu::launch((g.n_tiles * 64u + 127u) / 128u, 128, 0, [&]() { qoifdec::fcompose_kernel(F, g.n_tiles, D.tile_img); }, seed);
	QoiDecParams I = D;
	I.use_index = 1;
	I.n_jobs = g.n_tiles;
	I.status = nullptr; /* fchain_kernel reported it */
	I.st = st;
	run_decode(I, seed);
}

static int first_diff(const unsigned char *a, const unsigned char *b, size_t n) {
	for (size_t i = 0; i < n; i++)
		if (a[i] != b[i]) return (int)i;
	return -1;
}

int main(int argc, char **argv) {
	if (argc < 2) return 2;
	std::string mode = argv[1];
	if (mode == "enc" && argc == 10) {
		unsigned w = atoi(argv[2]), h = atoi(argv[3]), ch = atoi(argv[4]);
		int n = atoi(argv[5]);
		unsigned grid = atoi(argv[6]), seed = atoi(argv[7]), mis = atoi(argv[8]);
		std::vector<unsigned char> file = read_file(argv[9]);
		size_t img_bytes = (size_t)w * h * ch;
		if (file.size() != img_bytes * n) { fprintf(stderr, "pixel file size mismatch\n"); return 2; }

		std::vector<QoiDescPOD> descs(n);
		std::vector<uint64_t> poff(n), soff(n);
		QoiDescPOD d0 = {w, h, (unsigned char)ch, 0};
		size_t slot = (qoi_max_stream(d0) + 37) | 1; /* odd slot size: every destination alignment occurs */
		for (int i = 0; i < n; i++) {
			descs[i] = d0;
			descs[i].colorspace = (unsigned char)(i & 1);
			poff[i] = mis + (uint64_t)i * img_bytes;
			soff[i] = 3 + (uint64_t)i * slot;
		}
		/* pixel buffer 16-byte aligned base + misalignment; guard bytes both sides */
		std::vector<unsigned char> pixstore(img_bytes * n + mis + 64 + 16);
		unsigned char *pixbase = (unsigned char *)(((uintptr_t)pixstore.data() + 15) & ~(uintptr_t)15);
		memcpy(pixbase + mis, file.data(), file.size());

		for (int dec_ch = 3; dec_ch <= 4; dec_ch++) {
			QoiPlanGeom g;
			if (qoi_plan_geometry(g, n, descs.data(), poff.data(), soff.data(), nullptr, dec_ch)) { fprintf(stderr, "geometry rejected\n"); return 2; }
			QoiStateLayout SL = qoi_state_layout(g.n_tiles);
			std::vector<unsigned char> state(SL.total + 256, 0);
			unsigned char *sbase = (unsigned char *)(((uintptr_t)state.data() + 255) & ~(uintptr_t)255);
			std::vector<unsigned char> streams(g.stream_bytes + 64, 0xEE);
			std::vector<int32_t> lens(n, -1);
			QoiEncParams P;
			memset(&P, 0, sizeof P);
			P.pixels = pixbase;
			P.streams = streams.data();
			P.lens = lens.data();
			P.imgs = g.metas.data();
			P.tile_img = g.tile_img.data();
			P.tile_order = g.tile_order.data();
			P.n_images = (uint32_t)n;
			P.n_tiles = g.n_tiles;
			qoi_state_bind(P.st, sbase, SL);
			bool aligned = ch == 4; /* the zero-copy kernel needs every image 16-byte aligned (the C ABI checks the same) */
			for (int i = 0; i < n; i++) aligned = aligned && (((uintptr_t)pixbase + poff[i]) % 16 == 0);
			for (int iter = 0; iter < 2; iter++) {
				P.epoch = 1 + iter;
				std::fill(streams.begin(), streams.end(), (unsigned char)0xEE);
				std::fill(lens.begin(), lens.end(), -1);
				unsigned sd = seed ? seed + 17 * iter : 0;
				if (ch == 4 && aligned && (iter == 0)) run_encode<4, false>(P, grid, sd);
				else if (ch == 4) run_encode<4, true>(P, grid, sd);
				else run_encode<3, true>(P, grid, sd);
				if (P.st.ticket[0] != 0 || P.st.ticket[1] != 0) { printf("FAIL ticket not rearmed\n"); return 1; }
				for (int i = 0; i < n; i++) {
					int olen = 0;
					unsigned char *ref = (unsigned char *)ora_encode(pixbase + poff[i], &descs[i], &olen);
					const unsigned char *got = streams.data() + soff[i];
					if (lens[i] != olen) {
						int fd = first_diff(got, ref, (size_t)(olen < lens[i] || lens[i] < 0 ? olen : lens[i]));
						printf("FAIL encode iter %d image %d: length %d != oracle %d (first differing byte %d)\n", iter, i, lens[i], olen, fd);
						return 1;
					}
					int fd = first_diff(got, ref, (size_t)olen);
					if (fd >= 0) { printf("FAIL encode iter %d image %d: byte %d is %02x, oracle %02x\n", iter, i, fd, got[fd], ref[fd]); return 1; }
					if (got[olen] != 0xEE || (soff[i] && got[-1] != 0xEE)) { printf("FAIL encode image %d: wrote outside its stream\n", i); return 1; }
					free(ref);
				}
			}
			/* decode what we wrote: with the restart index, then as a foreign stream */
			for (int use_index = 1; use_index >= (getenv("QOI_EMU_SKIP_FOREIGN") ? 1 : 0); use_index--) { /* encoder-variant runs skip the any-stream decoder */
				std::vector<unsigned char> out(g.decoded_bytes + 16, 0xAB);
				std::vector<int32_t> status(n, -1);
				QoiDecParams D;
				memset(&D, 0, sizeof D);
				D.streams = streams.data();
				D.lens = lens.data();
				D.out = out.data();
				D.status = status.data();
				D.imgs = g.metas.data();
				D.tile_img = g.tile_img.data();
				D.n_images = (uint32_t)n;
				D.n_jobs = use_index ? g.n_tiles : (uint32_t)n;
				D.out_channels = dec_ch;
				D.use_index = use_index;
				D.st = P.st;
				if (use_index) run_decode(D, seed);
				else run_foreign(D, g, seed, getenv("QOI_EMU_FROUNDS") ? atoi(getenv("QOI_EMU_FROUNDS")) : QOI_FROUNDS);
				for (int i = 0; i < n; i++) {
					QoiDescPOD dd;
					unsigned char *ref = (unsigned char *)ora_decode(streams.data() + soff[i], lens[i], &dd, dec_ch);
					size_t nb = (size_t)w * h * dec_ch;
					int fd = first_diff(out.data() + g.metas[i].dec_off, ref, nb);
					if (fd >= 0 || status[i] != 0) { printf("FAIL decode(index=%d, ch=%d) image %d: byte %d status %d\n", use_index, dec_ch, i, fd, status[i]); return 1; }
					free(ref);
				}
				if (out[g.decoded_bytes] != 0xAB) { printf("FAIL decode wrote past the end\n"); return 1; }
			}
		}
		printf("PASS enc %ux%ux%u n=%d grid=%u seed=%u mis=%u\n", w, h, ch, n, grid, seed, mis);
		return 0;
	}
	if (mode == "dec" && argc == 8) {
		unsigned w = atoi(argv[2]), h = atoi(argv[3]), ch = atoi(argv[4]);
		int n = atoi(argv[5]);
		unsigned seed = atoi(argv[6]);
		std::vector<unsigned char> file = read_file(argv[7]);
		std::vector<QoiDescPOD> descs(n);
		std::vector<uint64_t> poff(n), soff(n);
		std::vector<int32_t> lens(n);
		QoiDescPOD d0 = {w, h, (unsigned char)ch, 0};
		size_t pos = 0;
		for (int i = 0; i < n; i++) {
			memcpy(&lens[i], file.data() + pos, 4);
			soff[i] = pos + 4;
			pos += 4 + (size_t)lens[i];
			descs[i] = d0;
			poff[i] = 0;
		}
		for (int dec_ch = 3; dec_ch <= 4; dec_ch++) {
			QoiPlanGeom g;
			/* stream extents are whatever the file holds, not max size: bypass the extent check by using the file directly */
			if (qoi_plan_geometry(g, n, descs.data(), poff.data(), soff.data(), nullptr, dec_ch)) return 2;
			std::vector<unsigned char> out(g.decoded_bytes + 16, 0xAB);
			std::vector<int32_t> status(n, -1);
			QoiDecParams D;
			memset(&D, 0, sizeof D);
			D.streams = file.data();
			D.lens = lens.data();
			D.out = out.data();
			D.status = status.data();
			D.imgs = g.metas.data();
			D.tile_img = g.tile_img.data();
			D.n_images = (uint32_t)n;
			D.n_jobs = (uint32_t)n;
			D.out_channels = dec_ch;
			D.use_index = 0;
			run_foreign(D, g, seed, getenv("QOI_EMU_FROUNDS") ? atoi(getenv("QOI_EMU_FROUNDS")) : QOI_FROUNDS);
			for (int i = 0; i < n; i++) {
				QoiDescPOD dd;
				unsigned char *ref = (unsigned char *)ora_decode(file.data() + soff[i], lens[i], &dd, dec_ch);
				if (!ref) { printf("FAIL oracle rejected stream %d\n", i); return 1; }
				int fd = first_diff(out.data() + g.metas[i].dec_off, ref, (size_t)w * h * dec_ch);
				if (fd >= 0 || status[i] != 0) { printf("FAIL foreign decode ch=%d stream %d: byte %d status %d\n", dec_ch, i, fd, status[i]); return 1; }
				free(ref);
			}
		}
		printf("PASS dec n=%d\n", n);
		return 0;
	}
	fprintf(stderr, "usage: see the header of emu_main.cpp\n");
	return 2;
}
