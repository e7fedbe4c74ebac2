/*
 * qoi_gpu.h -- batch / device-resident C ABI of the B200-native QOI codec.
 *
 * Additive to the reference API (which has no batch or device interface): the
 * four reference entry points are declared in include/qoi.h and exported by the
 * same shared object.  Everything here is plain C: pointers, sizes, an opaque
 * context and an opaque plan.  `cuda_stream` arguments are a cudaStream_t
 * passed as void* (NULL = the legacy default stream) so that a caller holding
 * device memory from any runtime (PyTorch, CuPy, raw CUDA) can order the work
 * on its own stream.
 *
 * Units.  A *plan* fixes the geometry of a batch of images (same channel count
 * for all images of a plan) and owns the scratch the kernels need:
 *   - encode replaces the reference loop qoi.h:406-478 for every image of the
 *     batch in ONE launch of a single-pass chained-scan kernel;
 *   - decode replaces qoi.h:540-587; it runs one lane per *segment*.  A stream
 *     that was written by qoi_gpu_encode of the same plan is cut into segments
 *     at the encoder's tile boundaries using the restart index the encoder
 *     leaves in the plan (px, previous pixel, byte offset, 64-entry table per
 *     tile) -- the index lives OUTSIDE the .qoi stream, which stays bit exact.
 *     Any other stream ("foreign") goes through a parse scan + speculative segment decode +
 *     verification + exact fix-up (cuda/qoi_fdec.cuh): exact for every input, parallel inside
 *     one stream as far as its content allows.
 *
 * All functions return QOI_GPU_OK (0) or a QOI_GPU_E_* code; none of them
 * falls back to the CPU.
 */
#ifndef QOI_GPU_H
#define QOI_GPU_H

#include <stddef.h>
#include <stdint.h>
#include "qoi.h"

#ifdef __cplusplus
extern "C" {
#endif

#define QOI_GPU_OK        0
#define QOI_GPU_E_ARG     1 /* NULL / malformed argument, geometry rejected like qoi.h:364-372 */
#define QOI_GPU_E_CUDA    2 /* a CUDA call failed (no device, launch error, ...) */
#define QOI_GPU_E_NOMEM   3
#define QOI_GPU_E_STREAM  4 /* a stream header does not match the plan (device-side check) */

#define QOI_GPU_TILE_PX   2048 /* pixels per encoder tile == pixels per indexed decode segment */

typedef struct qoi_gpu_ctx qoi_gpu_ctx;
typedef struct qoi_gpu_plan qoi_gpu_plan;

/* number of CUDA devices visible, 0 if none / no driver */
int qoi_gpu_device_count(void);

/* one context per device; owns a stream for the host-buffer entry points */
int qoi_gpu_ctx_create(qoi_gpu_ctx **ctx, int device);
void qoi_gpu_ctx_destroy(qoi_gpu_ctx *ctx);
/* text of the last CUDA error seen by this context ("" if none) */
const char *qoi_gpu_last_error(const qoi_gpu_ctx *ctx);

/* worst-case stream size of one image: w*h*(channels+1) + 14 + 8  (qoi.h:374-376) */
size_t qoi_gpu_max_stream_size(const qoi_desc *desc);

/*
 * Geometry of a batch.  For image i:
 *   pixel_off[i]   byte offset of its raw pixels inside the pixel buffer (encode input), stride = channels;
 *   stream_off[i]  byte offset of its stream slot inside the stream buffer; the slot must hold
 *                  qoi_gpu_max_stream_size(&descs[i]) bytes;
 *   decoded_off[i] byte offset of its decoded pixels inside the decode output buffer (may be NULL:
 *                  images are then packed back to back with the out_channels stride given here).
 * Any alignment is accepted; 16-byte aligned pixel offsets take the zero-copy path for RGBA.
 * All descs must pass the reference's checks (qoi.h:364-372) and share one channel count.
 */
int qoi_gpu_plan_create(qoi_gpu_ctx *ctx, qoi_gpu_plan **plan, int n_images, const qoi_desc *descs,
                        const uint64_t *pixel_off, const uint64_t *stream_off, const uint64_t *decoded_off,
                        int decoded_channels);
void qoi_gpu_plan_destroy(qoi_gpu_plan *plan);

/* totals a caller needs to size its buffers */
uint64_t qoi_gpu_plan_pixels(const qoi_gpu_plan *plan);       /* sum of w*h */
uint32_t qoi_gpu_plan_tiles(const qoi_gpu_plan *plan);        /* encoder tiles == indexed decode segments */

/*
 * Encode the whole batch.  d_pixels, d_streams, d_lens are DEVICE pointers;
 * d_lens receives n_images int32 stream lengths (header and end marker included).
 * Asynchronous on cuda_stream.  Also (re)writes the plan's restart index.
 */
int qoi_gpu_encode(qoi_gpu_plan *plan, const void *d_pixels, void *d_streams, int32_t *d_lens, void *cuda_stream);

/*
 * Decode the whole batch to `decoded_channels` (3 or 4) bytes per pixel.
 * d_lens: device int32[n_images] stream sizes (as qoi_decode's `size`).
 * use_index != 0: the streams are the ones the last qoi_gpu_encode on this plan wrote; decode them
 *                 segment-parallel with the plan's restart index.
 * use_index == 0: treat every stream as foreign -- no index needed (what qoi_decode of the C API uses).
 * d_status (may be NULL): device int32[n_images], set to 0 or QOI_GPU_E_STREAM per image when the
 *                 header in the stream disagrees with the plan (magic, size, channels, colorspace).
 */
int qoi_gpu_decode(qoi_gpu_plan *plan, const void *d_streams, const int32_t *d_lens, void *d_decoded,
                   int use_index, int32_t *d_status, void *cuda_stream);

/*
 * Host-buffer forms of the two calls above (copies inside): h_* are HOST pointers, pinned
 * (qoi_gpu_host_alloc) or pageable.  Synchronous.  h_streams uses the plan's stream_off slots.
 */
int qoi_gpu_encode_host(qoi_gpu_plan *plan, const void *h_pixels, void *h_streams, int32_t *h_lens);
int qoi_gpu_decode_host(qoi_gpu_plan *plan, const void *h_streams, const int32_t *h_lens, void *h_decoded,
                        int use_index);

/*
 * Restart index as a side-car.  The index the last qoi_gpu_encode left in the plan -- per 2048-pixel tile:
 * where its first op starts (pixel, stream offset, colour before it) and the reference decoder's index[64] at that point --
 * can be copied out and stored next to the .qoi files (0.13 bytes per pixel; the streams themselves stay bit exact and
 * decode anywhere without it).  A plan of the same geometry that imports it decodes those streams with use_index = 1,
 * i.e. at the segment-parallel speed, after they have been read back from disk.
 *   qoi_gpu_index_bytes   size of the blob for this plan's geometry
 *   qoi_gpu_index_export  device -> host blob (HOST pointer, capacity >= qoi_gpu_index_bytes); needs an index in the plan
 *   qoi_gpu_index_import  host blob -> plan; QOI_GPU_E_ARG if the blob does not describe this plan's images
 * Both are synchronous on cuda_stream.
 */
size_t qoi_gpu_index_bytes(const qoi_gpu_plan *plan);
int qoi_gpu_index_export(qoi_gpu_plan *plan, void *h_index, size_t capacity, void *cuda_stream);
int qoi_gpu_index_import(qoi_gpu_plan *plan, const void *h_index, size_t bytes, void *cuda_stream);

/* pinned host memory for the *_host calls */
void *qoi_gpu_host_alloc(size_t bytes);
void qoi_gpu_host_free(void *p);

/* Diagnostics of the last qoi_gpu_decode(use_index = 0) on the plan: out[r] = segments state round r walked (r < 60;
   out[0] = 0 stands for "all"), out[64 + r] = streams round r's scan still had to look at (0: settled, later launches return
   at once); out[120..127] = 1 + index of the first images whose streams were finished by the sequential fall-back (0: none).
   Synchronises the device.  QOI_GPU_E_ARG if the plan has never decoded a foreign stream. */
int qoi_gpu_plan_fstats(qoi_gpu_plan *plan, uint32_t out[128]);

/* number of kernel launches issued by this library since load (bench.py's gpu_launches) */
uint64_t qoi_gpu_launch_count(void);

#ifdef __cplusplus
}
#endif
#endif /* QOI_GPU_H */
