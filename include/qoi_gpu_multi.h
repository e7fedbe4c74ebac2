/*
 * qoi_gpu_multi.h -- one process, N devices: the batch API of qoi_gpu.h sharded over the GPUs of a node (libqoi_gpu_nccl.so).
 *
 * Images are independent units (qoi.h:393-400 and 533-537 reset all codec state per call), so a batch shards by image:
 * rank r owns the contiguous block [r * ceil(n / G), ...).  Nothing crosses GPUs inside the codec.  What
 * does cross, over NCCL (NVLink 5 / NVSwitch):
 *   - scatter   root -> ranks: the raw pixels of each rank's block (grouped ncclSend / ncclRecv), when the batch starts on one GPU;
 *   - all-gather of the per-image stream lengths (int32), so that every rank -- and the host -- knows every stream's size;
 *   - gather    ranks -> root: the streams, packed back to back at the exclusive scan of the lengths.
 * The scatter and the gather are optional per call: a batch that is already resident in the shards (loaded or generated in
 * place) needs only the length exchange.  NVLink moves ~0.9 TB/s per direction per GPU against ~6.5 TB/s of HBM, so a
 * root-sourced scatter costs more than the encode itself; qoi_gpu_mencode reports the three phases separately.
 *
 * Additive API, not in the reference (which is single-threaded CPU code: nothing to mirror).  Plain C.
 */
#ifndef QOI_GPU_MULTI_H
#define QOI_GPU_MULTI_H

#include "qoi_gpu.h"

#ifdef __cplusplus
extern "C" {
#endif

typedef struct qoi_gpu_mctx qoi_gpu_mctx;   /* N devices: one qoi_gpu_ctx, one stream and one NCCL communicator each */
typedef struct qoi_gpu_mplan qoi_gpu_mplan; /* a batch sharded over them, with the shard buffers */

/* devices: CUDA ordinals, devices[0] is the root.  ncclCommInitAll over them. */
int qoi_gpu_mctx_create(qoi_gpu_mctx **mctx, const int *devices, int n_devices);
void qoi_gpu_mctx_destroy(qoi_gpu_mctx *mctx);
int qoi_gpu_mctx_size(const qoi_gpu_mctx *mctx);
const char *qoi_gpu_mctx_last_error(const qoi_gpu_mctx *mctx);

/* All images of a plan share one channel count (as in qoi_gpu_plan_create).  Allocates, per rank, device buffers for its
   block: pixels (packed), stream slots, lengths, decoded pixels (decoded_channels bytes per pixel). */
int qoi_gpu_mplan_create(qoi_gpu_mctx *mctx, qoi_gpu_mplan **mplan, int n_images, const qoi_desc *descs, int decoded_channels);
void qoi_gpu_mplan_destroy(qoi_gpu_mplan *mplan);

/* rank r's block and its DEVICE buffers (on devices[r]); pixel_off / stream_off are offsets inside them, one per image of the block */
typedef struct {
	int device, first_image, n_images;
	void *d_pixels, *d_streams, *d_decoded;
	int32_t *d_lens;
	const uint64_t *pixel_off, *stream_off, *decoded_off; /* host arrays, n_images entries */
	uint64_t pixel_bytes, stream_bytes, decoded_bytes;
} qoi_gpu_shard;
int qoi_gpu_mplan_shard(const qoi_gpu_mplan *mplan, int rank, qoi_gpu_shard *out);

/* milliseconds on the device clocks (CUDA events), max over ranks per phase; 0 for a phase that was skipped */
typedef struct {
	float scatter_ms, kernel_ms, lens_ms, gather_ms;
} qoi_gpu_mtimes;

/*
 * Encode the whole batch.
 *   d_root_pixels   on devices[0]: all images packed back to back in batch order (w*h*channels each), or NULL when the
 *                   shards' d_pixels already hold the blocks (no scatter);
 *   d_root_streams  on devices[0]: receives all streams packed back to back in batch order, or NULL (streams stay in the
 *                   shards' slots; no gather);
 *   h_lens          host int32[n_images]: every stream's length (from the all-gather);
 *   h_root_off      host uint64[n_images] or NULL: where each stream starts inside d_root_streams.
 * Synchronous.  Bytes are the same as qoi_gpu_encode of one plan over the whole batch (images are independent).
 */
int qoi_gpu_mencode(qoi_gpu_mplan *mplan, const void *d_root_pixels, void *d_root_streams, int32_t *h_lens, uint64_t *h_root_off,
                    qoi_gpu_mtimes *times);

/*
 * Decode what the last qoi_gpu_mencode wrote (use_index != 0: segment-parallel with each shard's restart index; 0: as foreign
 * streams).  d_root_decoded on devices[0]: all images packed back to back (w*h*decoded_channels each), or NULL (pixels stay
 * in the shards).  Synchronous.
 */
int qoi_gpu_mdecode(qoi_gpu_mplan *mplan, int use_index, void *d_root_decoded, qoi_gpu_mtimes *times);

#ifdef __cplusplus
}
#endif
#endif
