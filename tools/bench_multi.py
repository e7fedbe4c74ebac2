#!/usr/bin/env python
"""bench_multi.py --gpus N [--w 8192 --h 8192 --ch 4 --per-gpu 32 --root-per-gpu 4 --dist 4 --reps 5]

BASELINE.json configs[3] (256 x 8192x8192 RGBA over 8 GPUs = 32 images per GPU) through the multi-device layer
(include/qoi_gpu_multi.h, libqoi_gpu_nccl.so): ONE process, N devices, batch sharded by image blocks.  Two numbers, as
SURVEY.md 8(e) asks:

  device-resident    every shard's pixels already sit on its GPU (loaded / generated in place): encode + all-gather of the
                     stream lengths, indexed decode; no scatter, no gather.  This is how the path is meant to be used.
  incl. scatter/gather   the batch starts and ends on GPU 0: NCCL scatter of the pixel blocks root -> ranks, encode, all-gather
                     of the lengths, gather of the streams to the root; decode, gather of the pixels to the root.  GPU 0 has
                     to hold the WHOLE batch (pixels + worst-case stream slots + decoded pixels), so this runs on
                     --root-per-gpu images per GPU (256 x 8192^2 would need 208 GiB on one 180 GB GPU).

Times are CUDA-event times on the devices, max over ranks per phase (qoi_gpu_mtimes).  Streams of the first and last image of
every shard are compared with the CPU checker (the compiled reference when it travelled with the repository).  MEASUREMENT TOOL."""
import argparse
import ctypes as C
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import qoi_testlib as q  # noqa: E402

import qoi_b200  # noqa: E402
from qoi_b200 import multi as qm  # noqa: E402

rt = C.CDLL("libcudart.so")
rt.cudaMemcpy.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_int]
rt.cudaMalloc.argtypes = [C.POINTER(C.c_void_p), C.c_size_t]
rt.cudaFree.argtypes = [C.c_void_p]
H2D, D2H = 1, 2


def ck(rc, what):
    if rc:
        raise RuntimeError(f"{what}: cuda error {rc}")


def run(n_dev, w, h, ch, per_gpu, dist, reps, with_root, cpu):
    chk = cpu.best()
    n = n_dev * per_gpu
    descs = [qoi_b200.Desc(w, h, ch, 0)] * n
    mctx = qm.MultiContext(list(range(n_dev)))
    plan = qm.MultiPlan(mctx, descs, decoded_channels=4)
    img_bytes = w * h * ch
    px_total = n * w * h
    # distinct images per slot of a shard, the same set on every shard would hide nothing but saves host time: image index = global index
    imgs = {}

    def image(i):
        if i not in imgs:
            imgs[i] = np.ascontiguousarray(cpu.image(w, h, ch, dist, i % 8))  # 8 distinct pictures
        return imgs[i]

    d_root = C.c_void_p()
    d_root_streams = C.c_void_p()
    d_root_dec = C.c_void_p()
    if with_root:
        ck(rt.cudaSetDevice(0), "set 0")
        ck(rt.cudaMalloc(C.byref(d_root), n * img_bytes), "root pixels")
        ck(rt.cudaMalloc(C.byref(d_root_streams), n * (w * h * (ch + 1) + 22)), "root streams")
        ck(rt.cudaMalloc(C.byref(d_root_dec), n * w * h * 4), "root decoded")
        for i in range(n):
            ck(rt.cudaMemcpy(d_root.value + i * img_bytes, image(i).ctypes.data, img_bytes, H2D), "fill root")
    else:
        for r in range(n_dev):
            s = plan.shard(r)
            ck(rt.cudaSetDevice(s.device), "set")
            for k in range(s.n_images):
                ck(rt.cudaMemcpy(s.d_pixels + s.pixel_off[k], image(s.first_image + k).ctypes.data, img_bytes, H2D), "fill shard")
    enc, dec = [], []
    lens = None
    for rep in range(reps + 1):
        lens, offs, te = plan.encode(d_root.value if with_root else None, d_root_streams.value if with_root else None)
        td = plan.decode(True, d_root_dec.value if with_root else None)
        if rep:  # first pass = warm-up
            enc.append((te.scatter_ms, te.kernel_ms, te.lens_ms, te.gather_ms))
            dec.append((td.scatter_ms, td.kernel_ms, td.lens_ms, td.gather_ms))
    # parity: first and last image of every shard against the CPU checker, and the decoded pixels of the same images
    checked = 0
    for r in range(n_dev):
        s = plan.shard(r)
        ck(rt.cudaSetDevice(s.device), "set")
        for k in sorted({0, s.n_images - 1}):
            gi = s.first_image + k
            buf = np.empty(int(lens[gi]), dtype=np.uint8)
            if with_root:
                ck(rt.cudaSetDevice(0), "set 0")
                ck(rt.cudaMemcpy(buf.ctypes.data, d_root_streams.value + int(offs[gi]), buf.size, D2H), "read stream")
            else:
                ck(rt.cudaMemcpy(buf.ctypes.data, s.d_streams + s.stream_off[k], buf.size, D2H), "read stream")
            want = chk.encode(image(gi), w, h, ch)
            assert buf.tobytes() == want, f"stream of image {gi} differs from the CPU checker"
            back = np.empty(w * h * 4, dtype=np.uint8)
            if with_root:
                ck(rt.cudaMemcpy(back.ctypes.data, d_root_dec.value + gi * w * h * 4, back.size, D2H), "read pixels")
            else:
                ck(rt.cudaSetDevice(s.device), "set")
                ck(rt.cudaMemcpy(back.ctypes.data, s.d_decoded + s.decoded_off[k], back.size, D2H), "read pixels")
            src = image(gi)
            ok = np.array_equal(back, src) if ch == 4 else np.array_equal(back.reshape(-1, 4)[:, :3], src.reshape(-1, 3))
            assert ok, f"decoded pixels of image {gi} differ"
            checked += 1
    med = lambda rows, k: float(np.median([r[k] for r in rows]))
    e_sc, e_k, e_l, e_g = (med(enc, k) for k in range(4))
    d_sc, d_k, d_l, d_g = (med(dec, k) for k in range(4))
    enc_bytes = int(lens.astype(np.int64).sum())
    total_ms = e_sc + e_k + e_l + e_g + d_sc + d_k + d_l + d_g
    out = {
        "tool": "bench_multi", "mode": "incl_scatter_gather" if with_root else "device_resident", "n_gpus": n_dev,
        "workload": f"{n} x {w}x{h}x{ch} ({per_gpu} per GPU), dist {dist}, encode + indexed decode to 4 channels",
        "gpx_per_s": px_total / (total_ms * 1e-3) / 1e9, "ms_total": total_ms,
        "encode_ms": {"scatter": e_sc, "kernel": e_k, "lens_allgather": e_l, "gather_streams": e_g},
        "decode_ms": {"kernel": d_k, "gather_pixels": d_g},
        "encoded_bytes": enc_bytes, "bytes_per_px": enc_bytes / px_total,
        "nvlink_gbs": {"scatter": (n - per_gpu) * img_bytes / (e_sc * 1e-3) / 1e9 if e_sc else None,
                       "gather_streams": enc_bytes * (n_dev - 1) / n_dev / (e_g * 1e-3) / 1e9 if e_g else None,
                       "gather_pixels": (n - per_gpu) * w * h * 4 / (d_g * 1e-3) / 1e9 if d_g else None},
        "checked_images": checked, "checker": "compiled reference" if chk is getattr(cpu, "ref", None) else "cpu checker",
    }
    print(json.dumps(out), flush=True)
    for p in (d_root, d_root_streams, d_root_dec):
        if p.value:
            ck(rt.cudaSetDevice(0), "set 0")
            rt.cudaFree(p)
    plan.close()
    mctx.close()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=2)
    ap.add_argument("--w", type=int, default=8192)
    ap.add_argument("--h", type=int, default=8192)
    ap.add_argument("--ch", type=int, default=4)
    ap.add_argument("--per-gpu", type=int, default=32)
    ap.add_argument("--root-per-gpu", type=int, default=4)
    ap.add_argument("--dist", type=int, default=4)
    ap.add_argument("--reps", type=int, default=5)
    a = ap.parse_args()
    cpu = q.cpu_libs()
    run(a.gpus, a.w, a.h, a.ch, a.per_gpu, a.dist, a.reps, False, cpu)
    if a.root_per_gpu > 0:
        run(a.gpus, a.w, a.h, a.ch, a.root_per_gpu, a.dist, a.reps, True, cpu)


if __name__ == "__main__":
    main()
