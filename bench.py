#!/usr/bin/env python
"""bench.py -- the headline benchmark of the B200-native QOI codec.

Metric (BASELINE.json): Gpixels/s of an encode + decode round trip over a batch of synthetic
4096x4096 RGBA images, per-GPU batch fixed (weak scaling), 1/2/4/8 GPUs, next to the fraction of the
HBM roofline of the dominant kernel and the reference qoi.h timed on the host cores.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--config headline|1080p_rgb_1024|8192_rgba_32|sweep] [--batch B] [--dist D] [--impl reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

One JSON line on stdout (rank 0).  A "step" = qoi_gpu_encode of the whole batch followed by
qoi_gpu_decode of the streams it wrote (segment-parallel, using the restart index the encoder left
in the plan).  `value` is device resident; `e2e` runs the same step through the host-buffer C ABI
(qoi_gpu_encode_host / qoi_gpu_decode_host) from pinned host memory, copies inside the timed region.
The oracle / compiled reference is executed only in the cpu_baseline leg and under --impl reference.
"""
import argparse
import ctypes as C
import json
import os
import statistics
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "Gpixels/s encode+decode round trip, 4096x4096 RGBA batch"
# BASELINE.json configs that are bench lines here (the others are parity-test cases): name -> (w, h, channels, images per GPU)
CONFIGS = {
    "headline": (4096, 4096, 4, 32),          # the configuration the metric is quoted on
    "1080p_rgb_1024": (1920, 1080, 3, 1024),  # configs[2]: batch of 1024 synthetic 1920x1080 RGB images, decode to 4 channels
    "8192_rgba_32": (8192, 8192, 4, 32),      # configs[3] per-GPU share: 256 x 8192x8192 RGBA over 8 GPUs
}
SWEEP_SIZES = [64, 128, 256, 512, 1024, 2048, 4096, 8192, 16384]  # configs[4]: size sweep x channels 3/4
W = H = 4096
CH = 4
DIST_NAMES = ["noise_rgba", "noise_opaque", "gradient", "grad2", "photo", "tiles32", "palette16", "const0"]


def shard_images(global_batch, world, rank):
    """Contiguous block partition of image indices: rank r owns [lo, hi)."""
    per = (global_batch + world - 1) // world
    lo = min(global_batch, rank * per)
    hi = min(global_batch, lo + per)
    return lo, hi


def load_synth():
    so = os.path.join(ROOT, "tools", "libqoi_synth.so")
    if not os.path.exists(so):
        subprocess.check_call(["make", "-s", "-C", os.path.join(ROOT, "oracle")], stdout=subprocess.DEVNULL)
    lib = C.CDLL(so)
    lib.qoi_synth_batch.argtypes = [C.c_void_p, C.c_size_t, C.c_int, C.c_uint, C.c_uint, C.c_int, C.c_int, C.c_uint]
    return lib


def load_cpu_codec():
    """(kind, lib, prefix): the compiled reference if it travelled with the repo, else the oracle port."""
    ref = os.path.join(ROOT, "oracle", "_ref", "libqoi_ref.so")
    ora = os.path.join(ROOT, "oracle", "liboracle.so")
    if os.path.exists(ref):
        kind, path, prefix = "reference", ref, "ref_"
    else:
        if not os.path.exists(ora):
            subprocess.check_call(["make", "-s", "-C", os.path.join(ROOT, "oracle")], stdout=subprocess.DEVNULL)
        kind, path, prefix = "port", ora, "ora_"
    lib = C.CDLL(path)
    eb = getattr(lib, prefix + "encode_batch")
    eb.restype = C.c_double
    eb.argtypes = [C.c_void_p, C.c_size_t, C.c_int, C.c_void_p, C.c_int, C.c_void_p, C.c_size_t, C.c_void_p]
    db = getattr(lib, prefix + "decode_batch")
    db.restype = C.c_double
    db.argtypes = [C.c_void_p, C.c_size_t, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_size_t]
    # torchrun exports OMP_NUM_THREADS=1; the CPU legs pass an explicit thread count (num_threads clause) instead
    try:
        threads = len(os.sched_getaffinity(0))
    except AttributeError:
        threads = os.cpu_count() or 1
    return kind, eb, db, max(1, threads)


class DescC(C.Structure):
    _fields_ = [("width", C.c_uint), ("height", C.c_uint), ("channels", C.c_ubyte), ("colorspace", C.c_ubyte)]


def workload_name(n):
    return f"{n} x {W}x{H} {'RGBA' if CH == 4 else 'RGB'} per GPU, encode+decode round trip (decode to 4 channels)"


def cpu_round_trip(eb, db, threads, pixels, n, keep_streams, keep_lens, slot):
    """one encode + decode pass of the CPU codec over n images; returns seconds (encode, decode)."""
    d = DescC(W, H, CH, 0)
    te = eb(pixels.ctypes.data, W * H * CH, n, C.byref(d), threads, keep_streams.ctypes.data, slot, keep_lens.ctypes.data)
    td = db(keep_streams.ctypes.data, slot, keep_lens.ctypes.data, n, 4, threads, None, 0)
    if te < 0 or td < 0:
        raise RuntimeError("CPU codec returned NULL")
    return te, td


class ClockSampler:
    """nvidia-smi polled every 20 ms in the background.  It is started before the warm-up (the tool needs a few hundred
    ms to come up) and the samples are cut to the wall-clock window of the timed region afterwards."""
    QUERY = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap,timestamp")

    def __init__(self, device_index):
        self.file = tempfile.NamedTemporaryFile(prefix="qoi_clocks_", suffix=".csv", delete=False)
        self.proc = None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(device_index), "--query-gpu=" + self.QUERY,
                                          "--format=csv,noheader,nounits", "-lms", "20"],
                                         stdout=self.file, stderr=subprocess.DEVNULL)
        except OSError:
            self.proc = None

    def stop(self, t_begin=None, t_end=None):
        """t_begin / t_end: time.time() around the timed region; samples outside are dropped (all kept if none falls inside)."""
        import datetime
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        if self.proc is None:
            return out
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        self.file.close()
        rows = []
        sm, mx, reasons = [], [], set()
        try:
            for line in open(self.file.name):
                f = [x.strip() for x in line.split(",")]
                if len(f) < 9:
                    continue
                try:
                    rec = (float(f[1]), float(f[2]), [v.lower().startswith("active") for v in f[5:9]])
                except ValueError:
                    continue
                ts = None
                if len(f) >= 10:
                    try:
                        ts = datetime.datetime.strptime(f[9], "%Y/%m/%d %H:%M:%S.%f").timestamp()
                    except ValueError:
                        ts = None
                rows.append((ts, rec))
            inside = [r for ts, r in rows if ts is not None and t_begin is not None and t_begin - 0.02 <= ts <= t_end + 0.02]
            use = inside if inside else [r for _, r in rows]
            out["window"] = "timed region" if inside else "whole run (no sample fell inside the timed region)"
            sm, mx, reasons = [], [], set()
            for a, b, flags in use:
                sm.append(a); mx.append(b)
                for name, on in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), flags):
                    if on:
                        reasons.add(name)
        finally:
            os.unlink(self.file.name)
        if sm:
            out.update(sm_mhz=statistics.median(sm), sm_max_mhz=max(mx), reasons=sorted(reasons), samples=len(sm))
        return out


def measured_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


def recorded_traffic():
    """dram bytes per launch of the two kernels from the committed ncu capture (a RECORDED value of the headline workload,
    not a measurement of this run: bench.py never runs under a profiler)."""
    p = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(p):
        try:
            return json.load(open(p))
        except Exception:
            return None
    return None


def pcie_ceiling(torch, nbytes=1 << 30):
    """bare pinned cudaMemcpyAsync of nbytes: host->device alone, device->host alone, both at once (GB/s per direction).
    The e2e leg moves the same bytes both ways at the same time; this is its ceiling on this box."""
    try:
        h_a = torch.empty(nbytes, dtype=torch.uint8).pin_memory()
        h_b = torch.empty(nbytes, dtype=torch.uint8).pin_memory()
        d_a = torch.empty(nbytes, dtype=torch.uint8, device="cuda")
        d_b = torch.empty(nbytes, dtype=torch.uint8, device="cuda")
        s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()

        def timed(fn, reps=3):
            fn(); torch.cuda.synchronize()
            t0 = time.perf_counter()
            for _ in range(reps):
                fn()
            torch.cuda.synchronize()
            return (time.perf_counter() - t0) / reps

        def h2d():
            with torch.cuda.stream(s1):
                d_a.copy_(h_a, non_blocking=True)

        def d2h():
            with torch.cuda.stream(s2):
                h_b.copy_(d_b, non_blocking=True)

        def both():
            h2d(); d2h()

        t1, t2, t3 = timed(h2d), timed(d2h), timed(both)
        return {"h2d_gbs": nbytes / t1 / 1e9, "d2h_gbs": nbytes / t2 / 1e9, "both_gbs_per_direction": nbytes / t3 / 1e9, "bytes": nbytes}
    except Exception as e:  # an extra; never let it stop the line
        return {"error": repr(e)}


def config_dict(args, world):
    """identical in both arms (the driver compares them): only what follows from the command line"""
    n = args.batch
    pixel_bytes = n * W * H * CH
    return {"workload": workload_name(n), "name": args.config, "width": W, "height": H, "channels": CH, "batch_per_gpu": n,
            "global_batch": n * world, "distribution": DIST_NAMES[args.dist], "decode_channels": 4,
            "parallelism": f"batch sharded over {world} GPU(s), no data-path collective",
            "l2": (f"inputs {pixel_bytes >> 20} MiB per GPU per step, larger than the 126 MB L2 (no flush needed)" if pixel_bytes > (126 << 20)
                   else "inputs smaller than L2: a 256 MiB buffer is written between timed steps (outside the step's events)"),
            "decode_path": "segment-parallel with the encoder's restart index (side table, 0.13 B/px, outside the bit-exact stream); "
                           "roofline.decode_foreign is the any-stream path"}


def run_reference_arm(args, rank, world):
    """--impl reference: the reference's CPU implementation on the host cores, all threads, same config."""
    if rank != 0:
        return
    kind, eb, db, threads = load_cpu_codec()
    synth = load_synth()
    # images are the only parallel axis of qoi.h: give every host thread one; bounded so that a step stays within seconds
    n = max(min(args.batch, 2 * threads), threads)
    while n > threads and n * W * H > (1 << 32):
        n -= 1
    pixels = np.empty(n * W * H * CH, dtype=np.uint8)
    synth.qoi_synth_batch(pixels.ctypes.data, W * H * CH, n, W, H, CH, args.dist, 0)
    slot = W * H * (CH + 1) + 22
    streams = np.empty(n * slot, dtype=np.uint8)
    lens = np.zeros(n, dtype=np.int32)
    for _ in range(args.warmup):
        cpu_round_trip(eb, db, threads, pixels, n, streams, lens, slot)
    t_enc = t_dec = 0.0
    for _ in range(args.steps):
        te, td = cpu_round_trip(eb, db, threads, pixels, n, streams, lens, slot)
        t_enc += te
        t_dec += td
    total = t_enc + t_dec
    px = n * W * H * args.steps
    value = px / total / 1e9
    sample = (f"{n} x {W}x{H}x{CH} images per step (a bounded sample of the workload: host threads, at most the per-GPU batch "
              f"if that is larger), OpenMP over images, "
              "malloc+free inside the call as qoibench.c:446-450,471-481")
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": "Gpx/s", "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * total / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "u8", "data": "synthetic",
        "config": config_dict(args, args.gpus),
        "cpu_baseline": {"value": value, "unit": "Gpx/s", "cores": threads, "kind": kind, "sample": sample,
                         "encode_gpx": px / t_enc / 1e9, "decode_gpx": px / t_dec / 1e9},
        "e2e": {"value": value, "unit": "Gpx/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def all_threads_sample(cpu_baseline, eb, db, threads, h_pix, n, m, img_bytes, slot):
    """cpu_baseline with every host thread busy, as in the --impl reference arm: one image per thread, the batch's images
    repeated (qoi.h codes one image per call: images are its only parallel axis).  The batch-sized pass stays in
    cpu_baseline.batch_sample; skipped when the buffers would not fit, and never allowed to cost the line."""
    try:
        if threads <= m or threads * (img_bytes + slot) > (24 << 30):
            return
        reps_px = np.concatenate([h_pix[:n * img_bytes]] * ((threads + n - 1) // n))[:threads * img_bytes]
        kk = np.empty(threads * slot, dtype=np.uint8)
        kl = np.zeros(threads, dtype=np.int32)
        cpu_round_trip(eb, db, threads, reps_px, threads, kk, kl, slot)
        tea, tda = cpu_round_trip(eb, db, threads, reps_px, threads, kk, kl, slot)
        cpu_baseline["batch_sample"] = {k: cpu_baseline[k] for k in ("value", "cores", "sample", "encode_gpx", "decode_gpx")}
        cpu_baseline.update({"value": threads * W * H / (tea + tda) / 1e9, "cores": threads,
                             "sample": f"{threads} x {W}x{H}x{CH} images (one per host thread: the batch's images repeated), one encode+decode pass "
                                       "after a warm-up pass, OpenMP over images, malloc+free inside the call (qoibench.c:446-450,471-481); the "
                                       "GPU streams of the batch were verified identical on the batch_sample pass",
                             "encode_gpx": threads * W * H / tea / 1e9, "decode_gpx": threads * W * H / tda / 1e9})
    except Exception as e:
        cpu_baseline["all_threads_error"] = repr(e)


def c_api_single_image(h_pix, args):
    """extra.c_api_single_image: the four-function drop-in itself, one image per call (see the note in the result)."""
    # the four-function drop-in itself, called the way qoibench.c calls qoi.h: one image per call, pageable
    # host memory, malloc'ed results (qoi_decode decodes a stream as ANY stream: no restart index)
    import qoi_b200
    from qoi_b200 import api as _api
    one = np.ascontiguousarray(h_pix[:W * H * CH])
    d1 = qoi_b200.Desc(W, H, CH, 0)
    L = _api.lib()
    libc = C.CDLL(None)
    libc.free.argtypes = [C.c_void_p]
    nbytes, d2 = C.c_int(0), qoi_b200.Desc(0, 0, 0, 0)

    def c_encode():
        p_ = L.qoi_encode(one.ctypes.data, C.byref(d1), C.byref(nbytes))
        if not p_:
            raise RuntimeError("qoi_encode returned NULL")
        return p_

    def c_decode(stream_ptr):
        q_ = L.qoi_decode(stream_ptr, nbytes.value, C.byref(d2), 4)
        if not q_:
            raise RuntimeError("qoi_decode returned NULL")
        return q_

    sp = c_encode()                          # first calls build the cached plan and its scratch (untimed)
    libc.free(c_decode(sp))
    reps = 5
    t_enc = t_dec = t_free = 0.0
    q = None
    for _ in range(reps):               # the calls and the caller's free() of their results are timed apart
        t0 = time.perf_counter()
        libc.free(sp)
        if q:
            libc.free(q)
        t1 = time.perf_counter()
        sp = c_encode()
        t2 = time.perf_counter()
        q = c_decode(sp)
        t3 = time.perf_counter()
        t_free += t1 - t0
        t_enc += t2 - t1
        t_dec += t3 - t2
    back = np.frombuffer(C.string_at(q, W * H * 4), dtype=np.uint8)
    libc.free(q)
    libc.free(sp)
    same = np.array_equal(back, one) if CH == 4 else np.array_equal(back.reshape(-1, 4)[:, :3], one.reshape(-1, 3))
    if not same:
        raise RuntimeError("C API round trip mismatch")
    kind1, eb1, db1, _t = load_cpu_codec()
    k1 = np.empty(W * H * (CH + 1) + 22, dtype=np.uint8)
    l1 = np.zeros(1, dtype=np.int32)
    cpu_round_trip(eb1, db1, 1, one, 1, k1, l1, k1.size)
    te1, td1 = cpu_round_trip(eb1, db1, 1, one, 1, k1, l1, k1.size)
    c_api = {"image": f"{W}x{H}x{CH} {DIST_NAMES[args.dist]}", "qoi_encode_ms": 1e3 * t_enc / reps,
             "qoi_decode_ms": 1e3 * t_dec / reps, "free_of_both_results_ms": 1e3 * t_free / reps,
             "cpu_kind": kind1, "cpu_qoi_encode_ms": 1e3 * te1, "cpu_qoi_decode_ms": 1e3 * td1,
             "note": "single calls of the reference API, host pointers in and out, copies and the any-stream decoder inside; the caller's "
                     "free() of the two malloc'ed results (munmap of 64 + 80 MiB whose pages the library's pool threads touched) is timed apart; "
                     "the CPU columns include malloc + free inside the call as qoibench.c:446-450,471-481 times them"}
    return c_api



def bind_to_gpu_numa_node(local_rank):
    """Run this rank on the CPUs of the NUMA node its GPU hangs off, so that the pinned host buffers it allocates from
    here on (and the DMA traffic of the e2e leg) stay on the GPU's side of the socket interconnect.  Returns a small
    dict for the JSON line, or None where the platform does not say (VMs often report numa_node = -1).
    QOI_BENCH_NUMA=0 switches it off.  The caller restores the original affinity before any CPU-codec leg."""
    if os.environ.get("QOI_BENCH_NUMA", "1") == "0":
        return None
    try:
        import torch

        p = torch.cuda.get_device_properties(local_rank)
        bdf = "%04x:%02x:%02x.0" % (p.pci_domain_id, p.pci_bus_id, p.pci_device_id)
        node = int(open(f"/sys/bus/pci/devices/{bdf}/numa_node").read())
        if node < 0:
            return None
        cpus = set()
        for part in open(f"/sys/devices/system/node/node{node}/cpulist").read().strip().split(","):
            a, _, b = part.partition("-")
            cpus.update(range(int(a), int(b or a) + 1))
        allowed = os.sched_getaffinity(0)
        use = cpus & allowed
        if not use or use == allowed:
            return {"node": node, "cpus": len(use), "bound": False}
        os.sched_setaffinity(0, use)
        return {"node": node, "cpus": len(use), "bound": True}
    except Exception:  # never let a missing sysfs file or attribute stop the benchmark
        return None


def set_workload(w, h, ch):
    global W, H, CH
    W, H, CH = w, h, ch


def run_b200(args, rank, world, local_rank, torch, dist, emit=True):
    """one workload on this rank's GPU; rank 0 returns (and prints) the JSON line"""
    import qoi_b200
    from qoi_b200.api import pinned_array

    dist_on = world > 1
    full_affinity = os.sched_getaffinity(0)
    numa = bind_to_gpu_numa_node(local_rank)  # before the context and the pinned buffers exist
    n = args.batch
    img_bytes = W * H * CH
    synth = load_synth()
    ctx = qoi_b200.Context(local_rank)
    descs = [qoi_b200.Desc(W, H, CH, 0)] * n
    plan = qoi_b200.Plan(ctx, descs, decoded_channels=4)

    # ---- inputs: this rank's shard of the global batch, generated on the host (untimed), pinned
    lo, hi = shard_images(n * world, world, rank)
    assert hi - lo == n
    h_pix = pinned_array(plan.pixel_bytes)
    synth.qoi_synth_batch(h_pix.ctypes.data, img_bytes, n, W, H, CH, args.dist, lo)
    d_pix = torch.empty(plan.pixel_bytes, dtype=torch.uint8, device="cuda")
    d_pix.copy_(torch.from_numpy(h_pix))
    d_streams = torch.empty(plan.stream_bytes, dtype=torch.uint8, device="cuda")
    d_lens = torch.zeros(n, dtype=torch.int32, device="cuda")
    d_dec = torch.empty(plan.decoded_bytes, dtype=torch.uint8, device="cuda")
    stream = torch.cuda.current_stream().cuda_stream
    small = plan.pixel_bytes <= (126 << 20)  # inputs fit into L2: flush between timed steps
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda") if small else None

    def step(events=None):
        if flush is not None:
            flush.fill_(1)
        if events is not None:
            events[0].record()
        plan.encode(d_pix.data_ptr(), d_streams.data_ptr(), d_lens.data_ptr(), stream)
        if events is not None:
            events[1].record()
        plan.decode(d_streams.data_ptr(), d_lens.data_ptr(), d_dec.data_ptr(), True, None, stream)
        if events is not None:
            events[2].record()

    def barrier():
        if dist_on:
            dist.barrier()
        torch.cuda.synchronize()

    def decoded_equals_input():
        if CH == 4:
            return torch.equal(d_dec, d_pix)
        a = d_dec.view(-1, 4)
        return bool((a[:, :3] == d_pix.view(-1, 3)).all()) and bool((a[:, 3] == 255).all())

    sampler = ClockSampler(local_rank) if rank == 0 else None  # started early, cut to the timed window later
    # ---- warm-up + correctness of the round trip (untimed)
    for _ in range(args.warmup):
        step()
    torch.cuda.synchronize()
    if not decoded_equals_input():
        raise SystemExit("round trip mismatch: decode(encode(x)) != x")
    lens = d_lens.cpu().numpy().astype(np.int64)
    enc_bytes = int(lens.sum())
    px_rank = n * W * H
    e_bpp = enc_bytes / px_rank

    # ---- timed region: exactly K steps, events on the launching stream, clocks sampled meanwhile
    evs = [[torch.cuda.Event(enable_timing=True) for _ in range(3)] for _ in range(args.steps)]
    launches0 = qoi_b200.launch_count()
    barrier()
    t_start = torch.cuda.Event(enable_timing=True)
    t_stop = torch.cuda.Event(enable_timing=True)
    wall_begin = time.time()
    t_start.record()
    for k in range(args.steps):
        step(evs[k])
    t_stop.record()
    barrier()
    launches = qoi_b200.launch_count() - launches0
    wall_end = time.time()
    clocks = sampler.stop(wall_begin, wall_end) if sampler else None
    enc_ms = sum(e[0].elapsed_time(e[1]) for e in evs) / args.steps
    dec_ms = sum(e[1].elapsed_time(e[2]) for e in evs) / args.steps
    total_ms = t_start.elapsed_time(t_stop) if flush is None else (enc_ms + dec_ms) * args.steps  # the L2 flush is not part of a step

    # ---- e2e: same step through the host-buffer C ABI, pinned host memory, copies inside
    e2e_s, h2d, d2h = None, 0, 0
    h_streams = h_lens = None
    if args.e2e_steps > 0:
        h_streams = pinned_array(plan.stream_bytes)
        h_lens = np.zeros(n, dtype=np.int32)
        h_dec = pinned_array(plan.decoded_bytes)

        def e2e_step():
            plan.encode_host(h_pix, h_streams, h_lens)
            plan.decode_host(h_streams, h_lens, h_dec, use_index=True)

        e2e_step()
        ok = np.array_equal(h_dec, h_pix) if CH == 4 else (np.array_equal(h_dec.reshape(-1, 4)[:, :3], h_pix.reshape(-1, 3)))
        if not ok:
            raise SystemExit("e2e round trip mismatch")
        barrier()
        t0 = time.perf_counter()
        for _ in range(args.e2e_steps):
            e2e_step()
        barrier()
        e2e_s = (time.perf_counter() - t0) / args.e2e_steps
        h2d = plan.pixel_bytes + enc_bytes + 4 * n
        d2h = enc_bytes + plan.decoded_bytes + 4 * n
        del h_dec

    # ---- foreign-stream decode (no restart index, the path qoi_decode of the C API takes): reported beside the headline, not part of it
    fe = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
    plan.decode(d_streams.data_ptr(), d_lens.data_ptr(), d_dec.data_ptr(), False, None, stream)
    fe[0].record()
    plan.decode(d_streams.data_ptr(), d_lens.data_ptr(), d_dec.data_ptr(), False, None, stream)
    fe[1].record()
    torch.cuda.synchronize()
    foreign_ms = fe[0].elapsed_time(fe[1])
    if not decoded_equals_input():
        raise SystemExit("any-stream decode mismatch")

    # ---- N > 1: every rank checks one of ITS streams against the CPU codec (sharding must not change bytes)
    shard_ok = 1.0
    if dist_on:
        kind_, eb_, db_, _t = load_cpu_codec()
        k1 = np.empty(W * H * (CH + 1) + 22, dtype=np.uint8)
        l1 = np.zeros(1, dtype=np.int32)
        pick = (rank * 7) % n
        one = np.ascontiguousarray(h_pix[pick * img_bytes:(pick + 1) * img_bytes])
        cpu_round_trip(eb_, db_, 1, one, 1, k1, l1, k1.size)
        o = int(plan.stream_off[pick])
        got = d_streams[o:o + int(l1[0])].cpu().numpy()
        shard_ok = 1.0 if (int(lens[pick]) == int(l1[0]) and np.array_equal(got, k1[:int(l1[0])])) else 0.0

    # ---- max over ranks
    vals = torch.tensor([total_ms, enc_ms, dec_ms, (e2e_s or 0.0) * 1e3, foreign_ms, 1.0 - shard_ok], dtype=torch.float64, device="cuda")
    sums = torch.tensor([float(enc_bytes), float(launches)], dtype=torch.float64, device="cuda")
    if dist_on:
        dist.all_reduce(vals, op=dist.ReduceOp.MAX)
        dist.all_reduce(sums, op=dist.ReduceOp.SUM)  # sizes exchange (the only collective; outside the timed region)
    total_ms, enc_ms, dec_ms, e2e_ms, foreign_ms, shard_bad = [float(x) for x in vals.tolist()]
    enc_bytes_all, launches_all = [float(x) for x in sums.tolist()]
    if shard_bad:
        raise SystemExit("a rank's stream differs from the CPU codec's (sharding changed bytes)")

    line = None
    if rank == 0:
        px_all = px_rank * world
        ms_per_step = total_ms / args.steps
        value = px_all / (ms_per_step * 1e-3) / 1e9
        peak, peak_src = measured_peak()
        enc_alg = px_rank * CH + enc_bytes          # bytes one encode launch must move (read ch B/px + write e B/px)
        dec_alg = enc_bytes + px_rank * 4           # bytes one decode launch must move (read e B/px + write 4 B/px)
        enc_gbs = enc_alg / (enc_ms * 1e-3) / 1e9
        dec_gbs = dec_alg / (dec_ms * 1e-3) / 1e9
        rt_gbs = (enc_alg + dec_alg) / ((enc_ms + dec_ms) * 1e-3) / 1e9
        for_gbs = dec_alg / (foreign_ms * 1e-3) / 1e9
        dominant = "decode" if dec_ms >= enc_ms else "encode"
        traffic = recorded_traffic() if args.config == "headline" else None
        tsrc = "recorded profiles/traffic.json (ncu --set full of the headline workload; not measured in this run)" if traffic else None
        roofline = {
            "bound": "hbm", "kernel": "qoi_decode_kernel<4,128>" if dominant == "decode" else f"qoi_encode_kernel<{CH},{'false' if CH == 4 else 'true'}>",
            "achieved": dec_gbs if dominant == "decode" else enc_gbs, "peak": peak, "unit": "GB/s",
            "frac": (dec_gbs if dominant == "decode" else enc_gbs) / peak, "peak_source": peak_src,
            "traffic": (traffic or {}).get(dominant), "traffic_source": tsrc,
            "encode": {"ms": enc_ms, "achieved": enc_gbs, "frac": enc_gbs / peak, "algorithmic_bytes": enc_alg,
                       "gpx_per_s": px_rank / (enc_ms * 1e-3) / 1e9, "traffic": (traffic or {}).get("encode")},
            "decode": {"ms": dec_ms, "achieved": dec_gbs, "frac": dec_gbs / peak, "algorithmic_bytes": dec_alg,
                       "gpx_per_s": px_rank / (dec_ms * 1e-3) / 1e9, "traffic": (traffic or {}).get("decode")},
            "round_trip": {"ms": enc_ms + dec_ms, "achieved": rt_gbs, "frac": rt_gbs / peak, "algorithmic_bytes": enc_alg + dec_alg},
            "decode_foreign": {"ms": foreign_ms, "achieved": for_gbs, "frac": for_gbs / peak, "gpx_per_s": px_rank / (foreign_ms * 1e-3) / 1e9,
                               "note": "same streams decoded WITHOUT the restart index: the path qoi_decode / qoi_read take (cuda/qoi_fdec.cuh)"},
        }
        cpu_baseline = None
        try:
            os.sched_setaffinity(0, full_affinity)  # the CPU codec legs get every host thread again
        except OSError:
            pass
        if world == 1 and not args.no_cpu:
            kind, eb, db, threads = load_cpu_codec()
            slot = W * H * (CH + 1) + 22
            # a bounded sample: one image per host thread (images are qoi.h's only parallel axis), at most the batch
            m = min(n, threads)
            keep = np.empty(m * slot, dtype=np.uint8)
            klens = np.zeros(m, dtype=np.int32)
            cpu_round_trip(eb, db, threads, h_pix, m, keep, klens, slot)  # first-touch warm-up
            te, td = cpu_round_trip(eb, db, threads, h_pix, m, keep, klens, slot)
            gpu_streams = d_streams.cpu().numpy()
            ok = all(int(lens[i]) == int(klens[i]) and
                     np.array_equal(keep[i * slot:i * slot + int(klens[i])], gpu_streams[int(plan.stream_off[i]):int(plan.stream_off[i]) + int(klens[i])])
                     for i in range(m))
            if not ok:
                raise SystemExit("GPU streams differ from the CPU codec's")
            # column A of BASELINE.md section 3: one thread, one image
            k1 = np.empty(slot, dtype=np.uint8)
            l1 = np.zeros(1, dtype=np.int32)
            cpu_round_trip(eb, db, 1, h_pix, 1, k1, l1, slot)
            te1, td1 = cpu_round_trip(eb, db, 1, h_pix, 1, k1, l1, slot)
            cpu_baseline = {"value": m * W * H / (te + td) / 1e9, "unit": "Gpx/s", "cores": min(threads, m), "kind": kind,
                            "sample": f"{m} x {W}x{H}x{CH} images (one per host thread, at most the batch), one encode+decode pass after a warm-up pass, "
                                      "OpenMP over images, malloc+free inside the call (qoibench.c:446-450,471-481); GPU streams of these images verified identical",
                            "encode_gpx": m * W * H / te / 1e9, "decode_gpx": m * W * H / td / 1e9,
                            "single_thread": {"value": W * H / (te1 + td1) / 1e9, "encode_gpx": W * H / te1 / 1e9, "decode_gpx": W * H / td1 / 1e9,
                                              "sample": "one image, one thread"}}
            all_threads_sample(cpu_baseline, eb, db, threads, h_pix, n, m, img_bytes, slot)
        try:
            c_api = c_api_single_image(h_pix, args) if (world == 1 and not args.no_cpu) else None
        except Exception as e:  # the headline line must not depend on this extra
            c_api = {"error": repr(e)}
        pcie = pcie_ceiling(torch) if (args.e2e_steps > 0 and not args.no_pcie) else None
        knobs = {k: v for k, v in os.environ.items() if k.startswith("QOI_GPU_")}
        e2e = None
        if e2e_s is not None:
            e2e = {"value": px_all / (e2e_ms * 1e-3) / 1e9, "unit": "Gpx/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                   "ms_per_step": e2e_ms, "api": "qoi_gpu_encode_host + qoi_gpu_decode_host (pinned host buffers)",
                   "gbs_per_direction": h2d / (e2e_ms * 1e-3) / 1e9}
            if pcie and "both_gbs_per_direction" in pcie:
                e2e["pcie_ceiling"] = pcie
                e2e["frac_of_pcie_ceiling"] = e2e["gbs_per_direction"] / pcie["both_gbs_per_direction"]
        line = {
            "metric": METRIC if args.config == "headline" else f"Gpixels/s encode+decode round trip, {W}x{H}x{CH} batch",
            "value": value, "unit": "Gpx/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "u8", "data": "synthetic",
            "config": config_dict(args, world),
            "roofline": roofline,
            "cpu_baseline": cpu_baseline,
            "e2e": e2e,
            "gpu_launches": int(launches_all),
            "clocks": clocks,
            "extra": {"encoded_bytes_per_px": e_bpp, "numa": numa, "encoded_bytes_all_ranks": enc_bytes_all, "c_api_single_image": c_api,
                      "library": os.path.basename(qoi_b200.library_path()), "knobs": knobs,
                      "shard_check": "every rank compared one of its streams with the CPU codec" if dist_on else None},
        }
        if emit:
            print(json.dumps(line), flush=True)
    plan.close()
    ctx.close()
    del d_pix, d_streams, d_dec, h_pix, h_streams
    torch.cuda.empty_cache()
    return line


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--config", default="headline", choices=sorted(CONFIGS) + ["sweep"],
                    help="headline (default): 32 x 4096x4096 RGBA per GPU; 1080p_rgb_1024: BASELINE configs[2]; 8192_rgba_32: configs[3] per-GPU share; "
                         "sweep: configs[4], one JSON line per size x channels")
    ap.add_argument("--batch", type=int, default=0, help="images per GPU (weak scaling); 0 = the config's own")
    ap.add_argument("--dist", type=int, default=4, help="synthetic distribution id (4 = photo, headline)")
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--e2e-steps", type=int, default=3)
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--no-pcie", action="store_true", help="skip the bare-copy PCIe ceiling measurement")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "b200" else args.warmup

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))

    if args.config != "sweep":
        w, h, ch, nb = CONFIGS[args.config]
        set_workload(w, h, ch)
        if not args.batch:
            args.batch = nb
    if args.impl == "reference":
        if args.config == "sweep":
            raise SystemExit("--impl reference runs one configuration (not the sweep)")
        run_reference_arm(args, rank, world)
        return

    import torch

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the product has no CPU path")
    torch.cuda.set_device(local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    if args.config == "sweep":
        # configs[4]: sizes 64^2 .. 16384^2 x channels 3/4, batch chosen so that a point holds ~1 Gpx (capped), one line per point
        user_batch = args.batch
        for ch in (3, 4):
            for sz in SWEEP_SIZES:
                set_workload(sz, sz, ch)
                args.batch = user_batch or max(1, min(4096, (1 << 30) // (sz * sz)))
                args.config = "sweep"
                run_b200(args, rank, world, local_rank, torch, dist)
    else:
        run_b200(args, rank, world, local_rank, torch, dist)
    if dist is not None:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
