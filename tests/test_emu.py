"""Runs the real kernel source on the CPU emulator (tests/emu) against the oracle -- CPU only.

This exercises the chained-scan / look-back logic of cuda/qoi_enc.cuh and the segment decoder of
cuda/qoi_dec.cuh on tile, warp-chunk and image boundaries with several CTAs interleaved in random
order.  It is a debugging aid for a container without a GPU, not a product path (see cuda_emu.h).
"""
import os
import subprocess
import tempfile

import numpy as np
import pytest

import qoi_testlib as q

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EMU = os.path.join(ROOT, "tests", "emu", "emu_main")


@pytest.fixture(scope="module")
def emu(cpu):
    subprocess.check_call(["make", "-s", "-C", os.path.join(ROOT, "tests", "emu")])
    return EMU


def run_enc(emu, img, w, h, ch, n, grid, seed, mis, extra_variants=()):
    with tempfile.NamedTemporaryFile(suffix=".raw") as f:
        f.write(np.ascontiguousarray(img, dtype=np.uint8).tobytes())
        f.flush()
        r = subprocess.run([emu, "enc", str(w), str(h), str(ch), str(n), str(grid), str(seed), str(mis), f.name],
                           capture_output=True, text=True, timeout=300)
        assert r.returncode == 0 and r.stdout.startswith("PASS"), r.stdout + r.stderr[-2000:]
        for extra in extra_variants:
            r = subprocess.run([emu, "enc", str(w), str(h), str(ch), str(n), str(grid), str(seed), str(mis), f.name],
                               capture_output=True, text=True, timeout=300, env=dict(os.environ, QOI_EMU_SKIP_FOREIGN="1", **extra))
            assert r.returncode == 0 and r.stdout.startswith("PASS"), str(extra) + ": " + r.stdout + r.stderr[-2000:]


@pytest.mark.parametrize("dist", range(8))
def test_emulated_distributions(emu, cpu, dist):
    run_enc(emu, cpu.batch(2, 160, 130, 4, dist), 160, 130, 4, 2, 3, 10 + dist, 0)
    run_enc(emu, cpu.batch(2, 131, 97, 3, dist), 131, 97, 3, 2, 2, 20 + dist, dist)


def test_emulated_stress(emu, cpu):
    rng = np.random.default_rng(7)
    for it in range(18):
        w = int(rng.integers(60, 300)); h = int(rng.integers(20, 120)); ch = int(rng.integers(3, 5)); n = int(rng.integers(1, 4))
        img = np.concatenate([q.stress_image(rng, w * h, ch, it % 6) for _ in range(n)])
        run_enc(emu, img, w, h, ch, n, int(rng.integers(1, 5)), it + 1, int(rng.integers(0, 17)) if it % 2 else 0)


def test_emulated_long_runs(emu, cpu):
    n = 40000
    a = np.zeros((n, 4), dtype=np.uint8); a[:, 3] = 255
    b = np.zeros((n, 4), dtype=np.uint8)
    c = np.zeros((n, 4), dtype=np.uint8); c[:, 3] = 255; c[n // 2:] = (9, 9, 9, 255); c[-5:] = (0, 0, 0, 255)
    d = np.full((n, 4), 200, dtype=np.uint8); d[12345] = (1, 2, 3, 4)
    for img in (a, b, c, d):
        run_enc(emu, img.reshape(-1), 400, 100, 4, 1, 4, 3, 0)
        run_enc(emu, np.ascontiguousarray(img[:, :3]).reshape(-1), 400, 100, 3, 1, 3, 4, 5)


def test_emulated_foreign_decode(emu, cpu):
    rng = np.random.default_rng(3)

    def run(streams, w, h, env_extra):
        recs = bytearray()
        for st in streams:
            recs += len(st).to_bytes(4, "little") + st
        with tempfile.NamedTemporaryFile(suffix=".bin") as f:
            f.write(bytes(recs)); f.flush()
            r = subprocess.run([emu, "dec", str(w), str(h), "4", str(len(streams)), "5", f.name], capture_output=True, text=True,
                               timeout=300, env=dict(os.environ, **env_extra))
        assert r.returncode == 0 and r.stdout.startswith("PASS"), str(env_extra) + r.stdout + r.stderr[-2000:]

    small = [q.foreign_stream(rng, 900, 4, w=30, h=30, truncate=(i % 4 == 0)) for i in range(40)]
    # header + end marker only, and a single op byte: no live segment / one tiny segment (qoi.h:540-587 fills with the colour)
    small += [q.header(30, 30, 4) + ops + q.END_MARK for ops in (b"", b"\x37", b"\xfe\x05\x06\x07", b"\xc5")]
    run(small, 30, 30, {})
    # streams of many segments (random op soup: the speculation rarely holds, the fix-up does the work), with
    # different numbers of speculative rounds (both parities of the pass-buffered state)
    big = [q.foreign_stream(rng, 40000, 4, w=200, h=200, truncate=(i == 2)) for i in range(4)]
    for rounds in ("1", "2", "3", "6", "9"):
        run(big, 200, 200, {"QOI_EMU_FROUNDS": rounds})


def test_emulated_many_tiles(emu, cpu):
    """Images of many tiles on few CTAs: look-back windows of many in-flight predecessors, sparse tables (palette), more
    images than CTAs and more CTAs than tiles in flight; indexed decode only (the any-stream decoder has its own tests)."""
    fast = ({},)
    with_index_only = lambda img, w, h, ch, n, grid, seed, mis: run_enc_only(emu, img, w, h, ch, n, grid, seed, mis)
    with_index_only(cpu.batch(1, 512, 512, 4, 4), 512, 512, 4, 1, 6, 7, 0)
    with_index_only(cpu.batch(1, 512, 300, 3, 6), 512, 300, 3, 1, 5, 9, 3)
    with_index_only(cpu.batch(1, 700, 400, 4, 7), 700, 400, 4, 1, 9, 2, 0)  # 137 tiles of runs only: window look-back over > 32 tiles
    for grid in (1, 2, 7):
        with_index_only(cpu.batch(5, 256, 64, 4, 6), 256, 64, 4, 5, grid, 3 + grid, 5)


def run_enc_only(emu, img, w, h, ch, n, grid, seed, mis):
    with tempfile.NamedTemporaryFile(suffix=".raw") as f:
        f.write(np.ascontiguousarray(img, dtype=np.uint8).tobytes())
        f.flush()
        r = subprocess.run([emu, "enc", str(w), str(h), str(ch), str(n), str(grid), str(seed), str(mis), f.name],
                           capture_output=True, text=True, timeout=600, env=dict(os.environ, QOI_EMU_SKIP_FOREIGN="1"))
        assert r.returncode == 0 and r.stdout.startswith("PASS"), r.stdout + r.stderr[-2000:]


def test_emulated_tile_edges(emu, cpu):
    """Shapes around the tile (2048 px) and warp-chunk (512 px) boundaries, runs crossing them, partial last tiles."""
    for k, (w, h, ch, n, dist) in enumerate([(2048, 1, 4, 3, 4), (2049, 1, 4, 3, 4), (2110, 1, 3, 2, 7), (2047, 3, 4, 2, 2), (513, 4, 4, 2, 5),
                                             (511, 5, 3, 2, 3), (4096, 1, 4, 1, 7), (4097, 1, 3, 1, 7), (1, 1, 4, 4, 4), (33, 1, 3, 3, 0), (300, 200, 3, 2, 5)]):
        run_enc_only(emu, cpu.batch(n, w, h, ch, dist), w, h, ch, n, 1 + k % 4, k + 1, (0, 5, 3)[k % 3] if ch == 4 else 1 + k % 3)


def test_emulated_anystream_settles_without_fallback(emu, cpu):
    """Regression of the round-2 convergence defects: photo image 111 of the 1920x1080 RGB set once guessed
    alpha 0 out of slot 0's initial zero and dragged it down the stream 48 segments per round -- 16 rounds were not enough and
    the stream went to the sequential fall-back (from segment 2009 of 6486).  It has to settle by itself, within a few rounds."""
    w, h, ch = 1920, 1080, 3
    img = cpu.image(w, h, ch, 4, 111)
    with tempfile.NamedTemporaryFile(suffix=".raw") as f:
        f.write(img.tobytes())
        f.flush()
        r = subprocess.run([emu, "enc", str(w), str(h), str(ch), "1", "4", "1", "0", f.name], capture_output=True, text=True, timeout=600,
                           env=dict(os.environ, QOI_EMU_FSTATS="1", QOI_EMU_FTRACE="1"))
    assert r.returncode == 0 and r.stdout.startswith("PASS"), r.stdout + r.stderr[-2000:]
    stats = [ln for ln in r.stderr.splitlines() if "fall-back from segment" in ln]
    assert stats, r.stderr[-2000:]
    words = stats[0].split()
    live, first = int(words[2]), int(words[words.index("segment") + 1])
    assert first == live, stats[0]           # nothing left for the fall-back
    walks = [int(ln.split(",")[1].split()[0]) for ln in r.stderr.splitlines() if ln.startswith("round ") and "to walk again" in ln]
    assert walks[:16].index(0) <= 5, walks   # no segment listed after at most five rounds
