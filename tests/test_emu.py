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
        for lookup in ("walk", "summary"):  # both look-up variants of the encoder (QOI_GPU_LOOKUP)
            # ... and the foreign decoder's verify pass with one range per stream / with several
            # ... and the decoders' stream reader with warp-uniform / lane-driven refill
            env = dict(os.environ, QOI_GPU_LOOKUP=lookup, QOI_EMU_FVER_MIN_RANGE="1024" if lookup == "walk" else "3",
                       QOI_EMU_REFILL="uniform" if lookup == "walk" else "lane")
            r = subprocess.run([emu, "enc", str(w), str(h), str(ch), str(n), str(grid), str(seed), str(mis), f.name],
                               capture_output=True, text=True, timeout=300, env=env)
            assert r.returncode == 0 and r.stdout.startswith("PASS"), lookup + ": " + r.stdout + r.stderr[-2000:]
        for extra in extra_variants:  # experimental kernel variants that are off by default
            r = subprocess.run([emu, "enc", str(w), str(h), str(ch), str(n), str(grid), str(seed), str(mis), f.name],
                               capture_output=True, text=True, timeout=300, env=dict(os.environ, **extra))
            assert r.returncode == 0 and r.stdout.startswith("PASS"), str(extra) + ": " + r.stdout + r.stderr[-2000:]


@pytest.mark.parametrize("dist", range(8))
def test_emulated_distributions(emu, cpu, dist):
    ahead = ({"QOI_EMU_REFILL": "ahead"},)  # indexed decoder reading its ring one word ahead (QOI_GPU_REFILL=ahead)
    run_enc(emu, cpu.batch(2, 160, 130, 4, dist), 160, 130, 4, 2, 3, 10 + dist, 0, ahead)
    run_enc(emu, cpu.batch(2, 131, 97, 3, dist), 131, 97, 3, 2, 2, 20 + dist, dist, ahead)


# cuda/qoi_enc4.cuh (QOI_GPU_ENC=striped | striped-match | striped-earlyw); the tests below alternate between the first two ...
STRIPED = ({"QOI_EMU_ENC": "striped"}, {"QOI_EMU_ENC": "striped-match"})
PIPE = ({"QOI_EMU_ENC": "slab-pipe"}, {"QOI_EMU_ENC": "slab-pipe-ballot"})  # slab mode with the copy-out behind the next tile's index pass
COMPACT = ({"QOI_EMU_ENC": "slab4"}, {"QOI_EMU_ENC": "slab4-pipe"})  # slab mode in the compact layout (one pixel buffer, 4 CTAs per SM)
SLAB = ({"QOI_EMU_ENC": "slab"}, {"QOI_EMU_ENC": "slab-match"})  # the rolled "slab" mode of the same kernel (QOI_GPU_ENC=slab | slab-match)
STRIPED_EARLYW = ({"QOI_EMU_ENC": "striped-earlyw"},)  # ... and run this one on a subset


def run_variants(emu, img, w, h, ch, n, grid, seed, mis, variants):
    """Only the experimental variants (the default kernels are covered by run_enc)."""
    with tempfile.NamedTemporaryFile(suffix=".raw") as f:
        f.write(np.ascontiguousarray(img, dtype=np.uint8).tobytes())
        f.flush()
        for extra in variants:
            r = subprocess.run([emu, "enc", str(w), str(h), str(ch), str(n), str(grid), str(seed), str(mis), f.name],
                               capture_output=True, text=True, timeout=300, env=dict(os.environ, QOI_EMU_SKIP_FOREIGN="1", **extra))
            assert r.returncode == 0 and r.stdout.startswith("PASS"), str(extra) + ": " + r.stdout + r.stderr[-2000:]


@pytest.mark.parametrize("dist", range(8))
def test_emulated_striped_encoder_distributions(emu, cpu, dist):
    """The experimental lane-per-pixel encoder: streams bit exact, its restart index decodes, nothing written out of bounds."""
    four = STRIPED + SLAB + PIPE + COMPACT
    run_variants(emu, cpu.batch(2, 160, 130, 4, dist), 160, 130, 4, 2, 3, 10 + dist, 0, four)
    run_variants(emu, cpu.batch(2, 131, 97, 3, dist), 131, 97, 3, 2, 2, 20 + dist, dist, four)
    if True:
        run_variants(emu, cpu.batch(2, 160, 130, 4, dist), 160, 130, 4, 2, 3, 30 + dist, 0, STRIPED_EARLYW)
        run_variants(emu, cpu.batch(2, 131, 97, 3, dist), 131, 97, 3, 2, 2, 40 + dist, dist, STRIPED_EARLYW)


def test_emulated_striped_encoder_stress_and_runs(emu, cpu):
    rng = np.random.default_rng(7)
    for it in range(18):
        w = int(rng.integers(60, 300)); h = int(rng.integers(20, 120)); ch = int(rng.integers(3, 5)); n = int(rng.integers(1, 4))
        img = np.concatenate([q.stress_image(rng, w * h, ch, it % 6) for _ in range(n)])
        nine = STRIPED + SLAB + STRIPED_EARLYW + PIPE + COMPACT
        run_variants(emu, img, w, h, ch, n, int(rng.integers(1, 5)), it + 1, int(rng.integers(0, 17)) if it % 2 else 0, (nine[it % 9], nine[(it + 4) % 9], COMPACT[it % 2]))
    n = 40000  # runs that cross warp chunks, tiles and the 62 boundary; runs from the image start (qoi.h:396-399)
    a = np.zeros((n, 4), dtype=np.uint8); a[:, 3] = 255
    b = np.zeros((n, 4), dtype=np.uint8)
    c = np.zeros((n, 4), dtype=np.uint8); c[:, 3] = 255; c[n // 2:] = (9, 9, 9, 255); c[-5:] = (0, 0, 0, 255)
    d = np.full((n, 4), 200, dtype=np.uint8); d[12345] = (1, 2, 3, 4)
    for k, img in enumerate((a, b, c, d)):
        nine = STRIPED + SLAB + STRIPED_EARLYW + PIPE + COMPACT
        run_variants(emu, img.reshape(-1), 400, 100, 4, 1, 4, 3, 0, nine[k::4] + COMPACT[1:])
        run_variants(emu, np.ascontiguousarray(img[:, :3]).reshape(-1), 400, 100, 3, 1, 3, 4, 5, nine[(k + 2) % 4::4] + PIPE[:1])


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
    run(small, 30, 30, {})
    run(small, 30, 30, {"QOI_EMU_REFILL": "lane"})
    # streams of many segments (random op soup: the speculation rarely holds, the fix-up does the work), with
    # different numbers of speculative rounds (both parities of the pass-buffered state)
    big = [q.foreign_stream(rng, 40000, 4, w=200, h=200, truncate=(i == 2)) for i in range(4)]
    for rounds in ("1", "2", "3", "6", "9"):
        run(big, 200, 200, {"QOI_EMU_FROUNDS": rounds})
        run(big, 200, 200, {"QOI_EMU_FROUNDS": rounds, "QOI_EMU_FVER_MIN_RANGE": "2"})  # several verify ranges per stream


def test_emulated_many_tiles(emu, cpu):
    """512 x 512 images (64 tiles each) on 6 CTAs: look-back windows of many in-flight predecessors, for the shipped encoder
    and every variant."""
    img = cpu.batch(1, 512, 512, 4, 4)
    run_variants(emu, img, 512, 512, 4, 1, 6, 7, 0, ({"QOI_GPU_LOOKUP": "walk"}, STRIPED[0], COMPACT[1]))
    pal = cpu.batch(1, 512, 300, 3, 6)
    run_variants(emu, pal, 512, 300, 3, 1, 5, 9, 3, (SLAB[0], PIPE[1]))
    for grid in (1, 2, 7):  # the pipelined hand-offs with one CTA, with fewer CTAs than images, with more CTAs than tiles in flight
        run_variants(emu, cpu.batch(5, 256, 64, 4, 6), 256, 64, 4, 5, grid, 3 + grid, 5, PIPE[:1] + COMPACT)


def test_emulated_two_lanes_per_tile_decode(emu, cpu):
    """QOI_GPU_DEC=half: the slab-mode encoders also leave a restart point at every tile's mid point (table = merge of the first
    four warps' tables with the incoming one); the indexed decoder then runs job = 2 * tile + half.  Shapes around the mid
    point, runs crossing it, partial tiles."""
    half = lambda v: dict(v, QOI_EMU_DEC="half")
    for k, (w, h, ch, n, dist) in enumerate([(160, 130, 4, 2, 4), (131, 97, 3, 2, 3), (400, 100, 4, 1, 7), (2048, 1, 4, 3, 4), (2049, 1, 4, 3, 4),
                                             (2110, 1, 3, 2, 7), (2047, 3, 4, 2, 2), (512, 512, 4, 1, 0), (256, 64, 4, 5, 6), (300, 200, 3, 2, 5)]):
        img = cpu.batch(n, w, h, ch, dist)
        variants = (SLAB + PIPE + COMPACT)
        run_variants(emu, img, w, h, ch, n, 1 + k % 4, k + 1, (0, 5, 3)[k % 3] if ch == 4 else 1 + k % 3, (half(variants[k % 6]), half(variants[(k + 3) % 6])))
