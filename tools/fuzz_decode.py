#!/usr/bin/env python
"""fuzz_decode.py <count> [seed] [max_body_bytes] -- the qoifuzz.c:20-31 contract on the GPU path.

Random bytes behind a valid header (random w, h, channels, colorspace), random `channels` argument: qoi_decode of the
library must never fault and must return exactly what the reference decoder (the CPU checker) returns -- pixels for
accepted streams, NULL for rejected ones.  Run plain for parity over many streams, and under compute-sanitizer
(memcheck / racecheck) on a smaller count for memory safety:
    compute-sanitizer --tool memcheck python tools/fuzz_decode.py 400 1 65536
TEST TOOL: imports the checker from tests/qoi_testlib (never part of the product)."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import qoi_testlib as q  # noqa: E402

import qoi_b200  # noqa: E402


def main():
    count = int(sys.argv[1]) if len(sys.argv) > 1 else 1000
    seed = int(sys.argv[2]) if len(sys.argv) > 2 else 1
    max_body = int(sys.argv[3]) if len(sys.argv) > 3 else 65536
    rng = np.random.default_rng(seed)
    chk = q.cpu_libs().best()
    accepted = rejected = 0
    for it in range(count):
        kind = it % 8
        nbody = int(rng.integers(0, max_body + 1)) if kind else int(rng.integers(0, 64))
        body = rng.integers(0, 256, size=nbody, dtype=np.uint8)
        if kind in (2, 3):      # bias towards long runs and index ops: the paths random bytes rarely reach
            body[rng.random(nbody) < 0.5] = 0xC0 | int(rng.integers(0, 62))
        elif kind == 4:
            body[:] = rng.integers(0, 64, size=nbody, dtype=np.uint8)
        w = int(rng.integers(1, 400)); h = int(rng.integers(1, 400))
        if kind == 5:           # far more pixels than the ops can fill: truncation fill (qoi.h:544)
            w, h = int(rng.integers(1000, 3000)), int(rng.integers(100, 700))
        ch_hdr = int(rng.integers(3, 5)) if kind != 6 else int(rng.integers(0, 8))   # kind 6: headers the reference rejects
        cs = int(rng.integers(0, 2)) if kind != 6 else int(rng.integers(0, 4))
        s = q.header(w, h, ch_hdr, cs) + body.tobytes()
        if kind == 7:
            s = s[:int(rng.integers(0, 30))]   # shorter than header + padding
        want_ch = int(rng.choice([0, 3, 4, 4, 4, 5 if kind == 6 else 4]))
        pa, da = chk.decode(s, want_ch)
        pb, db = qoi_b200.decode(s, want_ch)
        if (pa is None) != (pb is None):
            print(f"FAIL stream {it}: reference {'rejects' if pa is None else 'accepts'}, GPU {'rejects' if pb is None else 'accepts'}")
            return 1
        if pa is None:
            rejected += 1
            continue
        accepted += 1
        if not (pa == pb).all() or bytes(da) != bytes(db):
            bad = int(np.flatnonzero(pa != pb)[0]) if pa.shape == pb.shape else -1
            print(f"FAIL stream {it} ({w}x{h} hdr ch {ch_hdr}, arg {want_ch}, {nbody} body bytes): first differing byte {bad}")
            return 1
    print(f"fuzz ok: {count} streams (seed {seed}, bodies up to {max_body} bytes): {accepted} decoded identically, {rejected} rejected by both")
    return 0


if __name__ == "__main__":
    sys.exit(main())
