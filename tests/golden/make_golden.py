"""Generates tests/golden/kat.json by running the UNMODIFIED reference (oracle/_ref/libqoi_ref.so,
built from /root/reference/qoi.h by oracle/Makefile).  Run in the build container only:

    python tests/golden/make_golden.py

The inputs are the known-answer cases of SURVEY.md Appendix A.4/A.5 plus digests of reference
streams for the synthetic distributions, so the fixtures stay small.
"""
import hashlib
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
import qoi_testlib as q  # noqa: E402

L = q.cpu_libs()
assert L.ref is not None, "reference library missing: run `make -C oracle` where /root/reference exists"

enc_cases = [
    ("zero_pixel_hits_slot0", 4, [(0, 0, 0, 0)]),
    ("start_pixel_is_run", 4, [(0, 0, 0, 255)]),
    ("initial_run_not_inserted", 4, [(0, 0, 0, 255), (9, 9, 9, 255), (0, 0, 0, 255)]),
    ("index_after_insert", 4, [(9, 9, 9, 255), (0, 0, 0, 255), (9, 9, 9, 255), (0, 0, 0, 255)]),
    ("run_62_62_5", 3, [(77, 77, 77)] * 130),
    ("run_split_from_pixel0", 4, [(0, 0, 0, 255)] * 63),
    ("luma_luma_rgb_rgb_index", 3, [(10, 10, 10), (11, 9, 12), (20, 30, 25), (200, 1, 2), (20, 30, 25)]),
    ("rgba_run_rgba", 4, [(1, 2, 3, 4), (1, 2, 3, 4), (1, 2, 3, 9)]),
    ("diff_wraparound", 4, [(255, 255, 255, 255), (0, 0, 0, 255), (1, 1, 1, 255)]),
    ("run_exactly_62", 4, [(5, 6, 7, 255)] + [(5, 6, 7, 255)] * 62),
    ("run_124_then_change", 3, [(1, 1, 1)] * 125 + [(2, 2, 2)]),
    ("luma_edges", 3, [(100, 100, 100), (68 + 100 - 100, 68, 68), (107, 99, 60)]),
]

out = {"encode": [], "decode": [], "digest": [], "reject": []}
for name, ch, px in enc_cases:
    arr = np.array(px, dtype=np.uint8).reshape(-1)
    s = L.ref.encode(arr, len(px), 1, ch)
    out["encode"].append({"name": name, "channels": ch, "width": len(px), "height": 1,
                          "pixels": arr.tolist(), "stream": s.hex()})

dec_cases = [
    ("index_after_run_insert", q.header(3, 1, 4) + bytes([0xC0, 0xFE, 9, 9, 9, 0x35]) + q.END_MARK, [0, 3, 4]),
    ("truncated_repeats_last", q.header(4, 1, 3) + bytes([0xFE, 5, 6, 7]) + q.END_MARK, [0, 3, 4]),
    ("bad_end_marker_ignored", q.header(2, 1, 4) + bytes([0xFE, 1, 2, 3, 0x55]) + bytes([9] * 8), [0, 4]),
    ("rgba_op_in_rgb_file", q.header(3, 1, 3) + bytes([0xFF, 1, 2, 3, 77, 0x6A, 0xFF, 9, 8, 7, 6]) + q.END_MARK, [0, 3, 4]),
    ("index_never_written_slot", q.header(3, 1, 4) + bytes([0x11, 0x6A, 0x00]) + q.END_MARK, [0, 4]),
    ("run_longer_than_image", q.header(5, 1, 4) + bytes([0xFE, 3, 2, 1, 0xFD]) + q.END_MARK, [4]),
    ("luma_wrap", q.header(3, 1, 3) + bytes([0xFE, 250, 250, 250, 0xBF, 0xFF, 0x80, 0x00]) + q.END_MARK, [3]),
]
for name, stream, chans in dec_cases:
    for c in chans:
        px, d = L.ref.decode(stream, c)
        out["decode"].append({"name": name, "channels_arg": c, "stream": stream.hex(),
                              "desc": [d.width, d.height, d.channels, d.colorspace], "pixels": px.tolist()})

for dist in range(8):
    for (w, h) in [(64, 64), (97, 23), (256, 256)]:
        for ch in (3, 4):
            img = L.image(w, h, ch, dist, index=dist + 1)
            s = L.ref.encode(img, w, h, ch)
            out["digest"].append({"dist": dist, "w": w, "h": h, "channels": ch, "index": dist + 1,
                                  "len": len(s), "sha256": hashlib.sha256(s).hexdigest(),
                                  "pixels_sha256": hashlib.sha256(img.tobytes()).hexdigest()})

# rejection behaviour (qoi.h:364-372, 497-521): True = returns NULL
rej = []
for (w, h, ch, cs) in [(0, 1, 4, 0), (1, 0, 4, 0), (1, 1, 2, 0), (1, 1, 5, 0), (1, 1, 4, 2), (20000, 20000, 4, 0)]:
    d = q.Desc(w, h, ch, cs)
    import ctypes as C
    n = C.c_int(0)
    one = (C.c_ubyte * 16)()
    p = L.ref._enc(one, C.byref(d), C.byref(n))
    rej.append({"kind": "encode", "w": w, "h": h, "ch": ch, "cs": cs, "null": not bool(p)})
for name, stream, carg in [
    ("bad_magic", b"qoix" + q.header(1, 1, 4)[4:] + bytes([0x00]) + q.END_MARK, 0),
    ("zero_width", q.header(0, 1, 4) + bytes([0x00]) + q.END_MARK, 0),
    ("zero_height", q.header(1, 0, 4) + bytes([0x00]) + q.END_MARK, 0),
    ("channels_2", q.header(1, 1, 2) + bytes([0x00]) + q.END_MARK, 0),
    ("colorspace_2", q.header(1, 1, 4, 2) + bytes([0x00]) + q.END_MARK, 0),
    ("too_many_pixels", q.header(20000, 20000, 4) + bytes([0x00]) + q.END_MARK, 0),
    ("too_short", q.header(1, 1, 4) + bytes(7), 0),
    ("channels_arg_5", q.header(1, 1, 4) + bytes([0x00]) + q.END_MARK, 5),
    ("minimal_ok", q.header(1, 1, 4) + q.END_MARK, 0),
]:
    px, d = L.ref.decode(stream, carg)
    rej.append({"kind": "decode", "name": name, "stream": stream.hex(), "channels_arg": carg, "null": px is None,
                "desc": [d.width, d.height, d.channels, d.colorspace]})
out["reject"] = rej

with open(os.path.join(HERE, "kat.json"), "w") as f:
    json.dump(out, f, indent=1)
print("wrote", len(out["encode"]), "encode,", len(out["decode"]), "decode,", len(out["digest"]), "digest,", len(rej), "reject cases")
