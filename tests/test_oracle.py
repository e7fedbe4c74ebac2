"""Pins the oracle (oracle/qoi_oracle.c) -- CPU only.

1. against the committed golden vectors (tests/golden/kat.json, produced by the reference);
2. against the compiled reference itself (oracle/_ref/libqoi_ref.so) when it is present,
   on seeded stress images, all synthetic distributions and foreign / truncated streams.
"""
import ctypes as C
import hashlib
import json
import os

import numpy as np
import pytest

import qoi_testlib as q

KAT = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "kat.json")))


def test_kat_encode(cpu):
    for case in KAT["encode"]:
        px = np.array(case["pixels"], dtype=np.uint8)
        s = cpu.oracle.encode(px, case["width"], case["height"], case["channels"])
        assert s.hex() == case["stream"], case["name"]


def test_kat_decode(cpu):
    for case in KAT["decode"]:
        px, d = cpu.oracle.decode(bytes.fromhex(case["stream"]), case["channels_arg"])
        assert px is not None, case["name"]
        assert px.tolist() == case["pixels"], case["name"]
        assert [d.width, d.height, d.channels, d.colorspace] == case["desc"], case["name"]


def test_kat_digests(cpu):
    for case in KAT["digest"]:
        img = cpu.image(case["w"], case["h"], case["channels"], case["dist"], case["index"])
        assert hashlib.sha256(img.tobytes()).hexdigest() == case["pixels_sha256"]
        s = cpu.oracle.encode(img, case["w"], case["h"], case["channels"])
        assert len(s) == case["len"] and hashlib.sha256(s).hexdigest() == case["sha256"], case


def test_kat_reject(cpu):
    for case in KAT["reject"]:
        if case["kind"] == "encode":
            d = q.Desc(case["w"], case["h"], case["ch"], case["cs"])
            n = C.c_int(-7)
            one = (C.c_ubyte * 16)()
            p = cpu.oracle._enc(one, C.byref(d), C.byref(n))
            assert (not bool(p)) == case["null"], case
            if not p:
                assert n.value == -7  # out_len untouched on failure (qoi.h:484)
        else:
            px, d = cpu.oracle.decode(bytes.fromhex(case["stream"]), case["channels_arg"])
            assert (px is None) == case["null"], case["name"]
            assert [d.width, d.height, d.channels, d.colorspace] == case["desc"], case["name"]


def test_null_arguments(cpu):
    d = q.Desc(1, 1, 4, 0)
    n = C.c_int(0)
    one = (C.c_ubyte * 16)()
    for codec in [cpu.oracle] + ([cpu.ref] if cpu.ref else []):
        assert not codec._enc(None, C.byref(d), C.byref(n))
        assert not codec._enc(one, None, C.byref(n))
        assert not codec._enc(one, C.byref(d), None)
        assert not codec._dec(None, 30, C.byref(d), 0)
        assert not codec._dec(one, 30, None, 0)


needs_ref = pytest.mark.skipif(not os.path.exists(q.REF_SO), reason="compiled reference not present")


@needs_ref
def test_oracle_equals_reference_on_stress_images(cpu):
    rng = np.random.default_rng(1234)
    for it in range(1500):
        w = int(rng.integers(1, 98))
        h = int(rng.integers(1, 24))
        ch = int(rng.integers(3, 5))
        img = q.stress_image(rng, w * h, ch, it % 6)
        a = cpu.ref.encode(img, w, h, ch, it & 1)
        b = cpu.oracle.encode(img, w, h, ch, it & 1)
        assert a == b, (it, w, h, ch)
        for c in (0, 3, 4):
            pa, da = cpu.ref.decode(a, c)
            pb, db = cpu.oracle.decode(a, c)
            assert (pa == pb).all() and bytes(da) == bytes(db)


@needs_ref
def test_oracle_equals_reference_on_distributions(cpu):
    for dist in range(8):
        for ch in (3, 4):
            img = cpu.image(640, 360, ch, dist, 3)
            a = cpu.ref.encode(img, 640, 360, ch)
            assert a == cpu.oracle.encode(img, 640, 360, ch)
            pa, _ = cpu.ref.decode(a, 4)
            pb, _ = cpu.oracle.decode(a, 4)
            assert (pa == pb).all()


@needs_ref
def test_oracle_equals_reference_on_foreign_streams(cpu):
    rng = np.random.default_rng(99)
    for it in range(600):
        n = int(rng.integers(1, 700))
        s = q.foreign_stream(rng, n, int(rng.integers(3, 5)), truncate=(it % 5 == 0))
        for c in (0, 3, 4):
            pa, da = cpu.ref.decode(s, c)
            pb, db = cpu.oracle.decode(s, c)
            assert pa is not None and (pa == pb).all() and bytes(da) == bytes(db), it


@needs_ref
def test_oracle_equals_reference_on_header_only_streams(cpu):
    """a valid header with zero or one op: qoi.h:540-587 fills the image with the start colour / that op's colour"""
    for (w, h, ch) in [(30, 30, 4), (30, 30, 3), (1, 1, 4)]:
        for ops in (b"", b"\x37", b"\xfe\x05\x06\x07", b"\xc5"):
            s = q.header(w, h, ch) + ops + q.END_MARK
            for c in (0, 3, 4):
                pa, da = cpu.ref.decode(s, c)
                pb, db = cpu.oracle.decode(s, c)
                assert pa is not None and (pa == pb).all() and bytes(da) == bytes(db), (w, h, ch, ops, c)
    pa, _ = cpu.ref.decode(q.header(3, 1, 4) + q.END_MARK, 4)
    assert pa.tolist() == [0, 0, 0, 255] * 3


def test_roundtrip_configs0(cpu):
    """BASELINE.json configs[0]: 256x256 RGBA encode -> decode round trip on the CPU, bit exact (qoibench.c:410-417)."""
    img = cpu.image(256, 256, 4, 4)
    s = cpu.best().encode(img, 256, 256, 4)
    px, d = cpu.best().decode(s, 4)
    assert (px == img).all() and (d.width, d.height, d.channels) == (256, 256, 4)
