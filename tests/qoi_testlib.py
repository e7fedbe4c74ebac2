"""ctypes access to the CHECKER libraries (oracle restatement, compiled reference, input generator).

Test infrastructure only: nothing under qoi_b200/ imports this module.
"""
import ctypes as C
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_SO = os.path.join(ROOT, "oracle", "liboracle.so")
REF_SO = os.path.join(ROOT, "oracle", "_ref", "libqoi_ref.so")
SYNTH_SO = os.path.join(ROOT, "tools", "libqoi_synth.so")

DIST_NAMES = ["noise_rgba", "noise_opaque", "gradient", "grad2", "photo", "tiles32", "palette16", "const0"]


class Desc(C.Structure):
    _fields_ = [("width", C.c_uint), ("height", C.c_uint), ("channels", C.c_ubyte), ("colorspace", C.c_ubyte)]


def build_checkers():
    """(Re)build oracle/liboracle.so, tools/libqoi_synth.so and, if the reference is present, oracle/_ref."""
    subprocess.check_call(["make", "-s", "-C", os.path.join(ROOT, "oracle")], stdout=subprocess.DEVNULL)


class Codec:
    """One CPU codec behind the reference's C signatures (qoi.h:278,289)."""

    def __init__(self, lib, enc, dec, prefix):
        self.lib = lib
        self.prefix = prefix
        self._enc = getattr(lib, enc)
        self._dec = getattr(lib, dec)
        self._enc.restype = C.c_void_p
        self._enc.argtypes = [C.c_void_p, C.POINTER(Desc), C.POINTER(C.c_int)]
        self._dec.restype = C.c_void_p
        self._dec.argtypes = [C.c_void_p, C.c_int, C.POINTER(Desc), C.c_int]
        self._libc = C.CDLL(None)
        self._libc.free.argtypes = [C.c_void_p]
        eb = getattr(lib, prefix + "encode_batch")
        eb.restype = C.c_double
        eb.argtypes = [C.c_void_p, C.c_size_t, C.c_int, C.POINTER(Desc), C.c_int, C.c_void_p, C.c_size_t, C.c_void_p]
        db = getattr(lib, prefix + "decode_batch")
        db.restype = C.c_double
        db.argtypes = [C.c_void_p, C.c_size_t, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_size_t]
        self.encode_batch_raw = eb
        self.decode_batch_raw = db
        ht = getattr(lib, prefix + "host_threads")
        ht.restype = C.c_int
        self.host_threads = ht

    def encode(self, pixels, w, h, ch, colorspace=0):
        """pixels: uint8 array of w*h*ch bytes -> bytes or None."""
        pixels = np.ascontiguousarray(pixels, dtype=np.uint8)
        d = Desc(w, h, ch, colorspace)
        n = C.c_int(0)
        p = self._enc(pixels.ctypes.data, C.byref(d), C.byref(n))
        if not p:
            return None
        out = C.string_at(p, n.value)
        self._libc.free(p)
        return out

    def decode(self, stream, channels=0):
        """stream: bytes -> (uint8 array, Desc) or (None, Desc)."""
        buf = (C.c_ubyte * max(len(stream), 1)).from_buffer_copy(stream if len(stream) else b"\0")
        d = Desc(0, 0, 0, 0)
        p = self._dec(buf, len(stream), C.byref(d), channels)
        if not p:
            return None, d
        ch = channels if channels else d.channels
        n = d.width * d.height * ch
        out = np.frombuffer(C.string_at(p, n), dtype=np.uint8).copy()
        self._libc.free(p)
        return out, d


class CpuLibs:
    def __init__(self):
        if not (os.path.exists(ORACLE_SO) and os.path.exists(SYNTH_SO)):
            build_checkers()
        self.oracle = Codec(C.CDLL(ORACLE_SO), "ora_encode", "ora_decode", "ora_")
        self.ref = Codec(C.CDLL(REF_SO), "qoi_encode", "qoi_decode", "ref_") if os.path.exists(REF_SO) else None
        self.synth = C.CDLL(SYNTH_SO)
        self.synth.qoi_synth_image.argtypes = [C.c_void_p, C.c_uint, C.c_uint, C.c_int, C.c_int, C.c_uint]
        self.synth.qoi_synth_batch.argtypes = [C.c_void_p, C.c_size_t, C.c_int, C.c_uint, C.c_uint, C.c_int, C.c_int, C.c_uint]

    def best(self):
        """The strongest checker available: the compiled reference if present, else the restatement."""
        return self.ref if self.ref is not None else self.oracle

    def image(self, w, h, ch, dist, index=0):
        out = np.empty(w * h * ch, dtype=np.uint8)
        self.synth.qoi_synth_image(out.ctypes.data, w, h, ch, dist, index)
        return out

    def batch(self, n, w, h, ch, dist, first=0, out=None):
        stride = w * h * ch
        if out is None:
            out = np.empty(n * stride, dtype=np.uint8)
        self.synth.qoi_synth_batch(out.ctypes.data, stride, n, w, h, ch, dist, first)
        return out


_LIBS = None


def cpu_libs():
    global _LIBS
    if _LIBS is None:
        _LIBS = CpuLibs()
    return _LIBS


# ---------------------------------------------------------------------------------------------
# seeded "stress" images: distributions chosen to hit every encoder branch (runs crossing
# any boundary, the initial-run rule, INDEX hits at every distance, alpha changes, wrap-around)
# ---------------------------------------------------------------------------------------------
def stress_image(rng, npx, ch, kind):
    """Returns a uint8 array of npx*ch bytes."""
    if kind == 0:  # pure noise
        px = rng.integers(0, 256, size=(npx, 4), dtype=np.uint8)
    elif kind == 1:  # small palette with {0,0,0,255} and 0 bias -> INDEX + initial-run rule
        pal = rng.integers(0, 256, size=(rng.integers(2, 24), 4), dtype=np.uint8)
        pal[0] = (0, 0, 0, 255)
        pal[1] = (0, 0, 0, 0)
        px = pal[rng.integers(0, len(pal), size=npx)]
    elif kind == 2:  # small diffs and runs (DIFF / LUMA with wrap-around)
        steps = rng.integers(-3, 4, size=(npx, 4)).astype(np.int64)
        steps[rng.random(npx) < 0.5] = 0
        steps[:, 3] = 0
        start = rng.integers(0, 256, size=4)
        start[3] = 255
        px = ((np.cumsum(steps, axis=0) + start) & 255).astype(np.uint8)
    elif kind == 3:  # luma walk with rare alpha change
        g = rng.integers(-20, 21, size=npx)
        steps = np.stack([g + rng.integers(-8, 8, size=npx), g, g + rng.integers(-8, 8, size=npx), np.zeros(npx, dtype=np.int64)], axis=1)
        steps[:, 3] = np.where(rng.random(npx) < 0.02, rng.integers(-5, 6, size=npx), 0)
        px = ((np.cumsum(steps, axis=0) + 128) & 255).astype(np.uint8)
    elif kind == 4:  # long runs of a few colours, including the start colour and zero
        cols = np.array([(0, 0, 0, 255), (0, 0, 0, 0), (9, 9, 9, 255), (200, 1, 2, 7)], dtype=np.uint8)
        lens = rng.integers(1, 200, size=max(2, npx // 20))
        ids = rng.integers(0, len(cols), size=len(lens))
        px = np.repeat(cols[ids], lens, axis=0)
        if len(px) < npx:
            px = np.concatenate([px, np.repeat(cols[:1], npx - len(px), axis=0)])
        px = px[:npx]
    else:  # mixture, piecewise
        parts = []
        left = npx
        while left > 0:
            n = int(min(left, rng.integers(1, 400)))
            parts.append(stress_image(rng, n, 4, int(rng.integers(0, 5))).reshape(n, 4))
            left -= n
        px = np.concatenate(parts, axis=0)
    px = np.ascontiguousarray(px[:, :ch])
    return px.reshape(-1)


def header(w, h, ch, cs=0):
    return b"qoif" + int(w).to_bytes(4, "big") + int(h).to_bytes(4, "big") + bytes([ch, cs])


END_MARK = bytes([0, 0, 0, 0, 0, 0, 0, 1])


def foreign_stream(rng, npx_target, ch_hdr, w=None, h=None, truncate=False):
    """A syntactically valid stream the reference ENCODER would never write: random op mix, INDEX into
    any slot (also never-written ones), RGBA ops in 3-channel files, back-to-back INDEX, runs of any length."""
    ops = bytearray()
    produced = 0
    while produced < npx_target:
        k = rng.integers(0, 100)
        if k < 15:
            ops += bytes([0xFE]) + bytes(rng.integers(0, 256, size=3, dtype=np.uint8))
            produced += 1
        elif k < 25:
            ops += bytes([0xFF]) + bytes(rng.integers(0, 256, size=4, dtype=np.uint8))
            produced += 1
        elif k < 45:
            ops += bytes([int(rng.integers(0, 64))])
            produced += 1
        elif k < 65:
            ops += bytes([0x40 | int(rng.integers(0, 64))])
            produced += 1
        elif k < 85:
            ops += bytes([0x80 | int(rng.integers(0, 64)), int(rng.integers(0, 256))])
            produced += 1
        else:
            r = int(rng.integers(0, 62))
            ops += bytes([0xC0 | r])
            produced += r + 1
    if w is None:
        w, h = npx_target, 1
    if truncate:
        ops = ops[: max(1, len(ops) // 2)]
    return header(w, h, ch_hdr) + bytes(ops) + END_MARK
