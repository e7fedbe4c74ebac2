#!/usr/bin/env python
"""capi_lat.py [w h ch dist reps] -- latency of the four-function drop-in (qoi_encode / qoi_decode of libqoi_gpu.so) on one
image per call, pageable memory in, malloc'ed result out, the way qoibench.c:446-450,471-481 calls qoi.h.  MEASUREMENT TOOL.
QOI_GPU_TRACE=1 makes the library print its stage times per call."""
import ctypes as C
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests"))
import qoi_testlib as q  # noqa: E402  (synthetic images + the CPU checker; test infrastructure)


class Desc(C.Structure):
    _fields_ = [("width", C.c_uint), ("height", C.c_uint), ("channels", C.c_ubyte), ("colorspace", C.c_ubyte)]


def main():
    w, h, ch, dist, reps = [int(x) for x in (sys.argv[1:6] + ["4096", "4096", "4", "4", "8"][len(sys.argv) - 1:])]
    L = C.CDLL(os.path.join(ROOT, "qoi_b200", "libqoi_gpu.so"))
    L.qoi_encode.restype = C.c_void_p
    L.qoi_encode.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
    L.qoi_decode.restype = C.c_void_p
    L.qoi_decode.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_int]
    libc = C.CDLL(None)
    libc.free.argtypes = [C.c_void_p]
    cpu = q.cpu_libs()
    img = np.ascontiguousarray(cpu.image(w, h, ch, dist))
    d, d2, n = Desc(w, h, ch, 0), Desc(), C.c_int(0)
    sp = L.qoi_encode(img.ctypes.data, C.byref(d), C.byref(n))
    assert sp
    want = cpu.best().encode(img, w, h, ch)
    assert C.string_at(sp, n.value) == want, "stream differs from the CPU checker"
    px = L.qoi_decode(sp, n.value, C.byref(d2), ch)
    assert px and np.array_equal(np.frombuffer(C.string_at(px, w * h * ch), dtype=np.uint8), img), "round trip"
    libc.free(px)
    te, td, tf = [], [], []
    for _ in range(reps):
        t0 = time.perf_counter()
        libc.free(sp)
        t1 = time.perf_counter()
        sp = L.qoi_encode(img.ctypes.data, C.byref(d), C.byref(n))
        t2 = time.perf_counter()
        px = L.qoi_decode(sp, n.value, C.byref(d2), ch)
        t3 = time.perf_counter()
        libc.free(px)
        t4 = time.perf_counter()
        te.append(t2 - t1); td.append(t3 - t2); tf.append((t1 - t0) + (t4 - t3))
    med = lambda v: 1e3 * sorted(v)[len(v) // 2]
    print(f"{w}x{h}x{ch} dist {dist}: qoi_encode {med(te):.2f} ms  qoi_decode {med(td):.2f} ms  (free of both results {med(tf):.2f} ms)  stream {n.value} bytes, checked against the CPU reference")


if __name__ == "__main__":
    main()
