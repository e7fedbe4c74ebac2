"""ctypes binding of libqoi_gpu.so -- see qoi_b200/__init__.py."""
import ctypes as C
import os

import numpy as np

QOI_SRGB = 0
QOI_LINEAR = 1
TILE_PX = 2048

_HERE = os.path.dirname(os.path.abspath(__file__))
# QOI_GPU_LIB selects another build of the same C ABI (qoi_b200.build.build_variant writes libqoi_gpu_<name>.so for A/B runs);
# it is still a CUDA library with no CPU path, and bench.py records the override in its JSON line.
_LIB_PATH = os.environ.get("QOI_GPU_LIB") or os.path.join(_HERE, "libqoi_gpu.so")
_lib = None

_ERRORS = {1: "QOI_GPU_E_ARG", 2: "QOI_GPU_E_CUDA", 3: "QOI_GPU_E_NOMEM", 4: "QOI_GPU_E_STREAM"}


class QoiGpuError(RuntimeError):
    pass


class Desc(C.Structure):
    """qoi_desc (reference qoi.h:236-241)."""

    _fields_ = [("width", C.c_uint), ("height", C.c_uint), ("channels", C.c_ubyte), ("colorspace", C.c_ubyte)]

    def __repr__(self):
        return f"Desc({self.width}x{self.height}, channels={self.channels}, colorspace={self.colorspace})"


def library_path():
    return _LIB_PATH


def lib():
    """The loaded C-ABI library.  Raises if it has not been built: there is no fallback path."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(_LIB_PATH):
        raise QoiGpuError(
            f"{_LIB_PATH} is missing. Build it with `python -m qoi_b200.build` (needs nvcc). "
            "qoi_b200 has no CPU implementation.")
    L = C.CDLL(_LIB_PATH)
    vp, u64p, i32p = C.c_void_p, C.POINTER(C.c_uint64), C.POINTER(C.c_int32)
    L.qoi_encode.restype = vp
    L.qoi_encode.argtypes = [vp, C.POINTER(Desc), C.POINTER(C.c_int)]
    L.qoi_decode.restype = vp
    L.qoi_decode.argtypes = [vp, C.c_int, C.POINTER(Desc), C.c_int]
    L.qoi_gpu_device_count.restype = C.c_int
    L.qoi_gpu_ctx_create.restype = C.c_int
    L.qoi_gpu_ctx_create.argtypes = [C.POINTER(vp), C.c_int]
    L.qoi_gpu_ctx_destroy.argtypes = [vp]
    L.qoi_gpu_last_error.restype = C.c_char_p
    L.qoi_gpu_last_error.argtypes = [vp]
    L.qoi_gpu_max_stream_size.restype = C.c_size_t
    L.qoi_gpu_max_stream_size.argtypes = [C.POINTER(Desc)]
    L.qoi_gpu_plan_create.restype = C.c_int
    L.qoi_gpu_plan_create.argtypes = [vp, C.POINTER(vp), C.c_int, C.POINTER(Desc), u64p, u64p, u64p, C.c_int]
    L.qoi_gpu_plan_destroy.argtypes = [vp]
    L.qoi_gpu_plan_pixels.restype = C.c_uint64
    L.qoi_gpu_plan_pixels.argtypes = [vp]
    L.qoi_gpu_plan_tiles.restype = C.c_uint32
    L.qoi_gpu_plan_tiles.argtypes = [vp]
    L.qoi_gpu_encode.restype = C.c_int
    L.qoi_gpu_encode.argtypes = [vp, vp, vp, vp, vp]
    L.qoi_gpu_decode.restype = C.c_int
    L.qoi_gpu_decode.argtypes = [vp, vp, vp, vp, C.c_int, vp, vp]
    L.qoi_gpu_encode_host.restype = C.c_int
    L.qoi_gpu_encode_host.argtypes = [vp, vp, vp, vp]
    L.qoi_gpu_decode_host.restype = C.c_int
    L.qoi_gpu_decode_host.argtypes = [vp, vp, vp, vp, C.c_int]
    L.qoi_gpu_host_alloc.restype = vp
    L.qoi_gpu_host_alloc.argtypes = [C.c_size_t]
    L.qoi_gpu_host_free.argtypes = [vp]
    L.qoi_gpu_launch_count.restype = C.c_uint64
    L.qoi_gpu_plan_fstats.restype = C.c_int
    L.qoi_gpu_plan_fstats.argtypes = [vp, C.POINTER(C.c_uint32)]
    L.qoi_gpu_index_bytes.restype = C.c_size_t
    L.qoi_gpu_index_bytes.argtypes = [vp]
    L.qoi_gpu_index_export.restype = C.c_int
    L.qoi_gpu_index_export.argtypes = [vp, vp, C.c_size_t, vp]
    L.qoi_gpu_index_import.restype = C.c_int
    L.qoi_gpu_index_import.argtypes = [vp, vp, C.c_size_t, vp]
    _lib = L
    return L


_libc = C.CDLL(None)
_libc.free.argtypes = [C.c_void_p]


def device_count():
    return lib().qoi_gpu_device_count()


def launch_count():
    return int(lib().qoi_gpu_launch_count())


def max_stream_size(width, height, channels):
    return int(lib().qoi_gpu_max_stream_size(C.byref(Desc(width, height, channels, 0))))


def _check(rc, ctx=None, what=""):
    if rc != 0:
        detail = ""
        if ctx is not None:
            detail = lib().qoi_gpu_last_error(ctx).decode(errors="replace")
        raise QoiGpuError(f"{what} failed: {_ERRORS.get(rc, rc)} {detail}")


# ------------------------------------------------------------------ the reference's interface
def encode(pixels, desc):
    """qoi_encode (reference qoi.h:278): pixels = uint8 array / bytes of w*h*channels; returns bytes or None."""
    buf = np.ascontiguousarray(np.frombuffer(pixels, dtype=np.uint8) if isinstance(pixels, (bytes, bytearray)) else pixels,
                               dtype=np.uint8)
    # the C function trusts the pointer (it has no length argument, qoi.h:278): refuse a short buffer here
    if 3 <= desc.channels <= 4 and buf.size < desc.width * desc.height * desc.channels:
        raise ValueError(f"pixel buffer holds {buf.size} bytes, the desc needs {desc.width * desc.height * desc.channels}")
    n = C.c_int(0)
    p = lib().qoi_encode(buf.ctypes.data, C.byref(desc), C.byref(n))
    if not p:
        return None
    out = C.string_at(p, n.value)
    _libc.free(p)
    return out


def decode(stream, channels=0):
    """qoi_decode (reference qoi.h:289): returns (uint8 array or None, Desc)."""
    data = bytes(stream)
    buf = (C.c_ubyte * max(len(data), 1)).from_buffer_copy(data if data else b"\0")
    d = Desc(0, 0, 0, 0)
    p = lib().qoi_decode(buf, len(data), C.byref(d), channels)
    if not p:
        return None, d
    ch = channels if channels else d.channels
    out = np.frombuffer(C.string_at(p, d.width * d.height * ch), dtype=np.uint8).copy()
    _libc.free(p)
    return out, d


# ------------------------------------------------------------------ batch / device API
class Context:
    def __init__(self, device=0):
        self._h = C.c_void_p()
        _check(lib().qoi_gpu_ctx_create(C.byref(self._h), device), None, f"qoi_gpu_ctx_create(device={device})")
        self.device = device

    def close(self):
        if self._h:
            lib().qoi_gpu_ctx_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class Plan:
    """Geometry of a batch (include/qoi_gpu.h qoi_gpu_plan_create).  Offsets default to back-to-back packing:
    pixels at stride w*h*channels, streams at stride max_stream_size rounded up to 256, decoded at w*h*decoded_channels."""

    def __init__(self, ctx, descs, pixel_off=None, stream_off=None, decoded_off=None, decoded_channels=0):
        self.ctx = ctx
        self.n = len(descs)
        self.descs = (Desc * self.n)(*descs)
        ch = descs[0].channels
        self.decoded_channels = decoded_channels or ch
        if pixel_off is None:
            pixel_off = np.concatenate([[0], np.cumsum([d.width * d.height * d.channels for d in descs])[:-1]])
        if stream_off is None:
            sizes = [(max_stream_size(d.width, d.height, d.channels) + 255) & ~255 for d in descs]
            stream_off = np.concatenate([[0], np.cumsum(sizes)[:-1]])
        self.pixel_off = np.ascontiguousarray(pixel_off, dtype=np.uint64)
        self.stream_off = np.ascontiguousarray(stream_off, dtype=np.uint64)
        self.decoded_off = None if decoded_off is None else np.ascontiguousarray(decoded_off, dtype=np.uint64)
        if self.decoded_off is None:
            self._dec_off_eff = np.concatenate([[0], np.cumsum([d.width * d.height * self.decoded_channels for d in descs])[:-1]]).astype(np.uint64)
        else:
            self._dec_off_eff = self.decoded_off
        u64p = C.POINTER(C.c_uint64)
        self._h = C.c_void_p()
        _check(lib().qoi_gpu_plan_create(
            ctx._h, C.byref(self._h), self.n, self.descs,
            self.pixel_off.ctypes.data_as(u64p), self.stream_off.ctypes.data_as(u64p),
            None if self.decoded_off is None else self.decoded_off.ctypes.data_as(u64p), decoded_channels),
            ctx._h, "qoi_gpu_plan_create")
        self.pixel_bytes = int(max(int(o) + d.width * d.height * d.channels for o, d in zip(self.pixel_off, descs)))
        self.stream_bytes = int(max(int(o) + d.width * d.height * 5 + 22 for o, d in zip(self.stream_off, descs)))
        self.decoded_bytes = int(max(int(o) + d.width * d.height * self.decoded_channels for o, d in zip(self._dec_off_eff, descs)))
        self.pixels = int(lib().qoi_gpu_plan_pixels(self._h))
        self.tiles = int(lib().qoi_gpu_plan_tiles(self._h))

    def close(self):
        if self._h:
            lib().qoi_gpu_plan_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # device pointers are plain integers (e.g. torch.Tensor.data_ptr()); stream = cudaStream_t as int (0 = default)
    def encode(self, d_pixels, d_streams, d_lens, stream=0):
        _check(lib().qoi_gpu_encode(self._h, d_pixels, d_streams, d_lens, stream), self.ctx._h, "qoi_gpu_encode")

    def decode(self, d_streams, d_lens, d_decoded, use_index=True, d_status=None, stream=0):
        _check(lib().qoi_gpu_decode(self._h, d_streams, d_lens, d_decoded, 1 if use_index else 0, d_status, stream),
               self.ctx._h, "qoi_gpu_decode")

    def fstats(self):
        """diagnostics of the last decode(use_index=False): (walks per state round, unsettled streams per round,
        indices of the first images that went to the sequential fall-back)"""
        out = (C.c_uint32 * 128)()
        _check(lib().qoi_gpu_plan_fstats(self._h, out), self.ctx._h, "qoi_gpu_plan_fstats")
        v = list(out)
        return v[:60], v[64:124][:56], [x - 1 for x in v[120:128] if x]

    # restart index as a side-car (include/qoi_gpu.h qoi_gpu_index_*): bytes out, bytes in
    def export_index(self, stream=0):
        n = int(lib().qoi_gpu_index_bytes(self._h))
        buf = np.empty(n, dtype=np.uint8)
        _check(lib().qoi_gpu_index_export(self._h, buf.ctypes.data, n, stream), self.ctx._h, "qoi_gpu_index_export")
        return buf.tobytes()

    def import_index(self, blob, stream=0):
        buf = np.frombuffer(blob, dtype=np.uint8)
        _check(lib().qoi_gpu_index_import(self._h, buf.ctypes.data, buf.size, stream), self.ctx._h, "qoi_gpu_index_import")

    # host buffers: numpy arrays (or raw addresses)
    def encode_host(self, h_pixels, h_streams, h_lens):
        _check(lib().qoi_gpu_encode_host(self._h, _addr(h_pixels), _addr(h_streams), _addr(h_lens)), self.ctx._h, "qoi_gpu_encode_host")

    def decode_host(self, h_streams, h_lens, h_decoded, use_index=False):
        _check(lib().qoi_gpu_decode_host(self._h, _addr(h_streams), _addr(h_lens), _addr(h_decoded), 1 if use_index else 0),
               self.ctx._h, "qoi_gpu_decode_host")


def _addr(x):
    if isinstance(x, int):
        return x
    return x.ctypes.data


def pinned_array(nbytes):
    """uint8 numpy array over pinned host memory from qoi_gpu_host_alloc.  The memory is released when the array (and
    every view of it) has been garbage collected: the ctypes buffer the array is built on carries a finalizer."""
    import weakref

    p = lib().qoi_gpu_host_alloc(nbytes)
    if not p:
        raise QoiGpuError("qoi_gpu_host_alloc failed")
    buf = (C.c_ubyte * nbytes).from_address(p)
    weakref.finalize(buf, lib().qoi_gpu_host_free, p)  # numpy keeps `buf` alive as the base of the array and its views
    arr = np.frombuffer(buf, dtype=np.uint8)
    arr.flags.writeable = True
    return arr
