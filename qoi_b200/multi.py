"""ctypes binding of the multi-device layer (include/qoi_gpu_multi.h, qoi_b200/libqoi_gpu_nccl.so).

One process, N devices: the batch is sharded by image blocks, NCCL moves pixels / lengths / streams between the GPUs.
Used by the tests and by tools/bench_multi.py; the codec itself is libqoi_gpu.so (qoi_b200.api)."""
import ctypes as C
import os

import numpy as np

from . import api

_HERE = os.path.dirname(os.path.abspath(__file__))
_lib = None


class Shard(C.Structure):
    _fields_ = [("device", C.c_int), ("first_image", C.c_int), ("n_images", C.c_int),
                ("d_pixels", C.c_void_p), ("d_streams", C.c_void_p), ("d_decoded", C.c_void_p), ("d_lens", C.c_void_p),
                ("pixel_off", C.POINTER(C.c_uint64)), ("stream_off", C.POINTER(C.c_uint64)), ("decoded_off", C.POINTER(C.c_uint64)),
                ("pixel_bytes", C.c_uint64), ("stream_bytes", C.c_uint64), ("decoded_bytes", C.c_uint64)]


class Times(C.Structure):
    _fields_ = [("scatter_ms", C.c_float), ("kernel_ms", C.c_float), ("lens_ms", C.c_float), ("gather_ms", C.c_float)]


def lib():
    global _lib
    if _lib is not None:
        return _lib
    api.lib()  # the codec library first (no CPU fallback: raises if it is missing)
    path = os.path.join(_HERE, "libqoi_gpu_nccl.so")
    if not os.path.exists(path):
        raise api.QoiGpuError("libqoi_gpu_nccl.so is not built (python -c 'import __graft_entry__ as g; g.build()')")
    L = C.CDLL(path)
    vp = C.c_void_p
    L.qoi_gpu_mctx_create.restype = C.c_int
    L.qoi_gpu_mctx_create.argtypes = [C.POINTER(vp), C.POINTER(C.c_int), C.c_int]
    L.qoi_gpu_mctx_destroy.argtypes = [vp]
    L.qoi_gpu_mctx_size.restype = C.c_int
    L.qoi_gpu_mctx_size.argtypes = [vp]
    L.qoi_gpu_mctx_last_error.restype = C.c_char_p
    L.qoi_gpu_mctx_last_error.argtypes = [vp]
    L.qoi_gpu_mplan_create.restype = C.c_int
    L.qoi_gpu_mplan_create.argtypes = [vp, C.POINTER(vp), C.c_int, C.POINTER(api.Desc), C.c_int]
    L.qoi_gpu_mplan_destroy.argtypes = [vp]
    L.qoi_gpu_mplan_shard.restype = C.c_int
    L.qoi_gpu_mplan_shard.argtypes = [vp, C.c_int, C.POINTER(Shard)]
    L.qoi_gpu_mencode.restype = C.c_int
    L.qoi_gpu_mencode.argtypes = [vp, vp, vp, vp, vp, C.POINTER(Times)]
    L.qoi_gpu_mdecode.restype = C.c_int
    L.qoi_gpu_mdecode.argtypes = [vp, C.c_int, vp, C.POINTER(Times)]
    _lib = L
    return L


class MultiContext:
    def __init__(self, devices):
        self._h = C.c_void_p()
        arr = (C.c_int * len(devices))(*devices)
        rc = lib().qoi_gpu_mctx_create(C.byref(self._h), arr, len(devices))
        if rc:
            raise api.QoiGpuError(f"qoi_gpu_mctx_create({list(devices)}) failed: rc {rc}")
        self.devices = list(devices)

    def error(self):
        return lib().qoi_gpu_mctx_last_error(self._h).decode()

    def close(self):
        if self._h:
            lib().qoi_gpu_mctx_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class MultiPlan:
    def __init__(self, mctx, descs, decoded_channels=0):
        self.mctx = mctx
        self.n = len(descs)
        self.descs = (api.Desc * self.n)(*descs)
        self._h = C.c_void_p()
        rc = lib().qoi_gpu_mplan_create(mctx._h, C.byref(self._h), self.n, self.descs, decoded_channels)
        if rc:
            raise api.QoiGpuError(f"qoi_gpu_mplan_create failed: rc {rc} {mctx.error()}")
        self.decoded_channels = decoded_channels or descs[0].channels

    def shard(self, rank):
        s = Shard()
        rc = lib().qoi_gpu_mplan_shard(self._h, rank, C.byref(s))
        if rc:
            raise api.QoiGpuError("qoi_gpu_mplan_shard failed")
        return s

    def encode(self, d_root_pixels=None, d_root_streams=None):
        """returns (lens int32[n], root offsets uint64[n], Times)"""
        lens = np.zeros(self.n, dtype=np.int32)
        offs = np.zeros(self.n, dtype=np.uint64)
        t = Times()
        rc = lib().qoi_gpu_mencode(self._h, d_root_pixels, d_root_streams, lens.ctypes.data, offs.ctypes.data, C.byref(t))
        if rc:
            raise api.QoiGpuError(f"qoi_gpu_mencode failed: rc {rc} {self.mctx.error()}")
        return lens, offs, t

    def decode(self, use_index=True, d_root_decoded=None):
        t = Times()
        rc = lib().qoi_gpu_mdecode(self._h, 1 if use_index else 0, d_root_decoded, C.byref(t))
        if rc:
            raise api.QoiGpuError(f"qoi_gpu_mdecode failed: rc {rc} {self.mctx.error()}")
        return t

    def close(self):
        if self._h:
            lib().qoi_gpu_mplan_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
