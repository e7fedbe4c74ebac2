"""qoi_b200 -- Python face of the B200-native QOI codec.

This package is a thin ctypes binding over the C ABI in ``include/qoi.h`` and ``include/qoi_gpu.h``
(library ``qoi_b200/libqoi_gpu.so``, built from ``cuda/qoi_gpu.cu`` for sm_100a).  It mirrors the
reference's interface for the hot path (reference ``qoi.h:236-241,278,289``):

    encode(pixels, desc) -> bytes          qoi_encode
    decode(stream, channels) -> (array, desc)   qoi_decode

plus the batch / device-resident API (``Context``, ``Plan``) that the benchmark and the multi-GPU
harness use.  There is no CPU implementation here: importing works anywhere, but every call needs
the CUDA library and a device, and fails loudly otherwise.
"""
from .api import (  # noqa: F401
    QOI_LINEAR,
    QOI_SRGB,
    Context,
    Desc,
    Plan,
    QoiGpuError,
    decode,
    device_count,
    encode,
    launch_count,
    lib,
    library_path,
    max_stream_size,
)
