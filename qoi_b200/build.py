"""Builds qoi_b200/libqoi_gpu.so (the C-ABI library with the sm_100a kernels) in-tree with nvcc.

nvcc cross-compiles without a GPU, so this runs in the build container; the resulting .so travels
to the GPU box with the repository snapshot.
"""
import os
import shutil
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "qoi_b200", "libqoi_gpu.so")
SOURCES = [os.path.join(ROOT, "cuda", f) for f in
           ("qoi_gpu.cu", "qoi_enc.cuh", "qoi_dec.cuh", "qoi_fdec.cuh", "qoi_dev.cuh", "qoi_common.h", "qoi_plan.h")] + \
          [os.path.join(ROOT, "include", f) for f in ("qoi.h", "qoi_gpu.h")]

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-lineinfo", "-std=c++17",
    "-Xcompiler", "-fPIC", "-shared",
    "-cudart", "shared",
]


def find_nvcc():
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found: cannot build libqoi_gpu.so (there is no CPU fallback)")


MULTI_LIB = os.path.join(ROOT, "qoi_b200", "libqoi_gpu_nccl.so")
MULTI_SOURCES = [os.path.join(ROOT, "cuda", "qoi_gpu_multi.cu"), os.path.join(ROOT, "include", "qoi_gpu_multi.h"), os.path.join(ROOT, "include", "qoi_gpu.h")]


def build_multi(force=False):
    """libqoi_gpu_nccl.so: the multi-device layer (include/qoi_gpu_multi.h) on top of libqoi_gpu.so's public C API + NCCL.
    Host code only; a separate library so that the codec library itself does not depend on NCCL."""
    if not force and os.path.exists(MULTI_LIB) and all(os.path.getmtime(s) <= os.path.getmtime(MULTI_LIB) for s in MULTI_SOURCES + [LIB]):
        return MULTI_LIB
    cmd = [find_nvcc(), "-O2", "-std=c++17", "-Xcompiler", "-fPIC", "-shared", "-cudart", "shared", "-o", MULTI_LIB,
           os.path.join(ROOT, "cuda", "qoi_gpu_multi.cu"), "-L" + os.path.join(ROOT, "qoi_b200"), "-lqoi_gpu", "-lnccl",
           "-Xlinker", "-rpath", "-Xlinker", "$ORIGIN"]
    if os.path.exists("/usr/bin/g++"):
        cmd[1:1] = ["-ccbin", "/usr/bin/g++"]
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError("nvcc failed (multi-device layer):\n" + res.stdout + res.stderr)
    return MULTI_LIB


def needs_build():
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    return any(os.path.getmtime(s) > t for s in SOURCES)


def build_library(force=False, verbose=False):
    if not force and not needs_build():
        return LIB
    cmd = [find_nvcc()] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + \
          ["-o", LIB, os.path.join(ROOT, "cuda", "qoi_gpu.cu")]
    env = dict(os.environ)
    # the image exports CC/CXX wrappers that lack parts of the toolchain; nvcc should use the system g++
    if os.path.exists("/usr/bin/g++"):
        cmd[1:1] = ["-ccbin", "/usr/bin/g++"]
    res = subprocess.run(cmd, env=env, capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError("nvcc failed:\n" + res.stdout + res.stderr)
    if verbose:
        print(res.stderr)
    return LIB


if __name__ == "__main__":
    print(build_library(force=True, verbose=True))
