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
    # One NCCL per process: PyTorch loads the libnccl.so.2 of the nvidia-nccl wheel; a second, older copy with the same SONAME
    # loaded first would be handed to torch as well and miss symbols.  So link against that copy when it exists (same file for
    # both), else the system library.
    nccl_inc, nccl_lib = [], ["-lnccl"]
    try:
        import importlib.util
        spec = importlib.util.find_spec("nvidia.nccl")
        base = list(spec.submodule_search_locations)[0] if spec and spec.submodule_search_locations else None
        if base and os.path.exists(os.path.join(base, "lib", "libnccl.so.2")):
            nccl_inc = ["-I" + os.path.join(base, "include")]
            nccl_lib = ["-Xlinker", os.path.join(base, "lib", "libnccl.so.2"), "-Xlinker", "-rpath", "-Xlinker", os.path.join(base, "lib")]
    except Exception:
        pass
    cmd = [find_nvcc(), "-O2", "-std=c++17", "-Xcompiler", "-fPIC", "-shared", "-cudart", "shared", "-o", MULTI_LIB] + nccl_inc + \
          [os.path.join(ROOT, "cuda", "qoi_gpu_multi.cu"), "-L" + os.path.join(ROOT, "qoi_b200"), "-lqoi_gpu"] + nccl_lib + \
          ["-Xlinker", "-rpath", "-Xlinker", "$ORIGIN"]
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


def build_variant(name, defines):
    """A second build of the same library with other compile-time settings, for A/B timing with bench/enc_ab in one GPU
    session (qoi_b200/libqoi_gpu_<name>.so; never loaded by the package)."""
    out = os.path.join(ROOT, "qoi_b200", f"libqoi_gpu_{name}.so")
    cmd = [find_nvcc()] + NVCC_FLAGS + ["-D" + d for d in defines] + ["-o", out, os.path.join(ROOT, "cuda", "qoi_gpu.cu")]
    if os.path.exists("/usr/bin/g++"):
        cmd[1:1] = ["-ccbin", "/usr/bin/g++"]
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError("nvcc failed:\n" + res.stdout + res.stderr)
    return out


if __name__ == "__main__":
    print(build_library(force=True, verbose=True))
