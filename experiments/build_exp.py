#!/usr/bin/env python
"""Builds the EXPERIMENTAL library and emulator from a patched COPY of the kernel sources.

The shipped kernels (cuda/*.cuh -> qoi_b200/libqoi_gpu.so) stay byte-for-byte what was validated on the B200.
Candidate changes that have only been checked on the CPU emulator live here as a patch series
(experiments/series, applied in order on top of each other).  This script

  1. copies cuda/, include/ and tests/emu/ to build/exp/,
  2. applies the series (kernel sources only; the host and emulator hooks of a variant sit in the tree behind
     #ifdef QOI_EXP_<NAME>, which only this build defines),
  3. compiles   qoi_b200/libqoi_gpu_exp.so   (nvcc, sm_100a; the same flags as qoi_b200/build.py) and
                build/exp/tests/emu/emu_main  (the CPU emulator of the patched kernels).

Every variant is selected at run time by an environment variable and is OFF by default, so the experimental library
with no variable set runs the shipped algorithms.  A/B on a GPU box:

    QOI_GPU_LIB=$PWD/qoi_b200/libqoi_gpu_exp.so QOI_GPU_ENC=earlyw python bench.py --no-cpu

(tools/gpu_ab_exp.sh runs parity + bench for every variant in one gpurun call).

usage: python experiments/build_exp.py [--emu-only] [--lib-only]
"""
import os
import shutil
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EXP = os.path.join(ROOT, "build", "exp")
LIB = os.path.join(ROOT, "qoi_b200", "libqoi_gpu_exp.so")
EMU = os.path.join(EXP, "tests", "emu", "emu_main")


def series():
    with open(os.path.join(ROOT, "experiments", "series")) as f:
        return [ln.split("#")[0].strip() for ln in f if ln.split("#")[0].strip()]


def prepare():
    """Fresh patched copy of the sources under build/exp/."""
    shutil.rmtree(EXP, ignore_errors=True)
    for sub in ("cuda", "include"):
        shutil.copytree(os.path.join(ROOT, sub), os.path.join(EXP, sub))
    os.makedirs(os.path.join(EXP, "tests", "emu"))
    for f in ("cuda_emu.h", "emu_main.cpp"):
        shutil.copy(os.path.join(ROOT, "tests", "emu", f), os.path.join(EXP, "tests", "emu", f))
    for name in series():
        patch = os.path.join(ROOT, "experiments", name)
        res = subprocess.run(["patch", "-p1", "--no-backup-if-mismatch", "-d", EXP, "-i", patch], capture_output=True, text=True)
        if res.returncode != 0:
            raise RuntimeError(f"{name} does not apply:\n{res.stdout}{res.stderr}")
    return EXP


def build_emulator():
    gxx = "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else "g++"
    subprocess.check_call(["make", "-s", "-C", os.path.join(ROOT, "oracle")])
    cmd = [gxx, "-std=c++17", "-O2", "-g", "-Wall", "-Wno-unknown-pragmas", "-Wno-comment", "-Wno-unused-variable", "-Wno-unused-function",
           "-DQOI_EMU", "-DQOI_EXP_EARLYW", "-DQOI_EXP_MID", "-I" + os.path.join(EXP, "cuda"), "-I" + os.path.join(EXP, "tests", "emu"),
           os.path.join(EXP, "tests", "emu", "emu_main.cpp"), "-o", EMU,
           "-L" + os.path.join(ROOT, "oracle"), "-loracle", "-Wl,-rpath," + os.path.join(ROOT, "oracle")]
    subprocess.check_call(cmd)
    return EMU


def build_library():
    sys.path.insert(0, ROOT)
    from qoi_b200.build import NVCC_FLAGS, find_nvcc

    cmd = [find_nvcc()] + (["-ccbin", "/usr/bin/g++"] if os.path.exists("/usr/bin/g++") else []) + NVCC_FLAGS + \
          ["-DQOI_EXP_EARLYW", "-DQOI_EXP_MID", "-I" + os.path.join(EXP, "include"), "-o", LIB, os.path.join(EXP, "cuda", "qoi_gpu.cu")]
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError("nvcc failed:\n" + res.stdout + res.stderr)
    return LIB


if __name__ == "__main__":
    prepare()
    if "--lib-only" not in sys.argv:
        print("emulator:", build_emulator())
    if "--emu-only" not in sys.argv:
        print("library:", build_library())
