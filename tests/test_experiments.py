"""experiments/: the patch series of candidate kernel changes still applies to the shipped sources and the patched
kernels pass the CPU emulator (CPU only; the library itself is built by experiments/build_exp.py for GPU A/B runs)."""
import os
import subprocess
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_patch_series_applies_and_passes_the_emulator(cpu):
    sys.path.insert(0, os.path.join(ROOT, "experiments"))
    import build_exp

    build_exp.prepare()  # raises if a patch of experiments/series no longer applies
    emu = build_exp.build_emulator()
    variants = {"encoder_earlyw.patch": [{"QOI_EMU_ENC": "earlyw"}],
                # the shipped encoder with mid-tile restart records, decoded with two lanes per tile
                "encoder_mid.patch": [{"QOI_EMU_ENC": "mid", "QOI_EMU_DEC": "half"}, {"QOI_EMU_ENC": "earlyw-mid", "QOI_EMU_DEC": "half"}]}
    assert set(build_exp.series()) == set(variants), "a patch in experiments/series has no emulator check here"
    for name in build_exp.series():
        for extra in variants[name]:
            for (w, h, ch, n, grid, seed, mis, dist) in [(160, 130, 4, 2, 3, 11, 0, 4), (131, 97, 3, 2, 2, 23, 3, 3), (400, 100, 4, 1, 4, 3, 0, 7),
                                                         (2049, 1, 4, 3, 2, 5, 0, 4), (2047, 3, 4, 2, 3, 6, 5, 2)]:
                img = cpu.batch(n, w, h, ch, dist)
                with tempfile.NamedTemporaryFile(suffix=".raw") as f:
                    f.write(np.ascontiguousarray(img).tobytes())
                    f.flush()
                    r = subprocess.run([emu, "enc", str(w), str(h), str(ch), str(n), str(grid), str(seed), str(mis), f.name],
                                       capture_output=True, text=True, timeout=300, env=dict(os.environ, QOI_EMU_SKIP_FOREIGN="1", **extra))
                assert r.returncode == 0 and r.stdout.startswith("PASS"), name + str(extra) + ": " + r.stdout + r.stderr[-2000:]


def test_library_override_is_explicit():
    """QOI_GPU_LIB points the binding at another build of the same ABI; unset, the shipped library is used."""
    code = "import qoi_b200; print(qoi_b200.library_path())"
    out = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, cwd=ROOT,
                         env={k: v for k, v in os.environ.items() if k != "QOI_GPU_LIB"}).stdout.strip()
    assert out == os.path.join(ROOT, "qoi_b200", "libqoi_gpu.so")
    out = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, cwd=ROOT,
                         env=dict(os.environ, QOI_GPU_LIB="/tmp/other.so")).stdout.strip()
    assert out == "/tmp/other.so"
