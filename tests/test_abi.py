"""The C-ABI library loads and exports every function include/*.h declares (no compute: CPU only)."""
import ctypes as C
import os
import re
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_functions(path):
    text = open(path).read()
    text = text.split("#ifdef QOI_IMPLEMENTATION")[0]        # declarations only, not the stdio bodies
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)        # comments
    text = re.sub(r"^\s*#.*$", "", text, flags=re.M)          # preprocessor lines
    return sorted(set(re.findall(r"\b(qoi_[a-z0-9_]+)\s*\(", text)))


def test_library_builds_loads_and_exports_everything():
    import qoi_b200
    from qoi_b200.build import build_library

    build_library()
    lib = C.CDLL(qoi_b200.library_path())
    names = declared_functions(os.path.join(ROOT, "include", "qoi.h")) + declared_functions(os.path.join(ROOT, "include", "qoi_gpu.h"))
    assert {"qoi_encode", "qoi_decode", "qoi_read", "qoi_write", "qoi_gpu_encode", "qoi_gpu_decode", "qoi_gpu_plan_create"} <= set(names)
    for n in names:
        assert hasattr(lib, n), f"{n} is declared in include/ but not exported"


def test_binding_covers_the_abi():
    import qoi_b200

    qoi_b200.lib()  # binds argtypes/restypes for every entry point; raises if a symbol is missing
    assert C.sizeof(qoi_b200.Desc) == 12  # reference sizeof(qoi_desc), SURVEY.md probe P1


def test_multi_device_library_exports_its_header():
    """include/qoi_gpu_multi.h: every declared entry point is exported by libqoi_gpu_nccl.so (no calls: needs GPUs + NCCL)."""
    import re
    from qoi_b200 import multi as qm

    L = qm.lib()
    text = open(os.path.join(ROOT, "include", "qoi_gpu_multi.h")).read()
    names = set(re.findall(r"\b(qoi_gpu_m[a-z_]+)\s*\(", text))
    assert {"qoi_gpu_mctx_create", "qoi_gpu_mplan_create", "qoi_gpu_mencode", "qoi_gpu_mdecode", "qoi_gpu_mplan_shard"} <= names
    for n in names:
        assert hasattr(L, n), n


def test_kernels_are_blackwell_native():
    """SASS evidence that the hot path uses bulk async copies (TMA, UBLKCP) for the pixel tiles, mbarriers, named barriers,
    warp votes, match.any (on every other row only: ~50 ADU cycles per warp on sm_100a, profiles/r2_enc_v1, r2_enc_v4), dp4a
    and 256-bit global accesses in the decoders."""
    import qoi_b200

    try:
        sass = subprocess.run(["cuobjdump", "-sass", qoi_b200.library_path()], capture_output=True, text=True, timeout=120).stdout
    except (OSError, subprocess.TimeoutExpired):
        import pytest
        pytest.skip("cuobjdump not available")
    assert "sm_100a" in sass
    for needle in ("UBLKCP.S.G", "SYNCS.", "BAR.ARV", "IDP.4A", "VOTE.ANY", "MATCH.ANY", "LDG.E.ENL2.256", "STG.E.ENL2.256"):
        assert needle in sass, needle


def test_no_product_file_touches_the_oracle():
    """only tests/, __graft_entry__.smoke and bench.py's cpu legs may reference oracle/ (task rule)."""
    for sub in ("cuda", "include", "qoi_b200"):
        for dirpath, _, files in os.walk(os.path.join(ROOT, sub)):
            for f in files:
                if f.endswith((".so", ".pyc")):
                    continue
                text = open(os.path.join(dirpath, f), errors="replace").read()
                assert "liboracle" not in text and "libqoi_ref" not in text and "ora_encode" not in text, os.path.join(dirpath, f)
