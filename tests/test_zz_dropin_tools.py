"""The reference's own tools, UNCHANGED, built against include/qoi.h + libqoi_gpu.so.

qoibench.c / qoiconv.c need <png.h>, stb_image.h and stb_image_write.h for their PNG I/O (qoibench.c:17-25,
qoiconv.c:23-26); none is in the image, so dropin/compat/ carries zlib-backed stand-ins with the same entry points.

(The file name sorts last on purpose: the GPU tests below drive separate processes and have not run on a box yet, so the
parity tests of test_gpu_parity.py come first under `pytest -x`.)

CPU part (here): the stand-ins are checked against independently computed pixels (numpy + a tiny PNG writer in this
file, and Pillow both ways) for every colour type / bit depth / interlace mode, and the two tools are built from the
byte-identical reference sources against the REFERENCE qoi.h (dropin/_ref/, no GPU code) to show the tool + stand-in
combination converts and benchmarks correctly.  GPU part (-m gpu): the same two sources linked against
libqoi_gpu.so run on the box; their .qoi files must equal the checker's byte for byte.
"""
import os
import shutil
import subprocess
import zlib

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
DROPIN = os.path.join(ROOT, "dropin")
REF = "/root/reference"

PIL = pytest.importorskip("PIL.Image")


def make(*targets):
    subprocess.check_call(["make", "-s", "-C", DROPIN] + list(targets), stdout=subprocess.DEVNULL)


@pytest.fixture(scope="module")
def selftest():
    make("compat/selftest")
    return os.path.join(DROPIN, "compat", "selftest")


# ----------------------------------------------------------------------------- a PNG writer for the odd formats
NS = {0: 1, 2: 3, 3: 1, 4: 2, 6: 4}
A7 = [(0, 0, 8, 8), (4, 0, 8, 8), (0, 4, 4, 8), (2, 0, 4, 4), (0, 2, 2, 4), (1, 0, 2, 2), (0, 1, 1, 2)]  # x0, y0, dx, dy


def chunk(kind, body=b""):
    return len(body).to_bytes(4, "big") + kind + body + (zlib.crc32(kind + body) & 0xFFFFFFFF).to_bytes(4, "big")


def pack_rows(samples, depth):
    """samples (h, w, ns) -> list of packed scanlines (bytes)."""
    h, w, ns = samples.shape
    rows = []
    for y in range(h):
        flat = samples[y].reshape(-1).astype(np.uint32)
        if depth == 16:
            rows.append(flat.astype(">u2").tobytes())
        elif depth == 8:
            rows.append(flat.astype(np.uint8).tobytes())
        else:
            per = 8 // depth
            pad = (-len(flat)) % per
            v = np.concatenate([flat, np.zeros(pad, dtype=np.uint32)]).reshape(-1, per)
            shifts = np.arange(per - 1, -1, -1, dtype=np.uint32) * depth
            rows.append((v << shifts).sum(axis=1).astype(np.uint8).tobytes())
    return rows


def paeth(a, b, c):
    p = a + b - c
    pa, pb, pc = abs(p - a), abs(p - b), abs(p - c)
    return a if (pa <= pb and pa <= pc) else (b if pb <= pc else c)


def filter_rows(rows, bpp, rng):
    out = bytearray()
    prev = None
    for r in rows:
        f = int(rng.integers(0, 5))
        out.append(f)
        for i, v in enumerate(r):
            a = r[i - bpp] if i >= bpp else 0
            b = prev[i] if prev is not None else 0
            c = prev[i - bpp] if (prev is not None and i >= bpp) else 0
            pred = [0, a, b, (a + b) >> 1, paeth(a, b, c)][f]
            out.append((v - pred) & 255)
        prev = r
    return bytes(out)


def write_png(path, samples, depth, ctype, rng, interlace=False, palette=None, trns=None, extra=()):
    h, w, ns = samples.shape
    assert ns == NS[ctype]
    bpp = max(1, depth * ns // 8)
    if interlace:
        raw = b""
        for (x0, y0, dx, dy) in A7:
            sub = samples[y0::dy, x0::dx]
            if sub.shape[0] and sub.shape[1]:
                raw += filter_rows(pack_rows(sub, depth), bpp, rng)
    else:
        raw = filter_rows(pack_rows(samples, depth), bpp, rng)
    z = zlib.compress(raw, 6)
    cut = max(1, len(z) // 3)
    data = b"\x89PNG\r\n\x1a\n" + chunk(b"IHDR", w.to_bytes(4, "big") + h.to_bytes(4, "big") + bytes([depth, ctype, 0, 0, 1 if interlace else 0]))
    for e in extra:
        data += e
    if palette is not None:
        data += chunk(b"PLTE", palette.astype(np.uint8).tobytes())
    if trns is not None:
        data += chunk(b"tRNS", trns)
    data += chunk(b"IDAT", z[:cut]) + chunk(b"IDAT", z[cut:2 * cut]) + chunk(b"IDAT", z[2 * cut:])  # split IDAT on purpose
    data += chunk(b"tEXt", b"Comment\0after the image data") + chunk(b"IEND")
    with open(path, "wb") as f:
        f.write(data)


def expected_rgba(samples, depth, ctype, palette=None, trns=None, scale_small_grey=True):
    """8-bit RGBA + the channel count stb_image reports for the file."""
    h, w, ns = samples.shape
    s = samples.astype(np.uint32)
    out = np.empty((h, w, 4), dtype=np.uint8)
    out[..., 3] = 255
    if ctype == 3:
        out[..., :3] = palette[s[..., 0]]
        nc = 3
        if trns is not None:
            alpha = np.full(256, 255, dtype=np.uint8)
            alpha[: len(trns)] = np.frombuffer(trns, dtype=np.uint8)
            out[..., 3] = alpha[s[..., 0]]
            nc = 4
        return out, nc
    if depth == 16:
        v = s >> 8
    elif depth < 8 and (scale_small_grey or trns is not None):
        v = s * 255 // ((1 << depth) - 1)
    else:
        v = s
    v = v.astype(np.uint8)
    if ctype in (0, 4):
        out[..., 0] = out[..., 1] = out[..., 2] = v[..., 0]
        if ctype == 4:
            out[..., 3] = v[..., 1]
        nc = ns
    else:
        out[..., :3] = v[..., :3]
        if ctype == 6:
            out[..., 3] = v[..., 3]
        nc = ns
    if trns is not None:
        key = np.frombuffer(trns, dtype=">u2").astype(np.uint32)
        out[..., 3] = np.where((s[..., : len(key)] == key).all(axis=2), 0, 255)
        nc += 1
    return out, nc


CASES = []
for _ct, _depths in ((0, (1, 2, 4, 8, 16)), (2, (8, 16)), (3, (1, 2, 4, 8)), (4, (8, 16)), (6, (8, 16))):
    for _d in _depths:
        for _il in (False, True):
            for _tr in ((False, True) if _ct in (0, 2, 3) else (False,)):
                CASES.append((_ct, _d, _il, _tr))


@pytest.mark.parametrize("ctype,depth,interlace,with_trns", CASES)
def test_standins_decode_every_png_format(tmp_path, selftest, ctype, depth, interlace, with_trns):
    rng = np.random.default_rng(ctype * 1000 + depth * 10 + interlace * 2 + with_trns)
    w, h = int(rng.integers(1, 40)), int(rng.integers(1, 30))
    ns = NS[ctype]
    samples = rng.integers(0, 1 << depth, size=(h, w, ns), dtype=np.uint32)
    if depth == 16:
        samples[rng.random((h, w)) < 0.3] = 0x1234  # make the colour key hit
    palette, trns = None, None
    if ctype == 3:
        palette = rng.integers(0, 256, size=(1 << depth, 3), dtype=np.uint8)
        if with_trns:
            trns = bytes(rng.integers(0, 256, size=max(1, (1 << depth) // 2), dtype=np.uint8))
    elif with_trns:
        key = samples[h // 2, w // 2]
        trns = b"".join(int(k).to_bytes(2, "big") for k in key)
    path = str(tmp_path / "t.png")
    extra = (chunk(b"gAMA", (45455).to_bytes(4, "big")), chunk(b"iCCP", b"x\0\0" + zlib.compress(b"not a profile")))
    write_png(path, samples, depth, ctype, rng, interlace, palette, trns, extra)

    want, nc = expected_rgba(samples, depth, ctype, palette, trns)
    out = subprocess.run([selftest, "decode", path, str(tmp_path / "a.raw")], capture_output=True, text=True)
    assert out.returncode == 0, out.stderr
    gw, gh, gch, used = map(int, out.stdout.split())
    assert (gw, gh, gch) == (w, h, nc)
    got = np.fromfile(tmp_path / "a.raw", dtype=np.uint8).reshape(h, w, used)
    assert (got == want[..., :used]).all()

    # the libpng call sequence of qoibench.c:183-251: grey below 8 bits stays unscaled unless a tRNS chunk makes it expand
    want_png, _ = expected_rgba(samples, depth, ctype, palette, trns, scale_small_grey=False)
    out = subprocess.run([selftest, "pngdecode", path, str(tmp_path / "b.raw")], capture_output=True, text=True)
    assert out.returncode == 0, out.stderr
    got = np.fromfile(tmp_path / "b.raw", dtype=np.uint8).reshape(h, w, 4)
    assert (got == want_png).all()

    # Pillow agrees on the formats it converts without ambiguity
    if depth == 8 and not with_trns:
        pil = np.asarray(PIL.open(path).convert("RGBA"))
        assert (pil == want).all()


def test_standins_reject_damaged_files(tmp_path, selftest):
    rng = np.random.default_rng(5)
    samples = rng.integers(0, 256, size=(9, 11, 3), dtype=np.uint32)
    path = str(tmp_path / "ok.png")
    write_png(path, samples, 8, 2, rng)
    good = open(path, "rb").read()
    assert subprocess.run([selftest, "decode", path, str(tmp_path / "o.raw")]).returncode == 0
    bad = {
        "magic": b"\x89PNX" + good[4:],
        "crc": good[:40] + bytes([good[40] ^ 1]) + good[41:],
        "short": good[: len(good) // 2],
        "empty": b"",
    }
    for name, blob in bad.items():
        p = str(tmp_path / (name + ".png"))
        open(p, "wb").write(blob)
        for mode in ("decode", "pngdecode"):
            if mode == "pngdecode" and name in ("short", "empty"):
                continue  # the caller's read callback aborts the process, as qoibench.c:170-177 does
            r = subprocess.run([selftest, mode, p, str(tmp_path / "o.raw")], capture_output=True)
            assert r.returncode == 1, (name, mode)


@pytest.mark.parametrize("ch", [3, 4])
def test_standins_encode_is_read_back_by_pillow(tmp_path, selftest, cpu, ch):
    w, h = 67, 45
    px = cpu.image(w, h, ch, 4)
    px.tofile(tmp_path / "in.raw")
    a, b = str(tmp_path / "stb.png"), str(tmp_path / "libpng.png")
    subprocess.check_call([selftest, "encode", str(tmp_path / "in.raw"), str(w), str(h), str(ch), a, b])
    for p in (a, b):
        im = PIL.open(p)
        assert im.mode == ("RGB" if ch == 3 else "RGBA") and im.size == (w, h)
        assert (np.asarray(im).reshape(-1) == px).all()
        assert os.path.getsize(p) < px.size  # filters + deflate did something (qoibench.c:108 needs size < raw)


# ----------------------------------------------------------------------------- the unchanged tools, CPU build
def png_corpus(cpu, directory):
    """A directory like the reference's benchmark suite: RGB and RGBA files, a sub-directory, a grey and a palette file."""
    os.makedirs(os.path.join(directory, "sub"), exist_ok=True)
    items = {}
    for i, (w, h, ch, dist) in enumerate([(64, 64, 4, 4), (97, 33, 3, 3), (128, 50, 4, 5), (31, 77, 3, 6), (256, 256, 4, 4)]):
        px = cpu.image(w, h, ch, dist, i).reshape(h, w, ch)
        name = os.path.join(directory, "sub" if i == 3 else "", f"img{i}.png")
        PIL.fromarray(px, "RGB" if ch == 3 else "RGBA").save(name)
        items[name] = (w, h, ch, px.reshape(-1))
    g = (np.add.outer(np.arange(40), np.arange(52)) * 3 % 256).astype(np.uint8)
    name = os.path.join(directory, "grey.png")
    PIL.fromarray(g, "L").save(name)
    items[name] = (52, 40, 4, np.stack([g, g, g, np.full_like(g, 255)], axis=2).reshape(-1))  # tools force non-RGB to RGBA
    pal = PIL.fromarray(g % 7, "P")
    pal.putpalette([(i * 37) % 256 for i in range(21)])
    name = os.path.join(directory, "pal.png")
    pal.save(name)
    items[name] = (52, 40, 3, np.asarray(pal.convert("RGB")).reshape(-1))
    return items


def run_tools_and_check(cpu, tmp_path, qoiconv, qoibench):
    chk = cpu.best()
    items = png_corpus(cpu, str(tmp_path / "images"))
    for name, (w, h, ch, px) in items.items():
        qoi = name[:-4] + ".qoi"
        r = subprocess.run([qoiconv, name, qoi], capture_output=True, text=True)
        assert r.returncode == 0, r.stdout + r.stderr
        assert open(qoi, "rb").read() == chk.encode(px, w, h, ch), name  # qoi_write -> qoi_encode, bit exact
        back_png = str(tmp_path / "back.png")
        r = subprocess.run([qoiconv, qoi, back_png], capture_output=True, text=True)  # qoi_read -> qoi_decode
        assert r.returncode == 0, r.stdout + r.stderr
        im = PIL.open(back_png)
        assert im.size == (w, h) and (np.asarray(im).reshape(-1) == px).all(), name
        os.remove(back_png)  # keep the benchmark directory to the original files
    r = subprocess.run([qoiconv, str(tmp_path / "missing.png"), str(tmp_path / "x.qoi")], capture_output=True, text=True)
    assert r.returncode == 1 and "Couldn't" in r.stdout
    # qoibench <runs> <dir>: verifies every image's round trip itself (qoibench.c:410-417) and aborts on a mismatch
    r = subprocess.run([qoibench, "2", str(tmp_path / "images")], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert "# Grand total for" in r.stdout and r.stdout.count("qoi:") >= len(items) + 2
    assert "libpng:" in r.stdout and "stbi:" in r.stdout
    r2 = subprocess.run([qoibench, "1", str(tmp_path / "images"), "--nopng", "--onlytotals"], capture_output=True, text=True, timeout=600)
    assert r2.returncode == 0 and "qoi:" in r2.stdout and "libpng:" not in r2.stdout and "## " + str(tmp_path) not in r2.stdout.replace("## Total", "")
    return r.stdout


@pytest.mark.skipif(not os.path.exists(os.path.join(REF, "qoibench.c")), reason="reference sources not present")
def test_unchanged_tools_with_standins_cpu_reference_build(tmp_path, cpu):
    make("ref-tools")
    out = run_tools_and_check(cpu, tmp_path, os.path.join(DROPIN, "_ref", "qoiconv_ref"), os.path.join(DROPIN, "_ref", "qoibench_ref"))
    assert "decode ms" in out


@pytest.mark.skipif(not os.path.exists(os.path.join(REF, "qoibench.c")), reason="reference sources not present")
def test_unchanged_tools_link_against_the_gpu_library():
    """Build check only (no device here): sources byte-identical to the reference, codec symbols undefined in the tool
    and resolved from libqoi_gpu.so, stdio wrappers from include/qoi.h."""
    import __graft_entry__ as g

    g.build()
    make("qoibench", "qoiconv")
    for tool in ("qoibench", "qoiconv"):
        src = os.path.join(DROPIN, tool + ".c")
        assert os.path.islink(src) and open(src, "rb").read() == open(os.path.join(REF, tool + ".c"), "rb").read()
        nm = subprocess.run(["nm", "-D", os.path.join(DROPIN, tool)], capture_output=True, text=True).stdout
        assert " U qoi_encode" in nm or " U qoi_decode" in nm or " U qoi_read" in nm
        ldd = subprocess.run(["ldd", os.path.join(DROPIN, tool)], capture_output=True, text=True).stdout
        assert "libqoi_gpu.so" in ldd and "not found" not in ldd.split("libqoi_gpu.so")[1].split("\n")[0]
    assert open(os.path.join(DROPIN, "qoi.h"), "rb").read() == open(os.path.join(ROOT, "include", "qoi.h"), "rb").read()


# ----------------------------------------------------------------------------- the unchanged tools on the GPU
@pytest.mark.gpu
def test_unchanged_tools_on_gpu(tmp_path, cpu):
    qoiconv, qoibench = os.path.join(DROPIN, "qoiconv"), os.path.join(DROPIN, "qoibench")
    if not (os.path.exists(qoiconv) and os.path.exists(qoibench)):
        pytest.skip("dropin/qoibench and dropin/qoiconv are built where the reference sources are (build()); not present")
    out = run_tools_and_check(cpu, tmp_path, qoiconv, qoibench)
    assert "decode ms" in out

# ----------------------------------------------------------------------------- qoiconv for directories (batch API)
def check_dir_tool(cpu, tmp_path, env):
    """dropin/qoiconv_dir: png -> qoi for a directory in several batches, bit exact; qoi -> png back, pixel exact."""
    chk = cpu.best()
    items = png_corpus(cpu, str(tmp_path / "images"))
    top = {k: v for k, v in items.items() if os.path.dirname(k) == str(tmp_path / "images")}  # the tool does not recurse
    out, back = tmp_path / "qoi", tmp_path / "back"
    out.mkdir(); back.mkdir()
    tool = os.path.join(DROPIN, "qoiconv_dir")
    r = subprocess.run([tool, str(tmp_path / "images"), str(out), "--budget-mpx", "0.01", "--verify"], capture_output=True, text=True, env=env, timeout=600)
    assert r.returncode == 0, r.stdout + r.stderr
    assert f"{len(top)} files" in r.stdout and " 0 failed" in r.stdout
    assert int(r.stdout.split(" files in ")[1].split()[0]) >= 3  # RGB and RGBA batches, split further by the tiny budget
    for name, (w, h, ch, px) in top.items():
        q = out / (os.path.basename(name)[:-4] + ".qoi")
        assert q.read_bytes() == chk.encode(px, w, h, ch), name
    r = subprocess.run([tool, str(out), str(back)], capture_output=True, text=True, env=env, timeout=600)
    assert r.returncode == 0 and " 0 failed" in r.stdout, r.stdout + r.stderr
    for name, (w, h, ch, px) in top.items():
        im = PIL.open(back / os.path.basename(name))
        assert im.size == (w, h) and (np.asarray(im).reshape(-1) == px).all(), name
    r = subprocess.run([tool, str(tmp_path / "nothing-here"), str(back)], capture_output=True, text=True, env=env)
    assert r.returncode == 1


def test_directory_converter_host_logic_with_cpu_double(tmp_path, cpu):
    """Batching, pinned-slot layout, file I/O and the reader/GPU/writer ring of dropin/qoiconv_dir.c, with the batch entry
    points of libqoi_gpu.so replaced (LD_PRELOAD) by a TEST DOUBLE over the CPU checker -- tests/shim/, never shipped."""
    import __graft_entry__ as g

    g.build()
    make("qoiconv_dir")
    double = os.path.join(ROOT, "tests", "shim", "libqoi_gpu_cpu_double.so")
    subprocess.check_call(["gcc", "-std=c99", "-O2", "-shared", "-fPIC", os.path.join(ROOT, "tests", "shim", "qoi_gpu_cpu_double.c"),
                           "-I" + os.path.join(ROOT, "include"), "-o", double, "-L" + os.path.join(ROOT, "oracle"), "-loracle",
                           "-Wl,-rpath," + os.path.join(ROOT, "oracle")])
    check_dir_tool(cpu, tmp_path, dict(os.environ, LD_PRELOAD=double, OMP_NUM_THREADS="4"))
    # without the double (and without a device) the tool refuses: no CPU path
    r = subprocess.run([os.path.join(DROPIN, "qoiconv_dir"), str(tmp_path / "images"), str(tmp_path / "qoi")], capture_output=True, text=True)
    assert r.returncode == 2 and "no usable CUDA device" in r.stderr


@pytest.mark.gpu
def test_directory_converter_on_gpu(tmp_path, cpu):
    if not os.path.exists(os.path.join(DROPIN, "qoiconv_dir")):
        pytest.skip("dropin/qoiconv_dir not built (build() builds it)")
    check_dir_tool(cpu, tmp_path, dict(os.environ))
