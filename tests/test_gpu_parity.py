"""Parity of the CUDA path against the oracle -- runs on the B200 box (pytest -m gpu).

Everything goes through the C ABI (include/qoi.h, include/qoi_gpu.h) via qoi_b200's ctypes binding.
The checker is the compiled reference (oracle/_ref/libqoi_ref.so) when present, else the restatement
(oracle/liboracle.so); both are pinned by tests/test_oracle.py.  Bit exact or fail.
"""
import ctypes as C
import json
import os

import numpy as np
import pytest

import qoi_testlib as q

pytestmark = pytest.mark.gpu

KAT = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "kat.json")))


@pytest.fixture(scope="module")
def gpu():
    import torch

    import qoi_b200

    assert torch.cuda.is_available(), "these tests need a CUDA device"
    assert os.path.exists(qoi_b200.library_path()), "libqoi_gpu.so missing: build() was not run"
    assert qoi_b200.device_count() >= 1
    ctx = qoi_b200.Context(0)
    yield qoi_b200, torch, ctx
    ctx.close()


def dev_u8(torch, arr):
    return torch.from_numpy(np.ascontiguousarray(arr)).cuda()


# ------------------------------------------------------------------ reference API (qoi_encode / qoi_decode)
def test_kat_through_c_api(gpu):
    qb, _, _ = gpu
    for case in KAT["encode"]:
        px = np.array(case["pixels"], dtype=np.uint8)
        s = qb.encode(px, qb.Desc(case["width"], case["height"], case["channels"], 0))
        assert s is not None and s.hex() == case["stream"], case["name"]
    for case in KAT["decode"]:
        px, d = qb.decode(bytes.fromhex(case["stream"]), case["channels_arg"])
        assert px is not None and px.tolist() == case["pixels"], case["name"]
        assert [d.width, d.height, d.channels, d.colorspace] == case["desc"], case["name"]


def test_reject_parity(gpu):
    qb, _, _ = gpu
    L = qb.lib()
    for case in KAT["reject"]:
        if case["kind"] == "encode":
            d = qb.Desc(case["w"], case["h"], case["ch"], case["cs"])
            n = C.c_int(-7)
            one = (C.c_ubyte * 16)()
            p = L.qoi_encode(one, C.byref(d), C.byref(n))
            assert (not bool(p)) == case["null"], case
            if not p:
                assert n.value == -7
        else:
            px, d = qb.decode(bytes.fromhex(case["stream"]), case["channels_arg"])
            assert (px is None) == case["null"], case["name"]
            assert [d.width, d.height, d.channels, d.colorspace] == case["desc"], case["name"]
    d = qb.Desc(1, 1, 4, 0)
    n = C.c_int(0)
    one = (C.c_ubyte * 32)()
    assert not L.qoi_encode(None, C.byref(d), C.byref(n))
    assert not L.qoi_encode(one, None, C.byref(n))
    assert not L.qoi_encode(one, C.byref(d), None)
    assert not L.qoi_decode(None, 30, C.byref(d), 0)
    assert not L.qoi_decode(one, 30, None, 0)


def test_stress_images_c_api(gpu, cpu):
    qb, _, _ = gpu
    chk = cpu.best()
    rng = np.random.default_rng(2024)
    for it in range(120):
        w = int(rng.integers(1, 400))
        h = int(rng.integers(1, 120))
        ch = int(rng.integers(3, 5))
        img = q.stress_image(rng, w * h, ch, it % 6)
        want = chk.encode(img, w, h, ch, it & 1)
        got = qb.encode(img, qb.Desc(w, h, ch, it & 1))
        assert got == want, (it, w, h, ch)
        for c in (0, 3, 4):
            pa, _ = chk.decode(want, c)
            pb, d = qb.decode(want, c)
            assert pb is not None and (pa == pb).all(), (it, c)
            assert (d.width, d.height, d.channels, d.colorspace) == (w, h, ch, it & 1)


def test_tiny_images(gpu, cpu):
    qb, _, _ = gpu
    chk = cpu.best()
    rng = np.random.default_rng(5)
    for (w, h) in [(1, 1), (2, 1), (1, 3), (61, 1), (62, 1), (63, 1), (64, 64), (4095, 1), (4096, 1), (4097, 1), (8193, 1)]:
        for ch in (3, 4):
            for kind in (1, 4):
                img = q.stress_image(rng, w * h, ch, kind)
                assert qb.encode(img, qb.Desc(w, h, ch, 0)) == chk.encode(img, w, h, ch)


def test_distributions_c_api(gpu, cpu):
    qb, _, _ = gpu
    chk = cpu.best()
    for dist in range(8):
        for ch in (3, 4):
            img = cpu.image(640, 360, ch, dist, 3)
            want = chk.encode(img, 640, 360, ch)
            assert qb.encode(img, qb.Desc(640, 360, ch, 0)) == want, (dist, ch)
            pa, _ = chk.decode(want, 4)
            pb, _ = qb.decode(want, 4)
            assert (pa == pb).all()


def test_foreign_streams(gpu, cpu):
    qb, _, _ = gpu
    chk = cpu.best()
    rng = np.random.default_rng(77)
    for it in range(150):
        n = int(rng.integers(1, 3000))
        s = q.foreign_stream(rng, n, int(rng.integers(3, 5)), truncate=(it % 5 == 0))
        for c in (0, 3, 4):
            pa, da = chk.decode(s, c)
            pb, db = qb.decode(s, c)
            assert pb is not None and (pa == pb).all() and bytes(da) == bytes(db), (it, c)
    # a valid header with no chunk bytes at all, and with a single op: every pixel is the start colour / that op's colour
    # (qoi.h:540-587 never parses past size - 8); a previous decode must not shine through
    for (w, h, ch) in [(30, 30, 4), (30, 30, 3), (7, 1, 4)]:
        qb.decode(q.foreign_stream(rng, w * h, ch, w=w, h=h), 4)  # leaves other pixels in the cached plan's output buffer
        for ops in (b"", b"\x37", b"\xfe\x05\x06\x07", b"\xc5"):
            s = q.header(w, h, ch) + ops + q.END_MARK
            for c in (0, 3, 4):
                pa, da = chk.decode(s, c)
                pb, db = qb.decode(s, c)
                assert pa is not None and pb is not None and (pa == pb).all() and bytes(da) == bytes(db), (w, h, ch, ops, c)
    # arbitrary bytes after a valid header never crash and match the reference (the qoifuzz.c contract)
    for it in range(60):
        body = bytes(rng.integers(0, 256, size=int(rng.integers(9, 400)), dtype=np.uint8))
        s = q.header(int(rng.integers(1, 40)), int(rng.integers(1, 40)), int(rng.integers(3, 5))) + body
        pa, _ = chk.decode(s, 4)
        pb, _ = qb.decode(s, 4)
        assert (pa is None) == (pb is None)
        if pa is not None:
            assert (pa == pb).all(), it


def test_config1_4096_rgba_bit_exact(gpu, cpu):
    """BASELINE.json configs[1]: single 4096x4096 RGBA encode, stream bit exact vs qoi.h."""
    qb, _, _ = gpu
    chk = cpu.best()
    for dist in (4, 0, 7):
        img = cpu.image(4096, 4096, 4, dist)
        want = chk.encode(img, 4096, 4096, 4)
        got = qb.encode(img, qb.Desc(4096, 4096, 4, 0))
        assert got == want, dist
        px, _ = qb.decode(got, 4)
        assert (px == img).all()


def test_foreign_large_streams(gpu, cpu):
    """Streams of many 4096-byte segments that the reference ENCODER would never write (op soup, INDEX into any slot,
    RGBA ops with changing alpha, truncation): the speculative decoder rarely settles these, the exact fix-up must."""
    qb, _, _ = gpu
    chk = cpu.best()
    rng = np.random.default_rng(2024)
    for it, (w, h) in enumerate([(640, 480), (1000, 333), (512, 512), (2048, 150)]):
        s = q.foreign_stream(rng, w * h, 4 if it % 2 == 0 else 3, w=w, h=h, truncate=(it == 1))
        for c in (4, 3):
            pa, da = chk.decode(s, c)
            pb, db = qb.decode(s, c)
            assert pb is not None and (pa == pb).all() and bytes(da) == bytes(db), (it, c)


def test_foreign_real_photographs(gpu, cpu):
    """Real photographic content (JPEG-decoded photographs that ship with scikit-learn, tiled 2x2), encoded by the CPU
    checker and decoded as foreign streams: INDEX ~10 %, RUN ~10 %, the op mix synthetic data does not have."""
    PIL = pytest.importorskip("PIL.Image")
    qb, _, _ = gpu
    chk = cpu.best()
    try:
        import sklearn.datasets
        base = os.path.join(os.path.dirname(sklearn.datasets.__file__), "images")
    except Exception:
        pytest.skip("scikit-learn not present")
    found = 0
    for name in ("china.jpg", "flower.jpg"):
        path = os.path.join(base, name)
        if not os.path.exists(path):
            continue
        found += 1
        for mode, ch in (("RGBA", 4), ("RGB", 3)):
            a = np.asarray(PIL.open(path).convert(mode))
            a = np.ascontiguousarray(np.tile(a, (2, 2, 1)))
            h, w = a.shape[:2]
            stream = chk.encode(a.reshape(-1), w, h, ch)
            px, d = qb.decode(stream, ch)
            assert px is not None and (px == a.reshape(-1)).all(), (name, mode)
            assert qb.encode(a.reshape(-1), qb.Desc(w, h, ch, 0)) == stream, (name, mode)
    if not found:
        pytest.skip("sample photographs not present")


# ------------------------------------------------------------------ batch / device API
def run_batch(gpu, cpu, descs_whc, dist_or_imgs, mis=0, dec_ch=4, odd_slots=False, repeat=1):
    qb, torch, ctx = gpu
    chk = cpu.best()
    descs = [qb.Desc(w, h, c, i & 1) for i, (w, h, c) in enumerate(descs_whc)]
    sizes = [w * h * c for (w, h, c) in descs_whc]
    poff = mis + np.concatenate([[0], np.cumsum(sizes)[:-1]])
    if odd_slots:
        slot = [((qb.max_stream_size(w, h, c) + 41) | 1) for (w, h, c) in descs_whc]
        soff = 5 + np.concatenate([[0], np.cumsum(slot)[:-1]])
    else:
        soff = None
    plan = qb.Plan(ctx, descs, pixel_off=poff, stream_off=soff, decoded_channels=dec_ch)
    host = np.zeros(plan.pixel_bytes, dtype=np.uint8)
    imgs = []
    for i, (w, h, c) in enumerate(descs_whc):
        img = dist_or_imgs[i] if isinstance(dist_or_imgs, list) else cpu.image(w, h, c, dist_or_imgs, i)
        imgs.append(img)
        host[int(poff[i]):int(poff[i]) + sizes[i]] = img
    d_pix = dev_u8(torch, host)
    d_streams = torch.full((plan.stream_bytes + 64,), 0xEE, dtype=torch.uint8, device="cuda")
    d_lens = torch.full((plan.n,), -1, dtype=torch.int32, device="cuda")
    d_dec = torch.full((plan.decoded_bytes + 16,), 0xAB, dtype=torch.uint8, device="cuda")
    d_status = torch.full((plan.n,), -1, dtype=torch.int32, device="cuda")
    for _ in range(repeat):
        d_streams.fill_(0xEE)
        plan.encode(d_pix.data_ptr(), d_streams.data_ptr(), d_lens.data_ptr())
        torch.cuda.synchronize()
        lens = d_lens.cpu().numpy()
        streams = d_streams.cpu().numpy()
        for i, (w, h, c) in enumerate(descs_whc):
            want = chk.encode(imgs[i], w, h, c, i & 1)
            o = int(plan.stream_off[i])
            assert lens[i] == len(want), (i, lens[i], len(want))
            assert streams[o:o + lens[i]].tobytes() == want, i
            assert streams[o + lens[i]] == 0xEE and (o == 0 or streams[o - 1] == 0xEE), "wrote outside the stream"
        for use_index in (True, False):
            d_dec.fill_(0xAB)
            plan.decode(d_streams.data_ptr(), d_lens.data_ptr(), d_dec.data_ptr(), use_index=use_index, d_status=d_status.data_ptr())
            torch.cuda.synchronize()
            assert (d_status.cpu().numpy() == 0).all()
            dec = d_dec.cpu().numpy()
            assert dec[plan.decoded_bytes] == 0xAB
            for i, (w, h, c) in enumerate(descs_whc):
                want_px, _ = chk.decode(streams[int(plan.stream_off[i]):int(plan.stream_off[i]) + lens[i]].tobytes(), dec_ch)
                o = int(plan._dec_off_eff[i])
                assert (dec[o:o + w * h * dec_ch] == want_px).all(), (i, use_index)
    plan.close()


def test_batch_rgba_aligned(gpu, cpu):
    run_batch(gpu, cpu, [(512, 256, 4)] * 6, 4, repeat=3)


def test_batch_all_distributions(gpu, cpu):
    for dist in range(8):
        run_batch(gpu, cpu, [(300, 200, 4)] * 3, dist)
        run_batch(gpu, cpu, [(300, 200, 3)] * 3, dist, dec_ch=4)


def test_batch_misaligned_and_mixed_sizes(gpu, cpu):
    run_batch(gpu, cpu, [(97, 23, 4), (1, 1, 4), (211, 87, 4), (64, 64, 4), (4096, 2, 4)], 3, mis=4, odd_slots=True)
    run_batch(gpu, cpu, [(97, 23, 4), (33, 5, 4)], 6, mis=7, odd_slots=True, dec_ch=3)
    run_batch(gpu, cpu, [(97, 23, 3), (1, 1, 3), (211, 87, 3), (64, 64, 3), (4097, 3, 3)], 4, mis=1, odd_slots=True, dec_ch=3)
    run_batch(gpu, cpu, [(129, 77, 3)] * 4, 2, mis=0, dec_ch=4)


def test_batch_stress_content(gpu, cpu):
    rng = np.random.default_rng(31337)
    for kind in range(6):
        for ch in (3, 4):
            whc = [(int(rng.integers(50, 500)), int(rng.integers(20, 200)), ch) for _ in range(4)]
            imgs = [q.stress_image(rng, w * h, ch, kind) for (w, h, _) in whc]
            run_batch(gpu, cpu, whc, imgs, dec_ch=4 if kind % 2 else ch)


def test_long_runs_across_many_tiles(gpu, cpu):
    """runs spanning dozens of tiles, the start-colour run, and the all-zero image (INDEX into an untouched slot)."""
    n = 300 * 1000
    a = np.zeros((n, 4), dtype=np.uint8); a[:, 3] = 255                      # start colour only
    b = np.zeros((n, 4), dtype=np.uint8)                                      # {0,0,0,0}
    c = np.zeros((n, 4), dtype=np.uint8); c[:, 3] = 255; c[n // 2:] = (9, 9, 9, 255); c[-5:] = (0, 0, 0, 255)
    d = np.full((n, 4), 200, dtype=np.uint8); d[100000] = (1, 2, 3, 4)
    run_batch(gpu, cpu, [(1000, 300, 4)] * 4, [a.reshape(-1), b.reshape(-1), c.reshape(-1), d.reshape(-1)])


def test_config2_small_batch_1080p_rgb(gpu, cpu):
    """BASELINE.json configs[2] shape (1920x1080 RGB, decode to 4 channels) at a batch the CPU checker finishes quickly."""
    run_batch(gpu, cpu, [(1920, 1080, 3)] * 8, 4, dec_ch=4)


def test_config2_batch_of_256_spans_waves(gpu, cpu):
    """configs[2] at a batch that fills the machine several times over (256 x 1920x1080 RGB = 259 k tiles): encode + both
    decodes on the device, round trip on every image, streams of a sample of images byte for byte against the CPU."""
    qb, torch, ctx = gpu
    chk = cpu.best()
    n, w, h = 256, 1920, 1080
    plan = qb.Plan(ctx, [qb.Desc(w, h, 3, 0)] * n, decoded_channels=3)
    host = cpu.batch(n, w, h, 3, 4)
    d_pix = dev_u8(torch, host)
    d_streams = torch.empty(plan.stream_bytes, dtype=torch.uint8, device="cuda")
    d_lens = torch.zeros(n, dtype=torch.int32, device="cuda")
    d_dec = torch.zeros(plan.decoded_bytes, dtype=torch.uint8, device="cuda")
    plan.encode(d_pix.data_ptr(), d_streams.data_ptr(), d_lens.data_ptr())
    for use_index in (True, False):
        d_dec.zero_()
        plan.decode(d_streams.data_ptr(), d_lens.data_ptr(), d_dec.data_ptr(), use_index=use_index)
        torch.cuda.synchronize()
        assert torch.equal(d_dec, d_pix), use_index
    lens = d_lens.cpu().numpy()
    for i in (0, 1, 77, 128, 255):
        want = chk.encode(host[i * w * h * 3:(i + 1) * w * h * 3], w, h, 3)
        o = int(plan.stream_off[i])
        assert lens[i] == len(want) and d_streams[o:o + len(want)].cpu().numpy().tobytes() == want, i
    plan.close()


def test_anystream_settles_without_fallback(gpu, cpu):
    """Round-2 regression: streams that are not the headline shape must settle in the state rounds too.
    64 x 1920x1080 RGB photographs with image indices 96..159 (index 111 once dragged a wrong alpha down its stream and went
    to the sequential fall-back) and one 8192x8192 RGBA image (4 x the segments of a 4096^2 one): decoded as foreign
    streams, pixels equal to the input, no image reported by qoi_gpu_plan_fstats."""
    qb, torch, ctx = gpu
    for (n, w, h, ch, first) in ((64, 1920, 1080, 3, 96), (1, 8192, 8192, 4, 23)):
        plan = qb.Plan(ctx, [qb.Desc(w, h, ch, 0)] * n, decoded_channels=ch)
        host = cpu.batch(n, w, h, ch, 4, first=first)
        d_pix = dev_u8(torch, host)
        d_streams = torch.empty(plan.stream_bytes, dtype=torch.uint8, device="cuda")
        d_lens = torch.zeros(n, dtype=torch.int32, device="cuda")
        d_dec = torch.zeros(plan.decoded_bytes, dtype=torch.uint8, device="cuda")
        plan.encode(d_pix.data_ptr(), d_streams.data_ptr(), d_lens.data_ptr())
        plan.decode(d_streams.data_ptr(), d_lens.data_ptr(), d_dec.data_ptr(), use_index=False)
        torch.cuda.synchronize()
        assert torch.equal(d_dec, d_pix), (w, h)
        walks, unsettled, fell_back = plan.fstats()
        assert fell_back == [], (w, h, fell_back, walks[:17])
        assert unsettled[8] == 0, (w, h, unsettled[:17])   # everything settled well before the 16th round
        plan.close()
        del d_pix, d_streams, d_dec


def test_index_sidecar_through_disk(gpu, cpu, tmp_path):
    """the restart index leaves the plan as a side-car, the streams and the side-car go to disk, a NEW plan
    reads both back and decodes with use_index -- same pixels as the reference decoder gives for the stored streams."""
    qb, torch, ctx = gpu
    chk = cpu.best()
    whc = [(700, 400, 4), (2048, 3, 4), (64, 64, 4), (1000, 777, 4), (1, 1, 4)]
    descs = [qb.Desc(w, h, c, 0) for (w, h, c) in whc]
    plan = qb.Plan(ctx, descs, decoded_channels=4)
    host = np.concatenate([cpu.image(w, h, c, (4, 7, 6, 3, 0)[i], i) for i, (w, h, c) in enumerate(whc)])
    d_pix = dev_u8(torch, host)
    d_streams = torch.zeros(plan.stream_bytes, dtype=torch.uint8, device="cuda")
    d_lens = torch.zeros(plan.n, dtype=torch.int32, device="cuda")
    with pytest.raises(qb.QoiGpuError):
        plan.export_index()  # nothing encoded yet
    plan.encode(d_pix.data_ptr(), d_streams.data_ptr(), d_lens.data_ptr())
    blob = plan.export_index()
    lens = d_lens.cpu().numpy()
    streams = d_streams.cpu().numpy()
    for i in range(plan.n):
        o = int(plan.stream_off[i])
        (tmp_path / f"img{i}.qoi").write_bytes(streams[o:o + lens[i]].tobytes())
    (tmp_path / "batch.qoix").write_bytes(blob)
    plan.close()
    # a fresh plan, everything from disk
    plan2 = qb.Plan(ctx, descs, decoded_channels=4)
    h_streams = np.zeros(plan2.stream_bytes, dtype=np.uint8)
    h_lens = np.zeros(plan2.n, dtype=np.int32)
    for i in range(plan2.n):
        data = (tmp_path / f"img{i}.qoi").read_bytes()
        assert data == chk.encode(host[int(plan2.pixel_off[i]):int(plan2.pixel_off[i]) + whc[i][0] * whc[i][1] * 4], *whc[i])
        o = int(plan2.stream_off[i])
        h_streams[o:o + len(data)] = np.frombuffer(data, dtype=np.uint8)
        h_lens[i] = len(data)
    d_streams2 = dev_u8(torch, h_streams)
    d_lens2 = torch.from_numpy(h_lens).cuda()
    d_dec = torch.zeros(plan2.decoded_bytes, dtype=torch.uint8, device="cuda")
    with pytest.raises(qb.QoiGpuError):
        plan2.decode(d_streams2.data_ptr(), d_lens2.data_ptr(), d_dec.data_ptr(), use_index=True)  # no index in this plan yet
    with pytest.raises(qb.QoiGpuError):
        plan2.import_index(blob[:-8])  # truncated side-car
    bad = bytearray(blob); bad[8] ^= 0xFF
    with pytest.raises(qb.QoiGpuError):
        plan2.import_index(bytes(bad))  # another tile size
    plan2.import_index((tmp_path / "batch.qoix").read_bytes())
    plan2.decode(d_streams2.data_ptr(), d_lens2.data_ptr(), d_dec.data_ptr(), use_index=True)
    torch.cuda.synchronize()
    assert torch.equal(d_dec, d_pix)
    assert len(blob) <= 0.14 * sum(w * h for (w, h, c) in whc) + 4096  # ~0.13 B/px
    plan2.close()


def test_c_api_large_image_through_staging(gpu, cpu):
    """qoi_encode / qoi_decode on an image of many staging chunks (8192 x 4096 RGBA = 128 MiB in 4 MiB pieces over the host
    threads): stream equal to the CPU's, pixels back."""
    qb, _, _ = gpu
    chk = cpu.best()
    w, h = 8192, 4096
    img = cpu.image(w, h, 4, 4)
    got = qb.encode(img, qb.Desc(w, h, 4, 0))
    assert got == chk.encode(img, w, h, 4)
    px, d = qb.decode(got, 4)
    assert px is not None and (px == img).all() and (d.width, d.height) == (w, h)
    px3, _ = qb.decode(got, 3)
    assert (px3.reshape(-1, 3) == img.reshape(-1, 4)[:, :3]).all()


def test_roundtrip_8192_device_resident(gpu, cpu):
    """BASELINE.json configs[3] image size: 8192x8192 RGBA, checked by the size-independent property
    decode(encode(x)) == x on the device, plus out_len against the CPU for one image."""
    qb, torch, ctx = gpu
    w = h = 8192
    n = 2
    plan = qb.Plan(ctx, [qb.Desc(w, h, 4, 0)] * n, decoded_channels=4)
    host = cpu.batch(n, w, h, 4, 4, first=11)
    d_pix = dev_u8(torch, host)
    d_streams = torch.empty(plan.stream_bytes, dtype=torch.uint8, device="cuda")
    d_lens = torch.zeros(n, dtype=torch.int32, device="cuda")
    d_dec = torch.zeros(plan.decoded_bytes, dtype=torch.uint8, device="cuda")
    plan.encode(d_pix.data_ptr(), d_streams.data_ptr(), d_lens.data_ptr())
    for use_index in (True, False):
        d_dec.zero_()
        plan.decode(d_streams.data_ptr(), d_lens.data_ptr(), d_dec.data_ptr(), use_index=use_index)
        torch.cuda.synchronize()
        assert torch.equal(d_dec, d_pix), use_index
    want = cpu.best().encode(host[: w * h * 4], w, h, 4)
    assert int(d_lens[0]) == len(want)
    o = int(plan.stream_off[0])
    assert d_streams[o:o + len(want)].cpu().numpy().tobytes() == want
    plan.close()


def test_host_buffer_api(gpu, cpu):
    qb, _, ctx = gpu
    chk = cpu.best()
    whc = [(320, 240, 4)] * 5
    descs = [qb.Desc(w, h, c, 0) for (w, h, c) in whc]
    plan = qb.Plan(ctx, descs, decoded_channels=4)
    host = cpu.batch(5, 320, 240, 4, 4)
    streams = np.zeros(plan.stream_bytes, dtype=np.uint8)
    lens = np.zeros(5, dtype=np.int32)
    plan.encode_host(host, streams, lens)
    for i in range(5):
        want = chk.encode(host[i * 320 * 240 * 4:(i + 1) * 320 * 240 * 4], 320, 240, 4)
        o = int(plan.stream_off[i])
        assert streams[o:o + lens[i]].tobytes() == want
    for use_index in (True, False):
        dec = np.zeros(plan.decoded_bytes, dtype=np.uint8)
        plan.decode_host(streams, lens, dec, use_index=use_index)
        assert (dec == host).all()
    plan.close()


def test_header_mismatch_is_reported(gpu, cpu):
    qb, torch, ctx = gpu
    plan = qb.Plan(ctx, [qb.Desc(64, 64, 4, 0)] * 2, decoded_channels=4)
    img = cpu.batch(2, 64, 64, 4, 4)
    d_pix = dev_u8(torch, img)
    d_streams = torch.zeros(plan.stream_bytes, dtype=torch.uint8, device="cuda")
    d_lens = torch.zeros(2, dtype=torch.int32, device="cuda")
    d_dec = torch.zeros(plan.decoded_bytes, dtype=torch.uint8, device="cuda")
    d_status = torch.zeros(2, dtype=torch.int32, device="cuda")
    plan.encode(d_pix.data_ptr(), d_streams.data_ptr(), d_lens.data_ptr())
    d_streams[int(plan.stream_off[1]) + 7] = 65  # corrupt the width of image 1
    plan.decode(d_streams.data_ptr(), d_lens.data_ptr(), d_dec.data_ptr(), use_index=False, d_status=d_status.data_ptr())
    torch.cuda.synchronize()
    assert d_status.cpu().tolist() == [0, 4]
    plan.close()


def test_size_sweep_config4(gpu, cpu):
    """BASELINE.json configs[4]: sizes 64x64 .. 16384x16384, channels 3 and 4.  Small sizes are compared byte for
    byte with the CPU; the largest ones through decode(encode(x)) == x on the device plus one full byte comparison."""
    qb, torch, ctx = gpu
    chk = cpu.best()
    for sz in (64, 128, 256, 512, 1024, 2048):
        for ch in (3, 4):
            run_batch(gpu, cpu, [(sz, sz, ch)] * 2, 4, dec_ch=4)
    for sz, ch, full in ((8192, 3, False), (16384, 4, True), (16384, 3, False)):
        plan = qb.Plan(ctx, [qb.Desc(sz, sz, ch, 0)], decoded_channels=ch)
        host = cpu.image(sz, sz, ch, 4, 5)
        d_pix = dev_u8(torch, host)
        d_streams = torch.empty(plan.stream_bytes, dtype=torch.uint8, device="cuda")
        d_lens = torch.zeros(1, dtype=torch.int32, device="cuda")
        d_dec = torch.zeros(plan.decoded_bytes, dtype=torch.uint8, device="cuda")
        plan.encode(d_pix.data_ptr(), d_streams.data_ptr(), d_lens.data_ptr())
        plan.decode(d_streams.data_ptr(), d_lens.data_ptr(), d_dec.data_ptr(), use_index=True)
        torch.cuda.synchronize()
        assert torch.equal(d_dec, d_pix), (sz, ch)
        if full:
            want = chk.encode(host, sz, sz, ch)
            n = int(d_lens[0])
            assert n == len(want)
            assert d_streams[:n].cpu().numpy().tobytes() == want
        plan.close()
        del d_pix, d_streams, d_dec
        torch.cuda.empty_cache()
