"""Multi-device layer (include/qoi_gpu_multi.h) on real GPUs: the same batch on 1 and on N devices must give the same bytes
per image (sharding must not change results), with the NCCL scatter of the pixels from the root,
the all-gather of the lengths and the gather of the streams / pixels back to the root.  Skips below 2 GPUs."""
import numpy as np
import pytest

import qoi_testlib as q

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def multi():
    torch = pytest.importorskip("torch")
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    import qoi_b200
    from qoi_b200 import multi as qm

    return qoi_b200, qm, torch


def run(multi, cpu, n_dev, whc, dist):
    qb, qm, torch = multi
    chk = cpu.best()
    descs = [qb.Desc(w, h, c, 0) for (w, h, c) in whc]
    imgs = [cpu.image(w, h, c, dist, i) for i, (w, h, c) in enumerate(whc)]
    host = np.concatenate(imgs)
    mctx = qm.MultiContext(list(range(n_dev)))
    plan = qm.MultiPlan(mctx, descs, decoded_channels=4)
    with torch.cuda.device(0):
        d_root = torch.from_numpy(host).cuda()
        d_streams = torch.full((sum(w * h * 5 + 22 for (w, h, c) in whc),), 0xEE, dtype=torch.uint8, device="cuda")
        d_dec = torch.zeros(sum(w * h * 4 for (w, h, c) in whc), dtype=torch.uint8, device="cuda")
    lens, offs, t = plan.encode(d_root.data_ptr(), d_streams.data_ptr())
    streams = d_streams.cpu().numpy()
    out = []
    for i, (w, h, c) in enumerate(whc):
        want = chk.encode(imgs[i], w, h, c)
        got = streams[int(offs[i]):int(offs[i]) + int(lens[i])].tobytes()
        assert got == want, (n_dev, i)
        out.append(got)
    assert int(offs[-1]) + int(lens[-1]) == sum(len(s) for s in out)
    for use_index in (True, False):
        d_dec.zero_()
        plan.decode(use_index, d_dec.data_ptr())
        dec = d_dec.cpu().numpy()
        o = 0
        for i, (w, h, c) in enumerate(whc):
            want_px, _ = chk.decode(out[i], 4)
            assert (dec[o:o + w * h * 4] == want_px).all(), (n_dev, i, use_index)
            o += w * h * 4
    # shards resident, no scatter / gather: only the length exchange
    lens2, _, t2 = plan.encode(None, None)
    assert (lens2 == lens).all() and t2.scatter_ms == 0 and t2.gather_ms == 0
    plan.close()
    mctx.close()
    return out


def test_one_device_is_the_plain_batch(multi, cpu):
    run(multi, cpu, 1, [(300, 200, 4), (64, 64, 4), (1000, 33, 4)], 4)


def test_bytes_do_not_depend_on_the_number_of_devices(multi, cpu):
    qb, qm, torch = multi
    ndev = torch.cuda.device_count()
    if ndev < 2:
        pytest.skip("needs at least 2 GPUs")
    whc = [(640, 480, 4), (333, 77, 4), (2048, 9, 4), (96, 96, 4), (1, 1, 4), (800, 600, 4), (50, 50, 4)]
    one = run(multi, cpu, 1, whc, 3)
    for n in sorted({2, min(4, ndev), ndev}):
        assert run(multi, cpu, n, whc, 3) == one, n
    # more devices than images: idle ranks still take part in the collectives
    assert run(multi, cpu, ndev, whc[:1], 6) == run(multi, cpu, 1, whc[:1], 6)
    # RGB input
    rgb = [(w, h, 3) for (w, h, c) in whc]
    assert run(multi, cpu, 2, rgb, 4) == run(multi, cpu, 1, rgb, 4)
