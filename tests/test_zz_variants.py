"""GPU parity of the encoder / decoder variants that are OFF by default (cuda/qoi_enc4.cuh, QOI_GPU_DEC=half).

Opt-in: these kernels have only run on the CPU emulator (the striped ones once on a B200), so they are kept out of the
round-end `pytest -m gpu` run -- a kernel that has never met real hardware must not be able to stall that tier.
Round 2 runs them explicitly, each case in its own process under a timeout:

    QOI_TEST_VARIANTS=1 python -m pytest tests/test_zz_variants.py -m gpu -q

Every case: batches of mixed content, RGB and RGBA, misaligned offsets, partial tiles -> streams byte-identical to the
CPU checker, indexed decode (optionally two lanes per tile) returns the input.
"""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ENC = ["striped", "striped-match", "striped-earlyw", "slab", "slab-match", "slab-pipe", "slab-pipe-ballot", "slab4", "slab4-pipe"]
CASES = [(e, "") for e in ENC] + [(e, "half") for e in ENC if e.startswith("slab")]

WORKER = r'''
import os, sys
import numpy as np
sys.path.insert(0, os.path.join(sys.argv[1], "tests")); sys.path.insert(0, sys.argv[1])
import torch
import qoi_testlib as q
import qoi_b200
cpu = q.cpu_libs(); chk = cpu.best()
ctx = qoi_b200.Context(0)                     # reads QOI_GPU_ENC / QOI_GPU_DEC
rng = np.random.default_rng(11)
shapes = [(4, 640, 480, 4, 4, 0), (3, 1000, 333, 3, 3, 0), (2, 2048, 2, 4, 6, 0), (5, 97, 61, 4, 0, 5), (2, 4096, 5, 3, 7, 1),
          (1, 2049, 1, 4, 2, 0), (3, 512, 512, 4, 5, 0), (1, 1920, 1080, 3, 4, 0)]
for (n, w, h, ch, dist, mis) in shapes:
    descs = [qoi_b200.Desc(w, h, ch, 0)] * n
    img = w * h * ch
    poff = np.array([mis + i * (img + 7 * (mis > 0)) for i in range(n)], dtype=np.uint64)
    plan = qoi_b200.Plan(ctx, descs, pixel_off=poff, decoded_channels=ch if ch == 3 else 4)
    host = np.zeros(int(poff[-1]) + img + 64, dtype=np.uint8)
    imgs = [cpu.image(w, h, ch, dist, i) if dist < 8 else q.stress_image(rng, w * h, ch, i % 6) for i in range(n)]
    for i in range(n):
        host[int(poff[i]):int(poff[i]) + img] = imgs[i]
    d_pix = torch.from_numpy(host).cuda()
    d_str = torch.full((plan.stream_bytes + 64,), 0xEE, dtype=torch.uint8, device="cuda")
    d_len = torch.zeros(n, dtype=torch.int32, device="cuda")
    d_dec = torch.zeros(plan.decoded_bytes + 64, dtype=torch.uint8, device="cuda")
    for rep in range(2):                      # twice: the second launch reuses the look-back records under a new epoch
        plan.encode(d_pix.data_ptr(), d_str.data_ptr(), d_len.data_ptr())
        plan.decode(d_str.data_ptr(), d_len.data_ptr(), d_dec.data_ptr(), use_index=True)
        torch.cuda.synchronize()
        lens = d_len.cpu().numpy(); streams = d_str.cpu().numpy(); dec = d_dec.cpu().numpy()
        och = ch if ch == 3 else 4
        for i in range(n):
            want = chk.encode(imgs[i], w, h, ch)
            o = int(plan.stream_off[i])
            assert lens[i] == len(want) and streams[o:o + lens[i]].tobytes() == want, ("stream", n, w, h, ch, dist, i, rep)
            assert streams[o + lens[i]] == 0xEE, ("wrote past the stream", n, w, h, ch, dist, i)
            ref_px, _ = chk.decode(want, och)
            do = int(plan._dec_off_eff[i])
            assert (dec[do:do + w * h * och] == ref_px).all(), ("decode", n, w, h, ch, dist, i, rep)
    plan.close()
ctx.close()
print("VARIANT OK")
'''


@pytest.mark.skipif(os.environ.get("QOI_TEST_VARIANTS") != "1", reason="opt-in: QOI_TEST_VARIANTS=1 (kernels that are off by default)")
@pytest.mark.parametrize("enc,dec", CASES)
def test_variant_parity(enc, dec):
    env = dict(os.environ, QOI_GPU_ENC=enc, QOI_GPU_DEC=dec)
    r = subprocess.run([sys.executable, "-c", WORKER, ROOT], capture_output=True, text=True, env=env, timeout=240)
    assert r.returncode == 0 and "VARIANT OK" in r.stdout, (enc, dec, r.stdout[-1500:], r.stderr[-3000:])
