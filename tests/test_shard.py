"""The N>1 host logic of bench.py on CPU: contiguous sharding of the batch + the sizes exchange, with
world_size 2 over gloo.  The per-image codec in this test is the oracle (there is no GPU here); what is
checked is that sharding does not change per-image results and that every image is owned exactly once."""
import os
import socket
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def test_partition_properties():
    from bench import shard_images

    for world in (1, 2, 3, 4, 8):
        for gb in (1, 7, 8, 32, 256, 1000):
            owned = []
            for r in range(world):
                lo, hi = shard_images(gb, world, r)
                assert 0 <= lo <= hi <= gb
                owned += list(range(lo, hi))
            assert owned == list(range(gb))
    # weak scaling configuration of bench.py: every rank gets exactly the per-GPU batch
    for world in (1, 2, 4, 8):
        assert all(shard_images(32 * world, world, r) == (32 * r, 32 * r + 32) for r in range(world))


def _worker(rank, world, port, gb, w, h, ch, dist_id, out_path):
    import torch
    import torch.distributed as dist

    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import qoi_testlib as q
    from bench import shard_images

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    cpu = q.cpu_libs()
    lo, hi = shard_images(gb, world, rank)
    per = (gb + world - 1) // world
    lens = torch.zeros(per, dtype=torch.int32)
    sums = torch.zeros(per, dtype=torch.int64)
    for k, i in enumerate(range(lo, hi)):
        s = cpu.oracle.encode(cpu.image(w, h, ch, dist_id, i), w, h, ch)
        lens[k] = len(s)
        sums[k] = int(np.frombuffer(s, dtype=np.uint8).astype(np.int64).sum())
    all_lens = [torch.zeros_like(lens) for _ in range(world)]
    all_sums = [torch.zeros_like(sums) for _ in range(world)]
    dist.all_gather(all_lens, lens)       # the "sizes all-gather" of SURVEY.md 8(e)
    dist.all_gather(all_sums, sums)
    t = torch.tensor([float(rank + 1)])
    dist.all_reduce(t, op=dist.ReduceOp.MAX)  # the max-over-ranks timing reduction bench.py uses
    if rank == 0:
        np.save(out_path, np.stack([torch.cat(all_lens).numpy()[:gb].astype(np.int64), torch.cat(all_sums).numpy()[:gb]]))
        assert t.item() == world
    dist.destroy_process_group()


def test_two_rank_sharding_matches_single_process(tmp_path, cpu):
    import torch.multiprocessing as mp

    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    gb, w, h, ch, dist_id = 6, 96, 64, 4, 4
    out = str(tmp_path / "gathered.npy")
    mp.spawn(_worker, args=(2, port, gb, w, h, ch, dist_id, out), nprocs=2, join=True)
    got = np.load(out)
    for i in range(gb):
        s = cpu.oracle.encode(cpu.image(w, h, ch, dist_id, i), w, h, ch)
        assert got[0, i] == len(s)
        assert got[1, i] == int(np.frombuffer(s, dtype=np.uint8).astype(np.int64).sum())
