"""Foreign-stream decode only (no restart index), for an ncu launch list: per-kernel time of the parse / chain /
speculative decode / verify / fix-up pipeline of cuda/qoi_fdec.cuh.  usage: foreign_prof.py [dist] [batch] [reps]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import qoi_b200
from qoi_b200.api import pinned_array
from bench import load_synth, W, H, CH

dist_id = int(sys.argv[1]) if len(sys.argv) > 1 else 4
n = int(sys.argv[2]) if len(sys.argv) > 2 else 8
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 2
ctx = qoi_b200.Context(0)
plan = qoi_b200.Plan(ctx, [qoi_b200.Desc(W, H, CH, 0)] * n, decoded_channels=4)
h_pix = pinned_array(plan.pixel_bytes)
load_synth().qoi_synth_batch(h_pix.ctypes.data, W * H * CH, n, W, H, CH, dist_id, 0)
d_pix = torch.from_numpy(h_pix).cuda()
d_streams = torch.empty(plan.stream_bytes, dtype=torch.uint8, device="cuda")
d_lens = torch.zeros(n, dtype=torch.int32, device="cuda")
d_dec = torch.empty(plan.decoded_bytes, dtype=torch.uint8, device="cuda")
s = torch.cuda.current_stream().cuda_stream
plan.encode(d_pix.data_ptr(), d_streams.data_ptr(), d_lens.data_ptr(), s)
ev = [torch.cuda.Event(enable_timing=True) for _ in range(reps + 1)]
ev[0].record()
for k in range(reps):
    plan.decode(d_streams.data_ptr(), d_lens.data_ptr(), d_dec.data_ptr(), False, None, s)
    ev[k + 1].record()
torch.cuda.synchronize()
assert torch.equal(d_dec, d_pix), "foreign decode mismatch"
ms = [ev[k].elapsed_time(ev[k + 1]) for k in range(reps)]
print(f"dist {dist_id} batch {n}: foreign decode ms {ms}  -> {n * W * H / (min(ms) * 1e-3) / 1e9:.3f} Gpx/s")
