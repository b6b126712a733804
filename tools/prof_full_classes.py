#!/usr/bin/env python
"""Full refresh with the library's per-kernel-class event timing on (sp_nnue_profile): where does a pass spend its time?
usage: python tools/prof_full_classes.py [n_positions] [shuffle]"""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from stormphrax_b200 import api, net as N
from bench import make_workload

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 20
boards, _, _ = make_workload(0, n)
if len(sys.argv) > 2:
    boards = boards[np.random.default_rng(0).permutation(n)]
ctx = api.Nnue(N.synthetic(1234).image, 0)
stream = torch.cuda.Stream(); torch.cuda.set_stream(stream); s = stream.cuda_stream
d_boards = torch.from_numpy(boards.view(np.uint8).reshape(-1)).cuda()
d_out = torch.empty(n, dtype=torch.int32, device="cuda")
ctx.eval_full_device(d_boards, n, d_out, s); ctx.sync(s)
ctx.profile(True); ctx.profile_read()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(stream)
for _ in range(3):
    ctx.eval_full_device(d_boards, n, d_out, s)
e1.record(stream); ctx.sync(s); torch.cuda.synchronize()
print(f"{n} positions x 3: {n * 3 / e0.elapsed_time(e1) / 1e3:.1f} Mpos/s, {e0.elapsed_time(e1) / 3:.3f} ms per pass")
print({k: (round(v[0] / 3, 3), v[1] // 3) for k, v in ctx.profile_read().items() if v[1]})
