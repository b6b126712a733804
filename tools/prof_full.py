#!/usr/bin/env python
"""A few full-refresh passes over N positions of the bench workload (device-resident), for ncu / timing variants.
usage: [SP_NNUE_LIB=...] python tools/prof_full.py [n_positions] [passes] [shuffle]"""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from stormphrax_b200 import api, net as N
from bench import make_workload

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 18
passes = int(sys.argv[2]) if len(sys.argv) > 2 else 3
boards, _, _ = make_workload(0, n)
if len(sys.argv) > 3:
    boards = boards[np.random.default_rng(0).permutation(n)]
ctx = api.Nnue(N.synthetic(1234).image, 0)
stream = torch.cuda.Stream(); torch.cuda.set_stream(stream); s = stream.cuda_stream
d_boards = torch.from_numpy(boards.view(np.uint8).reshape(-1)).cuda()
d_out = torch.empty(n, dtype=torch.int32, device="cuda")
ctx.eval_full_device(d_boards, n, d_out, s); ctx.sync(s)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(stream)
for _ in range(passes):
    ctx.eval_full_device(d_boards, n, d_out, s)
e1.record(stream); ctx.sync(s); torch.cuda.synchronize()
print(f"{n} positions x {passes}: {n * passes / e0.elapsed_time(e1) / 1e3:.1f} Mpos/s")
