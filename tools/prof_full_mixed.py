#!/usr/bin/env python
"""Ordered, shuffled and ordered positions again through ONE context: the full refresh falls back to the warp kernel while its
launches mostly overflow (unshared rows) and returns to the group kernel afterwards.  usage: python tools/prof_full_mixed.py"""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from stormphrax_b200 import api, net as N
from bench import make_workload

n = 1 << 20
boards, _, _ = make_workload(0, n)
shuffled = boards[np.random.default_rng(0).permutation(n)]
ctx = api.Nnue(N.synthetic(1234).image, 0)
stream = torch.cuda.Stream(); torch.cuda.set_stream(stream); s = stream.cuda_stream
d = {"ordered": torch.from_numpy(boards.view(np.uint8).reshape(-1)).cuda(), "shuffled": torch.from_numpy(shuffled.view(np.uint8).reshape(-1)).cuda()}
d_out = torch.empty(n, dtype=torch.int32, device="cuda")
want = {}
for name in ("ordered", "shuffled", "ordered", "shuffled", "ordered"):
    for rep in range(6):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        ctx.eval_full_device(d[name], n, d_out, s)
        e1.record(stream); ctx.sync(s); torch.cuda.synchronize()
        got = d_out.cpu().numpy()
        if name not in want:
            want[name] = got.copy()
        assert (got == want[name]).all(), (name, rep)
        print(f"{name:9s} pass {rep}: {n / e0.elapsed_time(e1) / 1e3:6.1f} Mpos/s")
assert (want["shuffled"] == want["ordered"][np.random.default_rng(0).permutation(n)]).all()
print("results identical across kernels and orders")
