#!/usr/bin/env python
"""Run the dense head alone a few times (for `ncu -k regex:head`): u8[M][1024] activations from real
positions, L2 flushed between calls.  usage: python tools/head_once.py [log2_M=20] [reps=3]"""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from stormphrax_b200 import api, net as N
from bench import make_workload

def main():
    m = 1 << (int(sys.argv[1]) if len(sys.argv) > 1 else 20)
    reps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
    boards, _, _ = make_workload(0, m)
    ctx = api.Nnue(N.synthetic(1234).image, 0)
    stream = torch.cuda.Stream(); torch.cuda.set_stream(stream); s = stream.cuda_stream
    d_boards = torch.from_numpy(boards.view(np.uint8).reshape(-1)).cuda()
    d_act = torch.empty(m * 1024, dtype=torch.uint8, device="cuda")
    d_bucket = torch.empty(m, dtype=torch.uint8, device="cuda")
    d_out = torch.empty(m, dtype=torch.int32, device="cuda")
    ctx.activations_device(d_boards, m, d_act, d_bucket, s); ctx.sync(s)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    for _ in range(reps):
        flush.fill_(0)
        ctx.forward_device(d_act, d_bucket, m, d_out, s)
    ctx.sync(s)
    print("head_once: M =", m, "checksum", int(d_out.cpu().numpy().astype(np.int64).sum()))

main()
