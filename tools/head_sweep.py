#!/usr/bin/env python
"""BASELINE config 4: the dense head (int8 IMMA L1 -> L2 -> L3) in isolation, batch sweep 2^10..2^20.
Activations come from real positions (FT output of the config-2 batch) and from uniform random bytes.
Prints a markdown table: Mpos/s, int8 TOP/s (2 * 65536 op/position), GB/s of (activations + bucket + eval).
usage: python tools/head_sweep.py > profiles/r1_head_sweep.md"""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from stormphrax_b200 import api, net as N
from bench import make_workload, measured_peak_hbm

def main():
    n_max = 1 << 20
    boards, _, _ = make_workload(0, n_max)
    ctx = api.Nnue(N.synthetic(1234).image, 0)
    stream = torch.cuda.Stream(); torch.cuda.set_stream(stream); s = stream.cuda_stream
    d_boards = torch.from_numpy(boards.view(np.uint8).reshape(-1)).cuda()
    d_act = torch.empty(n_max * 1024, dtype=torch.uint8, device="cuda")
    d_bucket = torch.empty(n_max, dtype=torch.uint8, device="cuda")
    d_out = torch.empty(n_max, dtype=torch.int32, device="cuda")
    ctx.activations_device(d_boards, n_max, d_act, d_bucket, s); ctx.sync(s)
    real = d_act.clone()
    rand = torch.randint(0, 128, (n_max * 1024,), dtype=torch.uint8, device="cuda")
    peak, _ = measured_peak_hbm()
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    print(f"# Round 1 -- dense head in isolation (BASELINE config 4), 1 x B200, SP_NNUE_HEAD={os.environ.get('SP_NNUE_HEAD', 'stream')}\n")
    print("`sp_nnue_forward_device`: u8[M][1024] activations + u8[M] buckets -> i32[M]; L2 flushed before every timed call;")
    print(f"HBM fraction = (1024 + 1 + 4) B/position against the measured {peak:.0f} GB/s; int8 TOP/s counts 2 x 1024 x 32 op/position.\n")
    print("| M | activations | us | Mpos/s | int8 TOP/s | GB/s | HBM frac |\n|---:|---|---:|---:|---:|---:|---:|")
    logms = [int(x) for x in os.environ["SWEEP_LOGM"].split(",")] if os.environ.get("SWEEP_LOGM") else range(10, 21)
    for logm in logms:
        m = 1 << logm
        for name, src in (("real FT output", real), ("uniform 0..127", rand)):
            for _ in range(3): ctx.forward_device(src, d_bucket, m, d_out, s)
            times = []
            for _ in range(7):
                flush.fill_(0)
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record(stream); ctx.forward_device(src, d_bucket, m, d_out, s); e1.record(stream)
                torch.cuda.synchronize(); times.append(e0.elapsed_time(e1))
            us = float(np.median(times)) * 1e3
            pos_s = m / (us * 1e-6)
            print(f"| 2^{logm} | {name} | {us:.1f} | {pos_s/1e6:.1f} | {pos_s*131072/1e12:.2f} | {pos_s*1029/1e9:.0f} | {pos_s*1029/1e9/peak:.3f} |")
    ctx.sync(s)

main()
