#!/usr/bin/env python
"""Quick kernel-level timing for tuning (not the official bench): device-resident inputs only.
usage: [SP_NNUE_LIB=variant.so] python tools/kbench.py [full|playouts|both] [n_positions]"""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from stormphrax_b200 import api, net as N
from bench import make_workload

def main():
    which = sys.argv[1] if len(sys.argv) > 1 else "both"
    n = int(sys.argv[2]) if len(sys.argv) > 2 else 1 << 20
    boards, moves, starts = make_workload(0, n, "playouts" if (which == "playouts" and len(sys.argv) <= 2) else "full")
    n = len(boards)
    if os.environ.get("KBENCH_SHUFFLE"):  # full refresh only: destroy the game order (no row sharing between neighbours)
        boards = boards[np.random.default_rng(0).permutation(n)]
    ctx = api.Nnue(N.synthetic(1234).image, 0)
    stream = torch.cuda.Stream(); torch.cuda.set_stream(stream); s = stream.cuda_stream
    d_boards = torch.from_numpy(boards.view(np.uint8).reshape(-1)).cuda()
    d_starts = torch.from_numpy(starts.astype(np.uint32).view(np.int32)).cuda()
    d_out = torch.empty(n, dtype=torch.int32, device="cuda")
    ref = None
    for name in (["full", "playouts"] if which == "both" else [which]):
        step = (lambda: ctx.eval_full_device(d_boards, n, d_out, s)) if name == "full" else \
               (lambda: ctx.eval_playouts_device(d_boards, d_starts, len(starts) - 1, n, d_out, s))
        for _ in range(3): step()
        ctx.sync(s)
        out = d_out.cpu().numpy().copy()
        if ref is None: ref = out
        ok = bool((out == ref).all())
        ctx.profile(True); ctx.profile_read()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        reps = 5
        e0.record(stream)
        for _ in range(reps): step()
        e1.record(stream); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / reps
        prof = {k: round(v[0] / reps, 3) for k, v in ctx.profile_read().items() if v[1]}
        ctx.profile(False)
        print(f"{os.environ.get('SP_NNUE_LIB','default')[-12:]:>12s} {name:9s} {ms:8.3f} ms/step {n/ms/1e3:8.1f} Mpos/s kernels(ms/step)={prof} consistent={ok} checksum={int(out.astype(np.int64).sum())}", flush=True)

main()
