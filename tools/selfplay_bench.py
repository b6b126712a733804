#!/usr/bin/env python
"""BASELINE config 5 analogue: batched self-play (sp_selfplay_run) on one GPU.  Prints one JSON line:
static evaluations per second through the lazy NnueState / EvalBatch protocol, nodes/s, positions/s.
usage: python tools/selfplay_bench.py [concurrency=16384] [threads=16] [depth=3] [nodes_per_move=5000] [max_plies=40] [resident=0] [games_per_slot=1]
resident=1: sp_selfplay_run_gpu (the searches run on the device; threads = concurrent driver instances).
games_per_slot > 1 keeps the slots busy (steady state); with 1 the run is mostly tail, the slots emptying one by one."""
import json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
from stormphrax_b200 import api, dist as D, net as N


def main():
    a = [int(x) for x in sys.argv[1:]]
    conc, threads, depth, nodes, plies, resident, gps = (a + [16384, 16, 3, 5000, 40, 0, 1][len(a):])[:7]
    # under torchrun every rank plays its own games (own seed) on its own GPU; NCCL only sums the counters
    rank, world, local_rank = D.env_rank()
    if world > 1:
        import torch
        torch.cuda.set_device(local_rank)
        D.init("nccl", local_rank)
    image = N.synthetic(1234).image
    api.selfplay(image, local_rank, concurrency=64, total_games=64, threads=1, depth=1, nodes_per_move=1, max_plies=4, resident=bool(resident))  # warm-up: context, library
    D.barrier()
    t0 = time.perf_counter()
    data, st = api.selfplay(image, local_rank, seed=42 + rank, concurrency=conc, total_games=conc * gps, threads=threads, depth=depth, nodes_per_move=nodes, max_plies=plies, resident=bool(resident))
    dt = D.max_over_ranks(time.perf_counter() - t0)
    keys = list(st)
    summed = D.allreduce_counters(np.array([st[k] for k in keys] + [len(data)], dtype=np.uint64))
    st, n_bytes = {k: int(v) for k, v in zip(keys, summed[:-1])}, int(summed[-1])
    if rank:
        return
    print(json.dumps({
        "workload": "batched self-play, stand-in alpha-beta search, every static eval on the GPU",
        "driver": "GPU-resident (sp_selfplay_run_gpu)" if resident else "host threads (sp_selfplay_run)",
        "concurrency": conc, "host_threads": threads, "host_cores": os.cpu_count(), "depth": depth, "nodes_per_move": nodes, "max_plies": plies, "games_per_slot": gps,
        "seconds": round(dt, 3), **st,
        "evals_per_s": round(st["evals"] / dt), "nodes_per_s": round(st["nodes"] / dt), "positions_per_s": round(st["positions"] / dt),
        "evals_per_batch": round(st["evals"] / max(1, st["batches"]), 1), "viriformat_bytes": n_bytes, "n_gpus": world,
    }))


main()
