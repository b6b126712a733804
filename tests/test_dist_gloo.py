"""N > 1 plumbing on CPU: two gloo ranks shard a workload, all-reduce counters and timings."""
import os
import socket

import numpy as np
import pytest
import torch.multiprocessing as mp

from stormphrax_b200 import dist as D


def _free_port() -> int:
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank: int, world: int, port: int, out):
    os.environ.update(RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank), MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    from stormphrax_b200 import api

    assert D.init("gloo")
    boards, _moves, starts = api.playouts(99, 40, 60, threads=1)  # every rank can regenerate the workload
    lo, hi = D.shard_range(len(boards), rank, world)
    glo, ghi = D.shard_games(starts, rank, world)
    counters = np.zeros(8, dtype=np.uint64)
    counters[0] = hi - lo               # "evals" of this rank
    counters[3] = 2 + rank              # "launches"
    total = D.allreduce_counters(counters)
    slowest = D.max_over_ranks(10.0 + rank)
    D.barrier()
    out.put((rank, lo, hi, glo, ghi, int(starts[glo]), int(starts[ghi]), total.tolist(), slowest, len(boards), len(starts) - 1))


def test_two_rank_sharding_and_counter_allreduce():
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    results = sorted(q.get(timeout=120) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    n, n_games = results[0][9], results[0][10]
    # position shards tile [0, n) without gaps or overlap
    assert results[0][1] == 0 and results[-1][2] == n
    assert all(results[i][2] == results[i + 1][1] for i in range(world - 1))
    # game shards tile [0, n_games) on game boundaries and are roughly balanced
    assert results[0][3] == 0 and results[-1][4] == n_games
    assert all(results[i][4] == results[i + 1][3] for i in range(world - 1))
    sizes = [r[6] - r[5] for r in results]
    assert sum(sizes) == n and max(sizes) - min(sizes) <= 61 * 2
    for r in results:
        assert r[7][0] == n            # evals summed over ranks
        assert r[7][3] == 2 + 3        # launches summed
        assert r[8] == 11.0            # max over ranks


def test_single_rank_helpers_are_identity():
    assert D.shard_range(10, 0, 1) == (0, 10)
    c = np.arange(8, dtype=np.uint64)
    assert (D.allreduce_counters(c) == c).all()
    assert D.max_over_ranks(3.5) == 3.5
    assert [D.shard_range(10, r, 3) for r in range(3)] == [(0, 3), (3, 6), (6, 10)]
