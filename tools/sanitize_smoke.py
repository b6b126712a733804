#!/usr/bin/env python
"""A short pass over every kernel of the library for compute-sanitizer (memcheck / racecheck / synccheck):

    compute-sanitizer --tool racecheck python tools/sanitize_smoke.py

Full refresh, playout walker (with planned rebuilds), slot refresh / update / evaluate-only, the dense head at
a search-sized and a chunk-sized batch, adjust + wdl epilogues; every result is compared with the CPU oracle, so a
race that changes a value fails here as well."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from oracle.bind import COracle  # noqa: E402
from stormphrax_b200 import api, net as N  # noqa: E402


def main():
    net = N.synthetic(1234)
    games = int(os.environ.get("SANITIZE_GAMES", "12"))
    boards, _moves, starts = api.playouts(3, games, 60, threads=2)
    oracle = COracle()
    oracle.load_net(net.image)
    want = oracle.eval_once(boards)
    for kernel in ("stream", "umma"):  # the sorted tensor-core heads (small launches would take the warp-per-position kernel)
        os.environ["SP_NNUE_HEAD"], os.environ["SP_NNUE_HEAD_DIRECT"] = kernel, "0"
        with api.Nnue(net.image, 0) as ctx:
            assert np.array_equal(ctx.eval_full(boards), want), "full refresh, head " + kernel
    os.environ["SP_NNUE_HEAD_DIRECT"] = "2048"
    with api.Nnue(net.image, 0) as ctx:
        assert np.array_equal(ctx.eval_full(boards), want), "full refresh"
        assert np.array_equal(ctx.eval_playouts(boards, starts), want), "playouts"
        adj = ctx.adjust(boards[:64], want[:64])
        assert adj.shape == (64,)
        norm, win, loss = ctx.wdl(boards[:64], want[:64])
        assert norm.shape == win.shape == loss.shape == (64,)
    n = len(starts) - 1
    first = starts[:-1].astype(np.int64)
    ids = np.arange(n, dtype=np.uint32)
    for small in ("1", "0"):  # the one-launch kernel for search-sized rounds, then the general slot kernels
        os.environ["SP_NNUE_SMALL"] = small
        with api.Nnue(net.image, 0) as ctx:
            ctx.slots_reserve(3 * n)
            ctx.refresh(2 * n + ids, boards[first])
            ctx.refresh(2 * ids, boards[first])
            assert np.array_equal(ctx.eval_slots(2 * ids), want[first]), "slot refresh + evaluate"
            got = ctx.update_eval(2 * ids, 2 * ids + 1, boards[first + 1])
            assert np.array_equal(got, want[first + 1]), "slot update + evaluate"
            r, u, e = ctx.batch(refresh=(2 * ids, boards[first + 2]), update=(2 * ids + 1, 2 * ids + 1, boards[first + 2]), evaluate=(2 * n + ids, None))
            assert np.array_equal(r, want[first + 2]) and np.array_equal(u, want[first + 2]) and np.array_equal(e, want[first]), "batch"
    print(f"sanitize_smoke ok: {len(boards)} positions, {n} games")


if __name__ == "__main__":
    main()
