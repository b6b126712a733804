"""Generate tests/golden/*.npz from the REFERENCE's own compiled code (oracle/_ref).

Run in the build container, where /root/reference exists and `make -C oracle ref` has been run:

    python tests/golden/make_golden.py

Inputs: the reference's move generator + RNG (spref_playouts) produce random legal playouts; the
synthetic network is stormphrax_b200.net.synthetic(1234) (numpy default_rng, deterministic).
Outputs: evaluateOnce results (nnue_state.cpp:612-634), incremental datagen-form results
(datagen.cpp:257-262), feature index lists, legal move lists, and special-move boards -- all as
produced by the reference.  The vectors travel to the GPU box; the reference does not.
"""
from __future__ import annotations

import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from oracle.bind import Reference  # noqa: E402
from stormphrax_b200 import net as N  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))

# positions that exercise castling (incl. Chess960), en passant, promotions, bare kings,
# every king bucket / mirror, and heavy threat lists
FENS = [
    "rnbqkbnr/pppppppp/8/8/8/8/PPPPPPPP/RNBQKBNR w KQkq - 0 1",
    "r3k2r/p1ppqpb1/bn2pnp1/3PN3/1p2P3/2N2Q1p/PPPBBPPP/R3K2R w KQkq - 0 1",
    "8/2p5/3p4/KP5r/1R3p1k/8/4P1P1/8 w - - 0 1",
    "r3k2r/Pppp1ppp/1b3nbN/nP6/BBP1P3/q4N2/Pp1P2PP/R2Q1RK1 w kq - 0 1",
    "rnbq1k1r/pp1Pbppp/2p5/8/2B5/8/PPP1NnPP/RNBQK2R w KQ - 1 8",
    "r4rk1/1pp1qppp/p1np1n2/2b1p1B1/2B1P1b1/P1NP1N2/1PP1QPPP/R4RK1 w - - 0 10",
    "4k3/8/8/8/8/8/8/4K3 w - - 0 1",
    "8/8/8/3k4/8/3K4/8/8 b - - 0 1",
    "8/P6k/8/8/8/8/p6K/8 w - - 0 1",
    "rnbqkbnr/ppp1p1pp/8/3pPp2/8/8/PPPP1PPP/RNBQKBNR w KQkq f6 0 3",
    "bnrqkrnb/pppppppp/8/8/8/8/PPPPPPPP/BNRQKRNB w FCfc - 0 1",
    "qqqqkqqq/qqqqqqqq/8/8/8/8/QQQQQQQQ/QQQQKQQQ w - - 0 1",
    "k7/8/8/8/8/8/8/7K w - - 0 1",
    "7k/8/8/8/8/8/8/K7 b - - 0 1",
    "r1b1k2r/ppppnppp/2n2q2/2b5/3NP3/2P1B3/PP3PPP/RN1QKB1R w KQkq - 0 1",
    "3Q4/1Q4Q1/4Q3/2Q4R/Q4Q2/3Q4/1Q4Rp/1K1BBNNk w - - 0 1",
]


def main() -> None:
    ref = Reference()
    net = N.synthetic(1234)
    ref.load_net(net.image)
    boards, moves, starts = ref.playouts(42, 40, 80)
    evals = ref.eval_once(boards)
    inc = np.concatenate(
        [ref.eval_playout(boards[starts[g]], moves[starts[g] : starts[g + 1] - 1]) for g in range(len(starts) - 1)]
    )
    assert (inc == evals).all(), "reference: incremental != evaluateOnce"
    dboards, dmoves, dstarts = ref.playouts(1337, 12, 60, dfrc=True)
    devals = ref.eval_once(dboards)

    fen_boards = np.concatenate([ref.board_from_fen(f) for f in FENS])
    fen_evals = ref.eval_once(fen_boards)

    # feature lists for a subset (variable length -> flat + offsets), perspective-major
    pick = np.arange(0, len(boards), 7)
    feats = {"psq": [[], []], "thr": [[], []]}
    for i in pick:
        for c in range(2):
            feats["psq"][c].append(ref.psq_features(boards[i], c))
            feats["thr"][c].append(np.sort(ref.threat_features(boards[i], c)))
    flat = {}
    for kind in ("psq", "thr"):
        for c in range(2):
            lists = feats[kind][c]
            flat[f"{kind}{c}_off"] = np.cumsum([0] + [len(x) for x in lists]).astype(np.uint32)
            flat[f"{kind}{c}"] = np.concatenate(lists).astype(np.uint32)

    # legal moves of a subset, sorted
    mpick = np.arange(0, len(boards), 11)
    mlists = [np.sort(ref.legal_moves(boards[i])) for i in mpick]
    np.savez_compressed(
        os.path.join(HERE, "playouts_seed42.npz"),
        boards=boards, moves=moves, starts=starts, evals=evals,
        dfrc_boards=dboards, dfrc_moves=dmoves, dfrc_starts=dstarts, dfrc_evals=devals,
        fen_boards=fen_boards, fen_evals=fen_evals, fens=np.array(FENS),
        feat_pick=pick.astype(np.uint32), **flat,
        move_pick=mpick.astype(np.uint32),
        move_off=np.cumsum([0] + [len(x) for x in mlists]).astype(np.uint32),
        move_flat=np.concatenate(mlists).astype(np.uint16),
        net_seed=np.int64(1234),
    )

    # eval post-processing: staticEvalOnce + adjustEval<false> from the reference, three parameter sets
    # (boards carry their real halfmove clocks; a few get large ones to exercise the 50-move damping)
    aboards = np.concatenate([boards[::3], fen_boards]).copy()
    aboards["halfmove"][::5] = (np.arange(len(aboards[::5])) * 7 % 101).astype(np.uint8)
    adjust = {"adjust_boards": aboards}
    for k, (contempt, optimism) in enumerate([((0, 0), (0, 0)), ((25, -25), (0, 0)), ((-40, 40), (120, -95))]):
        raw, adj = ref.adjusted_eval(aboards, contempt, optimism)
        adjust["adjust_raw"] = raw
        adjust[f"adjust_params{k}"] = np.array([*contempt, *optimism], dtype=np.int32)
        adjust[f"adjust_out{k}"] = adj
    np.savez_compressed(os.path.join(HERE, "adjust_seed42.npz"), **adjust)

    # the stress network (wrapping everywhere) on the same boards
    stress = N.synthetic(99, stress=True)
    ref.load_net(stress.image)
    np.savez_compressed(
        os.path.join(HERE, "stress_seed99.npz"),
        evals=ref.eval_once(boards), fen_evals=ref.eval_once(fen_boards), net_seed=np.int64(99),
    )
    print("wrote", len(boards), "positions,", len(dboards), "dfrc positions,", len(FENS), "fens")


if __name__ == "__main__":
    main()
