"""Generate tests/golden/datagen_seed42.npz from the REFERENCE's own compiled code (oracle/_ref):
viriformat records (src/datagen/viriformat.cpp) and wdl::normalizeScore<false> values (src/wdl.cpp)
for the games of playouts_seed42.npz, and double-Fischer-random start positions (Position::fromDfrcIndex).  Run where /root/reference exists, after `make -C oracle ref`:

    python tests/golden/make_datagen_golden.py
"""
from __future__ import annotations

import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from oracle.bind import Reference  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))


def main() -> None:
    ref = Reference()
    g = np.load(os.path.join(HERE, "playouts_seed42.npz"))
    records, offsets, outcomes = [], [0], []
    for prefix in ("", "dfrc_"):
        boards, moves, starts, evals = g[prefix + "boards"], g[prefix + "moves"], g[prefix + "starts"], g[prefix + "evals"]
        for k in range(len(starts) - 1):
            lo, hi = int(starts[k]), int(starts[k + 1])
            scores = np.clip(evals[lo : hi - 1], -32768, 32767).astype(np.int16)  # any int16 stream will do for the format
            outcome = k % 3
            rec = ref.viriformat(boards[lo], moves[lo : hi - 1], scores, outcome)
            records.append(rec)
            offsets.append(offsets[-1] + len(rec))
            outcomes.append(outcome)
    # score normalisation: every 5th board x a spread of scores incl. 0, decisive ones and both signs
    nb = g["boards"][::5]
    rng = np.random.default_rng(42)
    scores = np.concatenate([rng.integers(-3000, 3001, len(nb) - 8), [0, 1, -1, 30000, 30001, -30001, 24999, -24999]]).astype(np.int32)
    mat, norm = zip(*(ref.normalize_score(b, int(s)) for b, s in zip(nb, scores)))
    # win / loss per mille (wdl::wdlModel) on the same boards and scores
    wdl = np.array([ref.wdl_model(b, int(s)) for b, s in zip(nb, scores)], dtype=np.int32)
    # double-Fischer-random start positions (Position::fromDfrcIndex)
    dfrc_index = np.unique(np.concatenate([np.arange(0, 960 * 960, 1543), [0, 518 * 960 + 518, 960 * 960 - 1, 959, 960, 518]])).astype(np.uint32)
    dfrc_boards = np.concatenate([ref.board_from_dfrc(int(i)) for i in dfrc_index])
    np.savez_compressed(
        os.path.join(HERE, "datagen_seed42.npz"),
        dfrc_index=dfrc_index, dfrc_boards=dfrc_boards, wdl_model=wdl,
        viri=np.concatenate(records), viri_off=np.array(offsets, dtype=np.uint32), viri_outcome=np.array(outcomes, dtype=np.uint8),
        norm_boards=nb, norm_scores=scores, norm_material=np.array(mat, dtype=np.int32), norm_out=np.array(norm, dtype=np.int32),
    )
    print("wrote", len(records), "viriformat records,", len(scores), "normalised scores")


if __name__ == "__main__":
    main()
