"""Pin the plain-C oracle (oracle/nnue_oracle.c) to the reference.

Two anchors: (i) golden vectors produced by the reference's own compiled code
(tests/golden/make_golden.py), which travel everywhere; (ii) the live reference build
(oracle/_ref) where it is runnable.  The reference tree holds no test vectors of its own
(SURVEY.md section 4), so (i)+(ii) are what "pinned" means here.
"""
import numpy as np
import pytest

from stormphrax_b200 import net as N


def test_oracle_matches_golden_evals(c_oracle, golden):
    assert (c_oracle.eval_once(golden["boards"]) == golden["evals"]).all()
    assert (c_oracle.eval_once(golden["dfrc_boards"]) == golden["dfrc_evals"]).all()
    assert (c_oracle.eval_once(golden["fen_boards"]) == golden["fen_evals"]).all()


def test_oracle_matches_golden_features(c_oracle, golden):
    boards = golden["boards"]
    for n, i in enumerate(golden["feat_pick"]):
        for c in range(2):
            lo, hi = golden[f"psq{c}_off"][n], golden[f"psq{c}_off"][n + 1]
            assert (c_oracle.psq_features(boards[i], c) == golden[f"psq{c}"][lo:hi]).all()
            lo, hi = golden[f"thr{c}_off"][n], golden[f"thr{c}_off"][n + 1]
            assert (np.sort(c_oracle.threat_features(boards[i], c)) == golden[f"thr{c}"][lo:hi]).all()


def test_oracle_stress_net_matches_golden(golden):
    """Full-range weights: int16 accumulators and int32 dense sums wrap constantly."""
    from oracle.bind import COracle

    stress = np.load(__import__("os").path.join(__import__("os").path.dirname(__file__), "golden", "stress_seed99.npz"))
    o = COracle()
    o.load_net(N.synthetic(99, stress=True).image)
    try:
        assert (o.eval_once(golden["boards"]) == stress["evals"]).all()
        assert (o.eval_once(golden["fen_boards"]) == stress["fen_evals"]).all()
    finally:
        o.load_net(N.synthetic(1234).image)  # the C oracle keeps one global network


def test_oracle_matches_live_reference(c_oracle, reference):
    boards, moves, starts = reference.playouts(2024, 30, 80)
    ref = reference.eval_once(boards)
    assert (c_oracle.eval_once(boards) == ref).all()
    # the reference's own invariant (datagen.cpp:262): incremental == from scratch, both engine forms
    for g in range(6):
        lo, hi = starts[g], starts[g + 1]
        assert (reference.eval_playout(boards[lo], moves[lo : hi - 1], mode=0) == ref[lo:hi]).all()
        lazy = reference.eval_playout(boards[lo], moves[lo : hi - 1], mode=1, stride=5)
        seen = lazy != np.iinfo(np.int32).min
        assert seen.any() and (lazy[seen] == ref[lo:hi][seen]).all()


def test_oracle_threat_index_table_matches_reference(c_oracle, reference):
    rng = np.random.default_rng(5)
    for _ in range(20000):
        c, k = int(rng.integers(2)), int(rng.integers(64))
        a, v = int(rng.integers(12)), int(rng.integers(12))
        asq, vsq = int(rng.integers(64)), int(rng.integers(64))
        if a >> 1 == 0 and (asq < 8 or asq >= 56):
            continue  # pawns never stand on the back ranks; the tables are not defined there
        r = reference.threat_index(c, k, a, asq, v, vsq)
        o = c_oracle.threat_index(c, k, a, asq, v, vsq)
        assert (r < 0 and o < 0) or r == o


def test_avx2_and_avx512_reference_builds_agree(net):
    from oracle.bind import Reference, ref_isa_available
    import os

    isas = [i for i in ref_isa_available() if os.path.exists(Reference.path(i))]
    if len(isas) < 2:
        pytest.skip("needs both reference ISA builds runnable")
    a, b = Reference(isas[0]), Reference(isas[1])
    a.load_net(net.image)
    b.load_net(net.image)
    boards, _, _ = a.playouts(11, 10, 60)
    assert (a.eval_once(boards) == b.eval_once(boards)).all()


def test_synthetic_net_format(net):
    assert net.image.size == N.FILE_BYTES
    N.validate_header(net.image[:64].tobytes())
    bad = net.image[:64].copy()
    bad[0] = ord("X")
    with pytest.raises(N.NetworkFormatError):
        N.validate_header(bad.tobytes())


def _adjust_restatement(boards, raw, contempt, optimism):
    """numpy restatement of adjustStatic + adjustEval<false> (src/eval/eval.cpp:25-67) with the default
    tunables (src/tunable.h:161-169); C++ int division truncates toward zero."""
    tdiv = lambda a, b: np.trunc(a / b).astype(np.int64)
    value = np.array([48, 442, 461, 637, 1223, 0, 637, 0], dtype=np.int64)  # P N B R Q K castling-rook -
    out = np.empty(len(boards), dtype=np.int64)
    for i, b in enumerate(boards):
        n = bin(int(b["occupancy"])).count("1")
        nibbles = np.stack([b["pieces"] & 0xF, b["pieces"] >> 4], axis=1).reshape(-1)[:n] & 7
        material = int(value[nibbles].sum())
        stm = 0 if b["stm_ep"] & 0x80 else 1
        e = int(np.clip(int(raw[i]) + contempt[stm], -24999, 24999))
        e = int(tdiv(e * (26000 + material) + optimism[stm] * (2024 + int(tdiv(material * 1005, 1024))), 32768))
        e = int(tdiv(e * (200 - int(b["halfmove"])), 200))
        out[i] = np.clip(e, -24999, 24999)
    return out


def test_adjust_restatement_matches_reference_golden():
    import os

    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "adjust_seed42.npz"))
    for k in range(3):
        c0, c1, o0, o1 = (int(x) for x in g[f"adjust_params{k}"])
        got = _adjust_restatement(g["adjust_boards"], g["adjust_raw"], (c0, c1), (o0, o1))
        assert (got == g[f"adjust_out{k}"]).all(), k
