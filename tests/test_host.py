"""Host side of the product (no GPU): C-ABI surface, board model, shared feature/delta code.

The feature-index and delta code in csrc/sp_features.h / sp_delta.h is compiled for host AND
device; exercising it here on the CPU checks the very functions the kernels run per lane.
"""
import ctypes as C
import os
import re

import numpy as np
import pytest

from stormphrax_b200 import api

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    header = open(os.path.join(ROOT, "include", "sp_nnue.h")).read()
    declared = set(re.findall(r"\b(sp_(?:nnue|host|selfplay)_[a-z_0-9]+)\s*\(", header))
    assert declared, "no declarations found"
    L = api.lib()
    for name in sorted(declared):
        assert hasattr(L, name), f"{name} declared in include/sp_nnue.h but not exported"
    assert declared == set(api.SIGNATURES), declared ^ set(api.SIGNATURES)


def test_board_record_layout():
    assert api.BOARD_DTYPE.itemsize == 32
    assert api.BOARD_DTYPE.fields["stm_ep"][1] == 24


def test_create_rejects_bad_networks(net):
    L = api.lib()
    h = C.c_void_p()
    img = net.image[:4096].copy()
    assert L.sp_nnue_create(img.ctypes.data, 10, 0, C.byref(h)) == api.SP_ERR_BAD_NETWORK
    bad = img.copy()
    bad[0] = 0
    assert L.sp_nnue_create(bad.ctypes.data, bad.size, 0, C.byref(h)) == api.SP_ERR_BAD_NETWORK
    assert b"magic" in L.sp_nnue_last_error(None)
    bad = img.copy()
    bad[9] = 4  # arch id
    assert L.sp_nnue_create(bad.ctypes.data, bad.size, 0, C.byref(h)) == api.SP_ERR_BAD_NETWORK
    # valid header, truncated payload
    assert L.sp_nnue_create(img.ctypes.data, img.size, 0, C.byref(h)) == api.SP_ERR_BAD_NETWORK
    assert b"too small" in L.sp_nnue_last_error(None)


def test_no_cpu_fallback_without_device(net):
    import torch

    if torch.cuda.is_available():
        pytest.skip("a device is present")
    with pytest.raises(api.NnueError) as e:
        api.Nnue(net.image, 0)
    assert e.value.status == api.SP_ERR_NO_DEVICE


def test_fen_roundtrip_and_golden_boards(golden):
    for fen, board in zip(golden["fens"], golden["fen_boards"]):
        mine = api.board_from_fen(str(fen))
        assert mine.tobytes() == board.tobytes(), fen
        assert api.board_from_fen(api.board_to_fen(mine)).tobytes() == board.tobytes()


def test_legal_moves_match_reference_golden(golden):
    boards = golden["boards"]
    for n, i in enumerate(golden["move_pick"]):
        lo, hi = golden["move_off"][n], golden["move_off"][n + 1]
        assert (np.sort(api.legal_moves(boards[i])) == golden["move_flat"][lo:hi]).all(), api.board_to_fen(boards[i])


def test_apply_move_matches_reference_golden(golden):
    for key in ("", "dfrc_"):
        boards, moves, starts = golden[key + "boards"], golden[key + "moves"], golden[key + "starts"]
        for g in range(len(starts) - 1):
            for i in range(starts[g], starts[g + 1] - 1):
                nxt = api.apply_move(boards[i], moves[i])
                assert nxt.tobytes() == boards[i + 1].tobytes(), (key, g, i, api.board_to_fen(boards[i]), hex(moves[i]))


def test_dfrc_legal_moves_match_live_reference(reference, golden):
    boards = golden["dfrc_boards"]
    for i in range(0, len(boards), 5):
        assert (np.sort(api.legal_moves(boards[i])) == np.sort(reference.legal_moves(boards[i]))).all(), api.board_to_fen(boards[i])


def test_playouts_deterministic_and_thread_independent():
    a = api.playouts(5, 16, 40, threads=1)
    b = api.playouts(5, 16, 40, threads=4)
    for x, y in zip(a, b):
        assert x.tobytes() == y.tobytes()
    boards, moves, starts = a
    assert starts[0] == 0 and starts[-1] == len(boards)
    # every game starts from the initial position and every board follows from its move
    for g in range(16):
        assert api.board_to_fen(boards[starts[g]]).startswith("rnbqkbnr/pppppppp/8/8/8/8/PPPPPPPP/RNBQKBNR w")
        for i in range(starts[g], starts[g + 1] - 1):
            assert api.apply_move(boards[i], moves[i]).tobytes() == boards[i + 1].tobytes()


def test_shared_feature_code_matches_golden(golden):
    boards = golden["boards"]
    for n, i in enumerate(golden["feat_pick"]):
        for c in range(2):
            lo, hi = golden[f"psq{c}_off"][n], golden[f"psq{c}_off"][n + 1]
            assert (np.sort(api.features(boards[i], c, 0)) == np.sort(golden[f"psq{c}"][lo:hi])).all()
            lo, hi = golden[f"thr{c}_off"][n], golden[f"thr{c}_off"][n + 1]
            assert (np.sort(api.features(boards[i], c, 1)) == golden[f"thr{c}"][lo:hi]).all()


def test_shared_feature_code_matches_oracle_on_fens(c_oracle, golden):
    for board in golden["fen_boards"]:
        for c in range(2):
            assert (np.sort(api.features(board, c, 0)) == np.sort(c_oracle.psq_features(board, c))).all()
            assert (np.sort(api.features(board, c, 1)) == np.sort(c_oracle.threat_features(board, c))).all()


def _multiset_delta(before, after):
    """features(after) - features(before) as (adds, subs) multisets."""
    from collections import Counter

    a, b = Counter(after.tolist()), Counter(before.tolist())
    return a - b, b - a


@pytest.mark.parametrize("key", ["", "dfrc_"])
def test_delta_generator_equals_feature_set_difference(c_oracle, golden, key):
    """sp_delta.h: (adds - subs) must equal features(after) - features(before) as multisets, for
    every move kind (quiet, capture, castling incl. 960, en passant, promotion)."""
    from collections import Counter

    boards, starts = golden[key + "boards"], golden[key + "starts"]
    checked = refreshes = 0
    for g in range(len(starts) - 1):
        for i in range(starts[g], starts[g + 1] - 1):
            for c in range(2):
                refresh, pa, ps, ta, ts = api.feature_delta(boards[i], boards[i + 1], c)
                if refresh:
                    refreshes += 1
                    continue
                for kind, add, sub in ((0, pa, ps), (1, ta, ts)):
                    want_add, want_sub = _multiset_delta(api.features(boards[i], c, kind), api.features(boards[i + 1], c, kind))
                    got = Counter(add.tolist())
                    got.subtract(Counter(sub.tolist()))
                    want = Counter(want_add)
                    want.subtract(want_sub)
                    assert {k: v for k, v in got.items() if v} == {k: v for k, v in want.items() if v}, (
                        key, g, i, c, kind, api.board_to_fen(boards[i]), api.board_to_fen(boards[i + 1]))
                checked += 1
    assert checked > 1000 and refreshes > 0


def test_board_records_round_trip_through_the_board_model(golden):
    """marlinformat record -> Position -> FEN -> Position -> record is the identity on every golden board (standard
    and Chess960 castling rights, en passant squares, clocks); the eval / wdl / extra fields are not board state."""
    for boards in (golden["boards"][::3], golden["dfrc_boards"][::3], golden["fen_boards"]):
        for b in boards:
            fen = api.board_to_fen(b)
            again = api.board_from_fen(fen)[0]
            for field in ("occupancy", "pieces", "stm_ep", "halfmove", "fullmove"):
                assert np.array_equal(again[field], b[field]), (fen, field)
            assert api.board_to_fen(again) == fen


def test_host_adjust_eval_matches_reference_golden():
    """The C++ mirror's per-position adjustStatic + adjustEval (host/nnue_state.h) on the vectors the reference's own
    staticEvalOnce / adjustEval<false> produced, and with a correction term against the restated arithmetic."""
    g = np.load(os.path.join(ROOT, "tests", "golden", "adjust_seed42.npz"))
    boards, raw = g["adjust_boards"], g["adjust_raw"]
    for k in range(3):
        c0, c1, o0, o1 = (int(x) for x in g[f"adjust_params{k}"])
        params = api.AdjustParams.defaults(contempt=(c0, c1), optimism=(o0, o1))
        assert (api.host_adjust(boards, raw, params) == g[f"adjust_out{k}"]).all(), k
    corr = np.random.default_rng(4).integers(-300000, 300000, len(boards)).astype(np.int32)
    base = g["adjust_out0"].astype(np.int64)
    want = np.clip(base + np.trunc(corr / 2048).astype(np.int64), -24999, 24999)
    inside = np.abs(base) < 24999  # where the golden value was not clamped, the pre-clamp value is known
    got = api.host_adjust(boards, raw, api.AdjustParams.defaults(), correction=corr)
    assert (got[inside] == want[inside]).all()


def test_zstd_flagged_network_decodes_to_the_same_payload(net):
    """eval::init accepts a zstd-compressed network (src/eval/nnue.cpp:215-247): header uncompressed, the arrays one
    zstd frame.  The loader's host half must return the identical logical payload; damage must be reported."""
    from stormphrax_b200 import net as N

    pytest.importorskip("pyarrow")
    z = N.compressed(net)
    assert z.size < net.image.size // 1.5 and z[6] & 1
    assert np.array_equal(api.net_payload(z), net.image[64:])
    assert np.array_equal(api.net_payload(net.image), net.image[64:])
    with pytest.raises(api.NnueError):
        api.net_payload(z[: z.size // 2])  # truncated frame
    short = N.compressed(type(net)(image=net.image[: 64 + 4096], **{k: getattr(net, k) for k in ("psq_w", "thr_w", "ft_b", "l1_w", "l1_b", "l2_w", "l2_b", "l3_w", "l3_b")}))
    with pytest.raises(api.NnueError):
        api.net_payload(short)  # a valid frame that inflates to too few bytes (nnue.cpp:243-246)
