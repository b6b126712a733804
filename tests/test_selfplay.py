"""Batched self-play driver (csrc/host/selfplay.h, sp_selfplay_run): SURVEY.md section 8 f3/f4.

CPU: the driver's host logic (resumable search == recursive search, scheduler protocol, viriformat
records) with a stand-in evaluator that exists in the test only.
GPU: the real thing through the C-ABI -- every recorded score must be reproduced by re-searching the
recorded position with the CPU oracle as the evaluator (tests only), which pins every one of the
millions of batched device evaluations behind those scores to the reference's arithmetic."""
import os
import subprocess
import sys

import numpy as np
import pytest

from stormphrax_b200 import api, build

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HOST = os.path.join(ROOT, "stormphrax_b200", "csrc", "host")


def test_selfplay_host_logic(tmp_path):
    exe = str(tmp_path / "test_selfplay_host")
    subprocess.run(
        ["g++", "-std=c++17", "-O2", "-Wall", "-o", exe, os.path.join(ROOT, "tests", "cpp", "test_selfplay_host.cpp"),
         os.path.join(HOST, "position.cpp")],
        check=True,
    )
    r = subprocess.run([exe], capture_output=True, text=True, timeout=600)
    sys.stderr.write(r.stderr[-2000:])
    assert r.returncode == 0, r.stdout + r.stderr[-2000:]
    assert "0 failures" in r.stdout


def test_pin_aware_movegen_equals_make_and_test(tmp_path):
    exe = str(tmp_path / "test_movegen")
    subprocess.run(
        ["g++", "-std=c++17", "-O2", "-Wall", "-o", exe, os.path.join(ROOT, "tests", "cpp", "test_movegen.cpp"), os.path.join(HOST, "position.cpp")],
        check=True,
    )
    r = subprocess.run([exe], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0 and " 0 failures" in r.stdout, r.stdout + r.stderr[-2000:]


def test_viriformat_move_encoding():
    """from | to << 6 | promo << 12 | type flags (viriformat.cpp:33-49) against hand-built moves."""
    b = api.board_from_fen("r3k2r/1P6/8/3pP3/8/8/8/R3K2R w KQkq d6 0 1")
    moves = api.legal_moves(b)
    by_kind = {}
    for m in moves:
        by_kind.setdefault(int(m) & 3, []).append(int(m))
    assert set(by_kind) == {0, 1, 2, 3}
    for m in moves:
        m = int(m)
        frm, to, kind = m >> 10, (m >> 4) & 63, m & 3
        viri = frm | to << 6 | ((m >> 2) & 3 if kind == 1 else 0) << 12 | [0x0000, 0xC000, 0x8000, 0x4000][kind]
        assert api.viri_to_move(b, viri) == m


def _research(oracle, board, depth, nodes_per_move):
    """The stand-in search of selfplay.h, recursively, with the CPU oracle as the static evaluator."""
    MATE, WIN = 32766, 25000
    state = {"nodes": 0, "best": 0, "prev": 0}

    def order_key(b, m, root):
        m = int(m)
        frm, to, kind = m >> 10, (m >> 4) & 63, m & 3
        mailbox = oracle_mailbox(b)
        victim = 0 if kind == 3 else (12 if kind == 2 else mailbox[to])
        k = 0
        if victim != 12:
            k = 16 + 2 * (victim >> 1) - (1 if (mailbox[frm] >> 1) > (victim >> 1) else 0)
        if kind == 1:
            k += 8
        if root and m == state["prev"]:
            k = 1000
        return k

    def search(b, depth, alpha, beta, ply):
        state["nodes"] += 1
        moves = [int(m) for m in api.legal_moves(b)]
        in_check = is_check(b)
        if not moves:
            return (-MATE + ply) if in_check else 0
        if ply > 0 and int(b["halfmove"]) >= 100:
            return 0
        if in_check and depth == 0 and ply + 1 < 24:
            depth = 1
        if not in_check or depth == 0:
            ev = int(oracle.eval_once(np.array([b], dtype=api.BOARD_DTYPE))[0])
            ev = max(-WIN + 1, min(WIN - 1, ev))
            if depth == 0:
                return ev
            if ply > 0 and depth <= 2 and ev - 120 * depth >= beta:
                return ev
        moves.sort(key=lambda m: -order_key(b, m, ply == 0))  # stable, like std::stable_sort
        best = -MATE
        for m in moves:
            if alpha >= beta:
                break
            v = -search(api.apply_move(b, m)[0], depth - 1, -beta, -alpha, ply + 1)
            if v > best:
                best = v
                if ply == 0:
                    state["best"] = m
            alpha = max(alpha, v)
        return best

    score = 0
    d = 1
    while True:
        state["best"] = 0
        score = search(board, d, -MATE, MATE, 0)
        state["prev"] = state["best"]
        if not (d < depth and state["nodes"] < nodes_per_move and abs(score) <= 25000):
            break
        d += 1
    return score, state["prev"]


def oracle_mailbox(b):
    """piece codes (type << 1 | colour, 12 = none) per square from a packed record"""
    box = [12] * 64
    occ = int(b["occupancy"])
    i = 0
    for sq in range(64):
        if occ >> sq & 1:
            nib = (int(b["pieces"][i // 2]) >> (4 * (i & 1))) & 15
            t = nib & 7
            t = 3 if t == 6 else t
            box[sq] = t << 1 | (0 if nib & 8 else 1)
            i += 1
    return box


def is_check(b):
    return api.in_check(b)


@pytest.mark.gpu
def test_selfplay_device_records_replay_and_scores_match_oracle(net, c_oracle):
    data, stats = api.selfplay(net.image, 0, concurrency=256, total_games=320, threads=2, depth=2, nodes_per_move=200, max_plies=60, seed=11)
    games = api.parse_viriformat(data)
    assert len(games) == stats["games"] == 320
    assert sum(len(g[1]) for g in games) == stats["positions"]
    assert stats["evals"] / stats["batches"] > 24, stats  # leaves really are coalesced
    rng = np.random.default_rng(5)
    checked = 0
    for start, moves, scores in games:
        assert int(start["wdl"]) in (0, 1, 2)
        b = start.copy()
        b["wdl"] = 0
        for i, (vm, sc) in enumerate(zip(moves, scores)):
            m = api.viri_to_move(b, int(vm))  # raises if the recorded move is not legal here
            if rng.random() < 0.02 and checked < 40:
                score, best = _research(c_oracle, b, 2, 200)
                white = score if not (int(b["stm_ep"]) & 0x80) else -score
                assert best == m, (api.board_to_fen(b), hex(best), hex(m))
                if not (i == len(moves) - 1 and sc == 0):
                    assert int(sc) == (0 if abs(white) <= 2 else white), (api.board_to_fen(b), int(sc), white)
                checked += 1
            b = api.apply_move(b, m)[0]
    assert checked >= 20


@pytest.mark.gpu
def test_selfplay_is_deterministic_whatever_the_thread_count(net):
    """Every (slot, game) has its own random stream and records are emitted slot-major."""
    kw = dict(concurrency=64, total_games=100, depth=2, nodes_per_move=100, max_plies=40, seed=3)
    a, sa = api.selfplay(net.image, 0, threads=1, **kw)
    b, sb = api.selfplay(net.image, 0, threads=3, **kw)
    assert np.array_equal(a, b)
    assert {k: v for k, v in sa.items() if k != "batches"} == {k: v for k, v in sb.items() if k != "batches"}


@pytest.mark.gpu
def test_gpu_resident_selfplay_plays_the_same_games_as_the_host_driver(net):
    """sp_selfplay_run_gpu: the search state machine runs per device thread; records must be byte-identical to the
    host driver's (same board code, same search, exact integer evaluations), and so must the node / eval counts."""
    kw = dict(concurrency=96, total_games=150, depth=3, nodes_per_move=400, max_plies=50, seed=21)
    host, sh = api.selfplay(net.image, 0, threads=2, **kw)
    dev, sd = api.selfplay(net.image, 0, resident=True, **kw)
    games = api.parse_viriformat(dev)
    assert len(games) == 150
    assert len({g[0].tobytes() + g[1].tobytes() for g in games}) == 150  # no game is played twice
    assert np.array_equal(host, dev)
    for k in ("games", "positions", "nodes", "evals", "searches"):
        assert sh[k] == sd[k], (k, sh[k], sd[k])
    assert sd["evals"] / sd["batches"] > 12  # one device batch per round of all running games (96 slots, long tail)
    # three concurrent driver instances (slot ranges) on the same device: still the same bytes
    dev3, sd3 = api.selfplay(net.image, 0, resident=True, threads=3, **kw)
    assert np.array_equal(host, dev3) and sd3["evals"] == sh["evals"]
    # `datagen dfrc`: random double-Fischer-random starts (Chess960 castling through the whole stack)
    kw = dict(concurrency=48, total_games=60, depth=2, nodes_per_move=200, max_plies=40, seed=5, dfrc=True)
    host, sh = api.selfplay(net.image, 0, threads=2, **kw)
    dev, sd = api.selfplay(net.image, 0, resident=True, **kw)
    assert np.array_equal(host, dev) and sh["evals"] == sd["evals"]
    starts = {api.board_to_fen(g[0]).split()[0] for g in api.parse_viriformat(dev)}
    assert len(starts) >= 59  # every game has its own random stream and start position


def _golden_datagen():
    path = os.path.join(ROOT, "tests", "golden", "datagen_seed42.npz")
    if not os.path.exists(path):
        pytest.skip("run tests/golden/make_datagen_golden.py where /root/reference exists")
    return np.load(path), np.load(os.path.join(ROOT, "tests", "golden", "playouts_seed42.npz"))


def test_viriformat_records_match_reference_golden():
    """Byte-for-byte against records written by the reference's own Viriformat class (standard + DFRC games)."""
    d, g = _golden_datagen()
    k = 0
    for prefix in ("", "dfrc_"):
        boards, moves, starts, evals = g[prefix + "boards"], g[prefix + "moves"], g[prefix + "starts"], g[prefix + "evals"]
        for i in range(len(starts) - 1):
            lo, hi = int(starts[i]), int(starts[i + 1])
            scores = np.clip(evals[lo : hi - 1], -32768, 32767).astype(np.int16)
            got = api.viriformat(boards[lo], moves[lo : hi - 1], scores, int(d["viri_outcome"][k]))
            want = d["viri"][d["viri_off"][k] : d["viri_off"][k + 1]]
            assert np.array_equal(got, want), f"{prefix}game {i}"
            # and the parser / move decoder round-trips what was written
            (start, vm, sc), = api.parse_viriformat(got)
            assert np.array_equal(sc, scores) and int(start["wdl"]) == int(d["viri_outcome"][k])
            b = boards[lo]
            for j, v in enumerate(vm[:12]):
                assert api.viri_to_move(b, int(v)) == int(moves[lo + j])
                b = api.apply_move(b, int(moves[lo + j]))[0]
            k += 1
    assert k == len(d["viri_outcome"])


def test_wdl_model_matches_reference_golden():
    d, _ = _golden_datagen()
    for b, s, want in zip(d["norm_boards"], d["norm_scores"], d["wdl_model"]):
        assert api.wdl_model(b, int(s)) == (int(want[0]), int(want[1])), (api.board_to_fen(b), int(s))


def test_dfrc_start_positions_match_reference_golden():
    d, _ = _golden_datagen()
    for i, want in zip(d["dfrc_index"], d["dfrc_boards"]):
        assert api.board_from_dfrc(int(i))[0].tobytes() == want.tobytes(), int(i)
    assert api.board_to_fen(api.board_from_dfrc(518 * 960 + 518)).startswith("rnbqkbnr/pppppppp/8/8/8/8/PPPPPPPP/RNBQKBNR w")


def test_normalize_score_matches_reference_golden():
    d, _ = _golden_datagen()
    for b, s, m, n in zip(d["norm_boards"], d["norm_scores"], d["norm_material"], d["norm_out"]):
        assert api.normalize_score(b, int(s)) == (int(m), int(n))
