"""GPU parity, incremental path (BASELINE.json config 3): accumulator slots and playout streams.

The reference's own invariant is `evaluate after any push/pop/apply sequence == evaluateOnce`
(src/datagen/datagen.cpp:262); golden `evals` are evaluateOnce results which the reference's
incremental code was checked against when the fixtures were generated (make_golden.py).
"""
import numpy as np
import pytest

from stormphrax_b200 import api

pytestmark = pytest.mark.gpu

INT32_MIN = np.iinfo(np.int32).min


def test_playout_walker_matches_golden(gpu_ctx, golden):
    assert (gpu_ctx.eval_playouts(golden["boards"], golden["starts"]) == golden["evals"]).all()


def test_playout_walker_dfrc_castling(gpu_ctx, golden):
    """Chess960 castling moves king and rook at once, sometimes onto each other's squares."""
    assert (gpu_ctx.eval_playouts(golden["dfrc_boards"], golden["dfrc_starts"]) == golden["dfrc_evals"]).all()


def test_playout_walker_stress_network(golden):
    import os

    from stormphrax_b200 import net as N

    stress = np.load(os.path.join(os.path.dirname(__file__), "golden", "stress_seed99.npz"))
    with api.Nnue(N.synthetic(99, stress=True).image, 0) as ctx:
        assert (ctx.eval_playouts(golden["boards"], golden["starts"]) == stress["evals"]).all()


def test_playout_walker_ragged_games(gpu_ctx, golden):
    """Empty games, single-board games and games cut mid-way."""
    boards, evals = golden["boards"], golden["evals"]
    starts = np.array([0, 0, 1, 1, 30, 81, 81, 100], dtype=np.uint32)
    # a cut game is still a legal chain as long as consecutive boards follow from each other
    out = gpu_ctx.eval_playouts(boards[:100], starts)
    assert (out == evals[:100]).all()


def test_walker_jump_between_unrelated_boards_falls_back_to_rebuild(gpu_ctx, golden):
    """A 'game' whose consecutive boards are unrelated (> 8 changed squares, or any king jump)."""
    idx = np.array([0, 500, 37, 2100, 36, 1200, 1201, 5, 3000], dtype=np.int64)
    boards = golden["boards"][idx]
    out = gpu_ctx.eval_playouts(boards, np.array([0, len(idx)], dtype=np.uint32))
    assert (out == golden["evals"][idx]).all()


def test_slots_refresh_update_eval(gpu_ctx, golden):
    boards, starts, evals = golden["boards"], golden["starts"], golden["evals"]
    n_games = len(starts) - 1
    gpu_ctx.slots_reserve(2 * n_games)
    slots = np.arange(n_games, dtype=np.uint32)
    gpu_ctx.refresh(slots, boards[starts[:-1]])
    assert (gpu_ctx.eval_slots(slots) == evals[starts[:-1]]).all()
    lengths = np.diff(starts)
    for ply in range(1, int(lengths.max())):
        live = np.nonzero(lengths > ply)[0]
        idx = starts[:-1][live] + ply
        # in-place update (datagen's applyImmediately form) fused with evaluation
        out = gpu_ctx.update_eval(slots[live], slots[live], boards[idx])
        assert (out == evals[idx]).all(), ply


def test_slots_push_pop_semantics(gpu_ctx, golden):
    """Search form: children are written to fresh slots, the parent slot stays valid (pop = reuse it)."""
    boards, starts, evals = golden["boards"], golden["starts"], golden["evals"]
    g = 3
    lo = int(starts[g])
    gpu_ctx.slots_reserve(64)
    gpu_ctx.refresh([0], boards[lo : lo + 1])
    # parent -> three different descendants in three slots, one call
    out = gpu_ctx.update_eval([0, 0, 0], [1, 2, 3], boards[[lo + 1, lo + 1, lo + 1]])
    assert (out == evals[lo + 1]).all()
    gpu_ctx.update([1], [4], boards[lo + 2 : lo + 3])
    assert gpu_ctx.eval_slots([4, 0, 1])[0] == evals[lo + 2]
    assert (gpu_ctx.eval_slots([0, 1]) == evals[[lo, lo + 1]]).all()  # parents untouched
    # explicit side to move (null-move children are evaluated on the parent's accumulators)
    stm_board = 0 if boards[lo]["stm_ep"] & 0x80 else 1
    flipped = gpu_ctx.eval_slots([0], stm=[1 - stm_board])
    same = gpu_ctx.eval_slots([0], stm=[stm_board])
    assert same[0] == evals[lo] and flipped[0] != INT32_MIN


def test_flipped_stm_matches_oracle(gpu_ctx, c_oracle, golden):
    board = golden["boards"][40:41].copy()
    gpu_ctx.slots_reserve(8)
    gpu_ctx.refresh([5], board)
    psq, thr = c_oracle.accumulators(board)
    bucket = (bin(int(board["occupancy"][0])).count("1") - 2) // 4
    for stm in (0, 1):
        assert gpu_ctx.eval_slots([5], stm=[stm])[0] == c_oracle.forward_acc(psq, thr, stm, bucket)


def test_read_slot_returns_logical_accumulators(gpu_ctx, c_oracle, golden):
    board = golden["boards"][123:124]
    gpu_ctx.slots_reserve(8)
    gpu_ctx.refresh([7], board)
    acc, stored = gpu_ctx.read_slot(7)
    psq, thr = c_oracle.accumulators(board)
    want = (psq.astype(np.int32) + thr.astype(np.int32)).astype(np.int16)  # only the wrapped sum reaches the net
    assert (acc == want).all()
    assert stored.tobytes() == board.tobytes()


def test_slot_errors(gpu_ctx, golden):
    gpu_ctx.slots_reserve(8)
    with pytest.raises(api.NnueError) as e:
        gpu_ctx.refresh([1 << 20], golden["boards"][:1])
    assert e.value.status == api.SP_ERR_INVALID
    with pytest.raises(api.NnueError) as e:
        # updating from a slot that was never filled is a malformed-board error, not garbage
        gpu_ctx.slots_reserve(4096)
        gpu_ctx.update([4095], [4094], golden["boards"][:1])
    assert e.value.status == api.SP_ERR_BAD_BOARD


def test_full_size_playouts_equal_full_refresh(gpu_ctx):
    """Config 3 size: 65,536 games.  Property: the incremental walker equals the from-scratch
    evaluator on every position (the reference's datagen assertion, at scale)."""
    boards, _, starts = api.playouts(4242, 65536, 80)
    full = gpu_ctx.eval_full(boards)
    inc = gpu_ctx.eval_playouts(boards, starts)
    assert (full == inc).all()


@pytest.mark.parametrize("env", [{"SP_NNUE_PLAN_REBUILDS": "0"}, {"SP_NNUE_PLAN_CAP": "7"}, {"SP_NNUE_GAMES_CHUNK": "5"}])
def test_walker_rebuild_plan_variants(net, golden, monkeypatch, env):
    """Rebuilds inside the walker, a rebuild plan that overflows after 7 items, and many small chunks
    sharing one plan: all bit-exact."""
    for k, v in env.items():
        monkeypatch.setenv(k, v)
    with api.Nnue(net.image, 0) as ctx:
        for key in ("", "dfrc_"):
            out = ctx.eval_playouts(golden[key + "boards"], golden[key + "starts"])
            assert (out == golden[key + "evals"]).all(), (env, key)
        # a second call reuses the plan scratch
        assert (ctx.eval_playouts(golden["boards"], golden["starts"]) == golden["evals"]).all()


@pytest.mark.parametrize("small", ["1", "copies", "0"])
def test_search_sized_rounds_small_kernel_and_general_kernels(net, c_oracle, golden, monkeypatch, small):
    """sp_nnue_batch and the slot entry points with a few dozen items: SP_NNUE_SMALL=1 (default) takes them through ONE fused
    launch (small_batch_kernel: slot update + activation + dp4a L1 + L2 + L3 per warp), SP_NNUE_SMALL=0 through the general
    kernels.  Both must give the reference's values -- here on the STRESS network, whose int16 / int32 sums wrap everywhere."""
    import os

    from stormphrax_b200 import net as N

    stress = np.load(os.path.join(os.path.dirname(__file__), "golden", "stress_seed99.npz"))
    boards, starts, evals = golden["boards"], golden["starts"], stress["evals"]
    assert (np.diff(starts) >= 3).all()
    monkeypatch.setenv("SP_NNUE_SMALL", "0" if small == "0" else "1")
    if small == "copies":  # the staged block goes through cudaMemcpyAsync instead of being read in place (rounds above 64 items)
        monkeypatch.setenv("SP_NNUE_SMALL_MAPPED", "0")
    with api.Nnue(N.synthetic(99, stress=True).image, 0) as ctx:
        n = len(starts) - 1
        first = starts[:-1].astype(np.int64)
        ctx.slots_reserve(3 * n)
        ids = np.arange(n, dtype=np.uint32)
        launches0 = int(ctx.counters()[api.CTR_LAUNCHES])
        r, _, _ = ctx.batch(refresh=(ids, boards[first]))
        assert (r == evals[first]).all()
        if small != "0":
            assert int(ctx.counters()[api.CTR_LAUNCHES]) - launches0 == 1  # one kernel for the whole round
        # one round with all three groups: refresh other slots, advance the first ones, evaluate-only with both sides
        stm_board = np.where(boards[first]["stm_ep"] & 0x80, 0, 1).astype(np.uint8)
        r, u, e = ctx.batch(refresh=(ids + 2 * n, boards[first + 2]), update=(ids, ids + n, boards[first + 1]), evaluate=(ids, stm_board))
        assert (r == evals[first + 2]).all() and (u == evals[first + 1]).all() and (e == evals[first]).all()
        # push / pop: the parents are untouched, children of children work, explicit side to move differs from the stored one
        assert (ctx.eval_slots(ids) == evals[first]).all()
        assert (ctx.update_eval(ids + n, ids + n, boards[first + 2]) == evals[first + 2]).all()
        flipped = ctx.eval_slots(ids[:4], stm=1 - stm_board[:4])
        c_oracle.load_net(N.synthetic(99, stress=True).image)  # the C oracle keeps ONE global network: put the session's back afterwards
        try:
            for k in range(4):
                b = boards[first[k] : first[k] + 1]
                psq, thr = c_oracle.accumulators(b)
                bucket = (bin(int(b["occupancy"][0])).count("1") - 2) // 4
                assert flipped[k] == c_oracle.forward_acc(psq, thr, int(1 - stm_board[k]), bucket)
        finally:
            c_oracle.load_net(net.image)
        # errors: a slot out of range, an update from a slot that was never filled
        with pytest.raises(api.NnueError) as err:
            ctx.batch(refresh=([1 << 30], boards[:1]))
        assert err.value.status == api.SP_ERR_INVALID
        ctx.slots_reserve(3 * n + 16)
        with pytest.raises(api.NnueError) as err:
            ctx.update_eval([3 * n + 8], [3 * n + 9], boards[:1])
        assert err.value.status == api.SP_ERR_BAD_BOARD
