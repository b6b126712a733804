"""GPU parity, full-refresh path (BASELINE.json config 2): every call goes through the C-ABI.

Bar: bit-exact.  Checked against (i) golden vectors produced by the reference's own code and
(ii) the pinned C oracle on seeded inputs, plus size-independent properties at the full 1M size.
"""
import numpy as np
import pytest

from stormphrax_b200 import api
from stormphrax_b200 import net as N

pytestmark = pytest.mark.gpu

INT32_MIN = np.iinfo(np.int32).min


def _dev(a: np.ndarray):
    import torch

    _stream()
    t = torch.from_numpy(np.ascontiguousarray(a).view(np.uint8).reshape(-1)).cuda()
    torch.cuda.synchronize()
    return t


_STREAM = None


def _stream() -> int:
    """A non-default torch stream (the library maps a NULL stream to the context's own)."""
    import torch

    global _STREAM
    if _STREAM is None:
        _STREAM = torch.cuda.Stream()
        torch.cuda.set_stream(_STREAM)
    return _STREAM.cuda_stream


def test_golden_playouts_bit_exact(gpu_ctx, golden):
    assert (gpu_ctx.eval_full(golden["boards"]) == golden["evals"]).all()


def test_golden_dfrc_and_special_fens_bit_exact(gpu_ctx, golden):
    assert (gpu_ctx.eval_full(golden["dfrc_boards"]) == golden["dfrc_evals"]).all()
    assert (gpu_ctx.eval_full(golden["fen_boards"]) == golden["fen_evals"]).all()


def test_stress_network_bit_exact(golden):
    """Full-range weights: every int16 / int32 wrap-around path is taken."""
    import os

    stress = np.load(os.path.join(os.path.dirname(__file__), "golden", "stress_seed99.npz"))
    with api.Nnue(N.synthetic(99, stress=True).image, 0) as ctx:
        assert (ctx.eval_full(golden["boards"]) == stress["evals"]).all()
        assert (ctx.eval_full(golden["fen_boards"]) == stress["fen_evals"]).all()


def test_ft_activations_match_oracle(gpu_ctx, c_oracle, small_playouts):
    import torch

    boards = small_playouts[0][::9]
    n = len(boards)
    d_boards = _dev(boards)
    d_act = torch.empty(n * 1024, dtype=torch.uint8, device="cuda")
    d_bucket = torch.empty(n, dtype=torch.uint8, device="cuda")
    s = _stream()
    gpu_ctx.activations_device(d_boards, n, d_act, d_bucket, s)
    gpu_ctx.sync(s)
    act = d_act.cpu().numpy().reshape(n, 1024)
    bucket = d_bucket.cpu().numpy()
    for i in range(n):
        want, b = c_oracle.ft_activations(boards[i])
        assert bucket[i] == b
        assert (act[i] == want).all(), i


HEAD_KERNELS = [("umma", "2048"), ("umma", "0"), ("stream", "0"), ("tiles", "0")]


@pytest.mark.parametrize("kernel,direct", HEAD_KERNELS)
def test_dense_head_matches_oracle_on_arbitrary_activations(net, c_oracle, monkeypatch, kernel, direct):
    """Config 4's isolated head: full u8 range (the FT only produces 0..127), all buckets, ragged n.  Every head kernel: the
    warp-per-position kernel small launches take by default (SP_NNUE_HEAD_DIRECT=2048), and with that switched off the sorted
    tensor-core kernels -- head_umma_kernel (tcgen05 L1, default), head_stream_kernel and head_kernel (mma.sync L1)."""
    import torch

    monkeypatch.setenv("SP_NNUE_HEAD", kernel)
    monkeypatch.setenv("SP_NNUE_HEAD_DIRECT", direct)
    rng = np.random.default_rng(3)
    with api.Nnue(net.image, 0) as ctx:
        for n in (1, 15, 16, 17, 333, 1500):
            act = rng.integers(0, 256, (n, 1024), dtype=np.uint8)
            act[0] = 0
            if n > 1:
                act[1] = 255
            bucket = rng.integers(0, 8, n, dtype=np.uint8)
            if n > 100:
                bucket[7] = 0xFF  # a board the feature transformer rejected
            d_out = torch.empty(n, dtype=torch.int32, device="cuda")
            s = _stream()
            ctx.forward_device(_dev(act), _dev(bucket), n, d_out, s)
            ctx.sync(s)
            got = d_out.cpu().numpy()
            want = np.array([c_oracle.forward(act[i], int(bucket[i])) if bucket[i] < 8 else INT32_MIN for i in range(n)], dtype=np.int32)
            assert (got == want).all(), n


def test_matches_oracle_on_seeded_playouts(gpu_ctx, c_oracle, small_playouts):
    boards = small_playouts[0]
    assert (gpu_ctx.eval_full(boards) == c_oracle.eval_once(boards)).all()


def test_empty_and_ragged_batches(gpu_ctx, golden):
    assert gpu_ctx.eval_full(golden["boards"][:0]).size == 0
    for n in (1, 2, 7, 16, 17, 31, 33, 257):
        assert (gpu_ctx.eval_full(golden["boards"][:n]) == golden["evals"][:n]).all(), n


def test_malformed_boards_are_reported_not_evaluated(gpu_ctx, golden):
    boards = golden["boards"][:40].copy()
    boards["occupancy"][5] = 0  # no kings at all
    boards["pieces"][9] = 0x77  # piece code 7 does not exist
    out = np.zeros(40, dtype=np.int32)
    with pytest.raises(api.NnueError) as e:
        gpu_ctx.eval_full(boards, out)
    assert e.value.status == api.SP_ERR_BAD_BOARD
    good = np.ones(40, dtype=bool)
    good[[5, 9]] = False
    assert (out[good] == golden["evals"][:40][good]).all()
    assert (out[~good] == INT32_MIN).all()
    # the context stays usable and the error does not stick
    assert (gpu_ctx.eval_full(golden["boards"][:40]) == golden["evals"][:40]).all()


def test_device_pointer_entry_point(gpu_ctx, golden):
    import torch

    boards = golden["boards"]
    d_out = torch.empty(len(boards), dtype=torch.int32, device="cuda")
    s = _stream()
    gpu_ctx.eval_full_device(_dev(boards), len(boards), d_out, s)
    gpu_ctx.sync(s)
    assert (d_out.cpu().numpy() == golden["evals"]).all()


def test_counters(net, golden):
    with api.Nnue(net.image, 0) as ctx:
        ctx.eval_full(golden["boards"][:100])
        c = ctx.counters()
        assert c[api.CTR_EVALS] == 100 and c[api.CTR_FULL_REFRESH] == 200 and c[api.CTR_LAUNCHES] >= 2


def test_full_size_properties(gpu_ctx, c_oracle):
    """BASELINE config 2 size (1,048,576 positions): permutation equivariance, duplicate
    consistency, and a random sample against the oracle."""
    boards, _, _ = api.playouts(42, 13200, 80)
    assert len(boards) >= 1 << 20
    boards = boards[: 1 << 20]
    out = gpu_ctx.eval_full(boards)
    rng = np.random.default_rng(0)
    perm = rng.permutation(len(boards))
    assert (gpu_ctx.eval_full(boards[perm]) == out[perm]).all()
    sample = rng.choice(len(boards), 3000, replace=False)
    assert (c_oracle.eval_once(boards[sample]) == out[sample]).all()
    # identical records evaluate identically wherever they sit in the batch
    start = boards[0].tobytes()
    same = np.array([b.tobytes() == start for b in boards[:: 81][:200]])
    assert same.sum() > 1 and len(set(out[::81][:200][same].tolist())) == 1


def test_split_extract_accumulate_path_bit_exact(net, golden, monkeypatch):
    """SP_NNUE_SPLIT=1: boards -> row lists (extract_kernel) -> TMA-staged accumulate_kernel -> head."""
    monkeypatch.setenv("SP_NNUE_SPLIT", "1")
    monkeypatch.setenv("SP_NNUE_CHUNK", "1024")  # several chunks, double-buffered row lists
    with api.Nnue(net.image, 0) as ctx:
        assert (ctx.eval_full(golden["boards"]) == golden["evals"]).all()
        assert (ctx.eval_full(golden["fen_boards"]) == golden["fen_evals"]).all()
        bad = golden["boards"][:20].copy()
        bad["occupancy"][3] = 0
        out = np.zeros(20, dtype=np.int32)
        with pytest.raises(api.NnueError):
            ctx.eval_full(bad, out)
        assert out[3] == INT32_MIN and (np.delete(out, 3) == np.delete(golden["evals"][:20], 3)).all()


def test_small_chunks_exercise_the_overlap_pipeline(net, golden, monkeypatch):
    """Many tiny chunks: double-buffered scratch, auxiliary-stream head, per-chunk uploads/downloads."""
    monkeypatch.setenv("SP_NNUE_CHUNK", "256")
    monkeypatch.setenv("SP_NNUE_GAMES_CHUNK", "3")
    with api.Nnue(net.image, 0) as ctx:
        for _ in range(3):
            assert (ctx.eval_full(golden["boards"]) == golden["evals"]).all()
            assert (ctx.eval_playouts(golden["boards"], golden["starts"]) == golden["evals"]).all()


def test_adjust_eval_matches_reference_golden(gpu_ctx):
    """adjustStatic + adjustEval<false> (eval.cpp:25-67) as a device epilogue, against
    values produced by the reference's own staticEvalOnce / adjustEval."""
    import os

    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "adjust_seed42.npz"))
    boards, raw = g["adjust_boards"], g["adjust_raw"]
    assert (gpu_ctx.eval_full(boards) == raw).all()
    for k in range(3):
        c0, c1, o0, o1 = (int(x) for x in g[f"adjust_params{k}"])
        params = api.AdjustParams.defaults(contempt=(c0, c1), optimism=(o0, o1))
        assert (gpu_ctx.adjust(boards, raw, params) == g[f"adjust_out{k}"]).all(), k
    # caller-supplied correction (the reference reads it from its history tables): + trunc(c / 2048), then clamp
    rng = np.random.default_rng(1)
    corr = rng.integers(-300000, 300000, len(boards)).astype(np.int32)
    base = g["adjust_out0"].astype(np.int64)
    want = np.clip(base + np.trunc(corr / 2048).astype(np.int64), -24999, 24999)
    inside = np.abs(base) < 24999  # where the golden value was not clamped, the pre-clamp value is known
    got = gpu_ctx.adjust(boards, raw, api.AdjustParams.defaults(), correction=corr)
    assert (got[inside] == want[inside]).all()


@pytest.mark.parametrize("kernel", ["umma", "stream"])
def test_dense_head_large_mixed_buckets(net, c_oracle, monkeypatch, kernel):
    """Config-4 size with rows of all eight buckets interleaved at random plus rejected rows: exercises the
    bucket grouping (counting sort, group padding, CTAs that straddle two groups) and the byte-limb L2."""
    import torch

    monkeypatch.setenv("SP_NNUE_HEAD", kernel)
    gpu_ctx = api.Nnue(net.image, 0)
    rng = np.random.default_rng(11)
    n = (1 << 16) + 17
    act = rng.integers(0, 128, (n, 1024), dtype=np.uint8)
    bucket = rng.integers(0, 8, n, dtype=np.uint8)
    bucket[rng.choice(n, 50, replace=False)] = 0xFF  # boards the FT rejected
    bucket[:300] = 3                                  # a long single-bucket run
    d_out = torch.empty(n, dtype=torch.int32, device="cuda")
    s = _stream()
    gpu_ctx.forward_device(_dev(act), _dev(bucket), n, d_out, s)
    gpu_ctx.sync(s)
    got = d_out.cpu().numpy()
    assert (got[bucket == 0xFF] == INT32_MIN).all()
    sample = rng.choice(np.nonzero(bucket != 0xFF)[0], 1500, replace=False)
    want = np.array([c_oracle.forward(act[i], int(bucket[i])) for i in sample], dtype=np.int32)
    assert (got[sample] == want).all()
    # the same rows in another order give the same values (grouping is order-independent)
    perm = rng.permutation(n)
    gpu_ctx.forward_device(_dev(act[perm]), _dev(bucket[perm]), n, d_out, s)
    gpu_ctx.sync(s)
    assert (d_out.cpu().numpy() == got[perm]).all()
    # a bucket array that starts at an odd address (the counting sort reads it 16 bytes at a time from the aligned address below)
    shifted = _dev(np.concatenate([np.full(5, 0xEE, dtype=np.uint8), bucket]))[5:]
    d_out.zero_()
    gpu_ctx.forward_device(_dev(act), shifted, n, d_out, s)
    gpu_ctx.sync(s)
    assert (d_out.cpu().numpy() == got).all()
    gpu_ctx.close()


@pytest.mark.parametrize("kernel,direct", [("umma", "0"), ("stream", "0"), ("umma", "2048")])
def test_dense_head_l2_paths(net, stress_net, monkeypatch, kernel, direct):
    """The streaming head picks its L2 form per tile: int16-range weights + inputs below 2^16 (two limbs each),
    the general four-limb form without the zero input limbs, and the full form after a wrapped square.
    All three must agree with the oracle: normal net (narrow weights) with FT-range and full-range activations,
    the same net with the narrow form switched off, and the stress net (full-range weights, wrapping squares)."""
    import torch

    from oracle.bind import COracle

    monkeypatch.setenv("SP_NNUE_HEAD", kernel)
    monkeypatch.setenv("SP_NNUE_HEAD_DIRECT", direct)
    rng = np.random.default_rng(17)
    n = 2000
    acts = {"ft_range": rng.integers(0, 128, (n, 1024), dtype=np.uint8), "full_range": rng.integers(0, 256, (n, 1024), dtype=np.uint8)}
    bucket = rng.integers(0, 8, n, dtype=np.uint8)
    s = _stream()
    oracle = COracle()
    try:
        for network, narrow in ((net, "1"), (net, "0"), (stress_net, "1")):
            monkeypatch.setenv("SP_NNUE_L2_NARROW", narrow)
            oracle.load_net(network.image)
            with api.Nnue(network.image, 0) as ctx:
                for name, act in acts.items():
                    d_out = torch.empty(n, dtype=torch.int32, device="cuda")
                    ctx.forward_device(_dev(act), _dev(bucket), n, d_out, s)
                    ctx.sync(s)
                    got = d_out.cpu().numpy()
                    pick = rng.choice(n, 400, replace=False)
                    want = np.array([oracle.forward(act[i], int(bucket[i])) for i in pick], dtype=np.int32)
                    assert (got[pick] == want).all(), (narrow, name)
    finally:
        oracle.load_net(net.image)  # the C oracle keeps one global network


def test_zstd_flagged_network_loads_and_evaluates_identically(net, golden):
    """sp_nnue_create on a zstd-compressed image (eval::init, src/eval/nnue.cpp:215-247) = the raw network."""
    pytest.importorskip("pyarrow")
    with api.Nnue(N.compressed(net), 0) as ctx:
        assert (ctx.eval_full(golden["boards"][:512]) == golden["evals"][:512]).all()
    damaged = N.compressed(net)[:100000]
    with pytest.raises(api.NnueError) as e:
        api.Nnue(damaged, 0)
    assert e.value.status == api.SP_ERR_BAD_NETWORK


def test_wdl_on_device_matches_reference_golden(gpu_ctx):
    """sp_nnue_wdl = wdl::normalizeScore<false> (bit-exact) and wdl::wdlModel (per mille; exp() may differ in its last
    bit between libm and the device: tolerance 1) on the reference-generated vectors (tests/golden/make_datagen_golden.py)."""
    import os

    d = np.load(os.path.join(os.path.dirname(__file__), "golden", "datagen_seed42.npz"))
    norm, win, loss = gpu_ctx.wdl(d["norm_boards"], d["norm_scores"])
    assert np.array_equal(norm, d["norm_out"].astype(np.int32))
    want = d["wdl_model"].astype(np.int64)
    assert np.abs(win.astype(np.int64) - want[:, 0]).max() <= 1 and np.abs(loss.astype(np.int64) - want[:, 1]).max() <= 1
    only_norm, none_w, none_l = gpu_ctx.wdl(d["norm_boards"], d["norm_scores"], model=False)
    assert np.array_equal(only_norm, norm) and none_w is None and none_l is None


def test_warp_per_position_kernel_still_bit_exact(net, golden, monkeypatch):
    """SP_NNUE_FT=warp: ft_full_kernel (one warp per position, ALU summation) instead of the tensor-core group kernel."""
    monkeypatch.setenv("SP_NNUE_FT", "warp")
    with api.Nnue(net.image, 0) as ctx:
        assert (ctx.eval_full(golden["boards"]) == golden["evals"]).all()
        assert (ctx.eval_full(golden["dfrc_boards"]) == golden["dfrc_evals"]).all()


def test_group_kernel_on_shuffled_positions_and_stress_network(golden):
    """The group kernel must not depend on neighbours sharing rows: shuffled positions (large unions), ragged tails,
    and the full-range network (every int16 wrap of the byte-plane recombination)."""
    import os

    stress = np.load(os.path.join(os.path.dirname(__file__), "golden", "stress_seed99.npz"))
    rng = np.random.default_rng(5)
    perm = rng.permutation(len(golden["boards"]))
    with api.Nnue(N.synthetic(99, stress=True).image, 0) as ctx:
        for n in (len(perm), 1, 15, 16, 17, 1000):
            assert (ctx.eval_full(golden["boards"][perm[:n]]) == stress["evals"][perm[:n]]).all(), n


def test_group_kernel_overflow_path(net, golden, monkeypatch):
    """SP_NNUE_GROUP_LIMIT caps a group's threat-row union far below what 16 positions need: most groups overflow and are
    redone by the per-position kernel; results stay exact and malformed boards are still reported."""
    monkeypatch.setenv("SP_NNUE_GROUP_LIMIT", "70")
    with api.Nnue(net.image, 0) as ctx:
        assert (ctx.eval_full(golden["boards"]) == golden["evals"]).all()
        bad = golden["boards"][:40].copy()
        bad["occupancy"][19] = 0
        out = np.zeros(40, dtype=np.int32)
        with pytest.raises(api.NnueError):
            ctx.eval_full(bad, out)
        assert out[19] == INT32_MIN and (np.delete(out, 19) == np.delete(golden["evals"][:40], 19)).all()
