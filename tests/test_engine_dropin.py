"""The drop-in claim, tested with the reference engine's OWN search: oracle/engine/Makefile compiles the reference's
search.cpp / thread.cpp / position.cpp / bench.cpp / datagen.cpp / uci.cpp (where they lie, unmodified) twice -- against
its stock CPU evaluation (sp_engine_cpu) and against this library through the adapter of INTEGRATION.md section 3
(sp_engine_b200).  Evaluations are exact integers, so both engines must walk the same trees: identical `bench` node counts
(src/bench.cpp:149-150, the engine's own determinism signature) and identical evaluation checksums along playouts, with
datagen's invariant staticEvalOnce == staticEval(nnueState) (src/datagen/datagen.cpp:262) holding at every ply.
"""
import json
import os
import re
import subprocess

import pytest

from stormphrax_b200 import net as N

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF_DIR = os.path.join(ROOT, "oracle", "_ref")
GOLDEN = os.path.join(ROOT, "tests", "golden", "engine_seed7.json")


def _engine(name):
    path = os.path.join(REF_DIR, name)
    if not os.path.exists(path):
        pytest.skip(f"{name} not built (make -C oracle/engine, where /root/reference exists)")
    return path


@pytest.fixture(scope="module")
def golden():
    with open(GOLDEN) as f:
        return json.load(f)


@pytest.fixture(scope="module")
def net_path(tmp_path_factory, golden):
    path = tmp_path_factory.mktemp("engine") / "tame.nnue"
    N.synthetic(golden["net"]["seed"], tame=True).image.tofile(path)
    return str(path)


def _run(engine, net_path, *args, timeout=600):
    p = subprocess.run([engine, net_path, *map(str, args)], capture_output=True, text=True, timeout=timeout)
    assert p.returncode == 0, p.stdout[-2000:] + p.stderr[-2000:]
    return p.stdout


def _bench_nodes(out):
    return int(re.search(r"^(\d+) nodes (\d+) nps", out, re.M).group(1))


def _evalcheck(out):
    m = re.search(r"evalcheck: (\d+) positions, (\d+) mismatches, checksum ([0-9a-f]+)", out)
    return int(m.group(1)), int(m.group(2)), m.group(3)


def test_reference_engine_cpu_reproduces_its_golden_signature(net_path, golden):
    """The CPU build of the reference engine is deterministic: the committed golden values are what it prints here."""
    from oracle.bind import ref_isa_available

    if "avx2" not in ref_isa_available():
        pytest.skip("host cannot run the AVX2 build of the reference")
    cpu = _engine("sp_engine_cpu")
    assert _bench_nodes(_run(cpu, net_path, "bench", 1)) == golden["bench_nodes"]["1"]
    e = golden["evalcheck"]
    assert _evalcheck(_run(cpu, net_path, "evalcheck", e["games"], e["seed"])) == (e["positions"], 0, e["checksum"])


@pytest.mark.gpu
def test_reference_search_on_the_b200_evaluator_walks_the_same_tree(net_path, golden):
    """The reference's own search, every static evaluation answered by the GPU: same node counts as on the CPU."""
    b200 = _engine("sp_engine_b200")
    for depth in ("1", "2"):
        assert _bench_nodes(_run(b200, net_path, "bench", depth)) == golden["bench_nodes"][depth], f"depth {depth}"


@pytest.mark.gpu
def test_datagen_invariant_and_eval_checksum_on_the_b200_evaluator(net_path, golden):
    """applyMove<BoardObserver> + applyImmediately + staticEval == staticEvalOnce at every ply (datagen.cpp:257-262), and the
    evaluations are the CPU engine's, value for value (checksum)."""
    b200 = _engine("sp_engine_b200")
    e = golden["evalcheck"]
    assert _evalcheck(_run(b200, net_path, "evalcheck", e["games"], e["seed"])) == (e["positions"], 0, e["checksum"])


def _searches(out):
    return [int(x) for x in re.search(r"search nodes:((?: \d+)+)", out).group(1).split()]


def _games(out):
    m = re.search(r"games: .* nodes: (\d+) nodes .* checksum ([0-9a-f]+) rounds (\d+) evaluations (\d+)", out)
    return int(m.group(1)), m.group(2), int(m.group(3)), int(m.group(4))


def test_reference_engine_cpu_reproduces_its_search_and_game_goldens(net_path, golden):
    from oracle.bind import ref_isa_available

    if "avx2" not in ref_isa_available():
        pytest.skip("host cannot run the AVX2 build of the reference")
    cpu = _engine("sp_engine_cpu")
    s, g = golden["searches"], golden["games"]
    assert _searches(_run(cpu, net_path, "searches", s["n"], s["depth"], 0)) == s["nodes"]
    nodes, checksum, _, _ = _games(_run(cpu, net_path, "games", g["n"], g["soft_nodes"], g["plies"], g["seed"], 0))
    assert (nodes, checksum) == (g["nodes"], g["checksum"])


@pytest.mark.gpu
def test_batched_reference_searches_as_fibers_walk_the_same_trees(net_path, golden):
    """SURVEY 8 f4, "host fibers first": many searches of the UNMODIFIED reference Searcher run as fibers of one host thread;
    every NnueState::evaluate yields, and each round's requests are answered by one sp_nnue_batch submission.  Node counts per
    search must equal the CPU engine's, searched one after the other."""
    b200 = _engine("sp_engine_b200")
    s = golden["searches"]
    assert _searches(_run(b200, net_path, "searches", s["n"], s["depth"], 1)) == s["nodes"]
    assert _searches(_run(b200, net_path, "searches", s["n"], s["depth"], 0)) == s["nodes"]  # the synchronous path too


@pytest.mark.gpu
def test_batched_reference_selfplay_games_as_fibers(net_path, golden):
    """Concurrent self-play games with datagen's per-move search (runDatagenSearch + applyImmediately, datagen.cpp:206-260):
    same moves and scores as on the CPU (checksum), evaluations answered in batches."""
    b200 = _engine("sp_engine_b200")
    g = golden["games"]
    nodes, checksum, rounds, evaluations = _games(_run(b200, net_path, "games", g["n"], g["soft_nodes"], g["plies"], g["seed"], 1))
    assert (nodes, checksum) == (g["nodes"], g["checksum"])
    assert rounds > 0 and evaluations / rounds > g["n"] / 2  # really batched: most games contribute to a round


@pytest.mark.gpu
def test_fiber_schedulers_on_several_host_threads(net_path, golden):
    """datagen's "N threads" (datagen.cpp:378-384) on one GPU: the searches / games are shared out to host threads, each a fiber
    scheduler with its own evaluator context (eval::createContext: network copy, stream, slot store) submitting its own batches.
    Same trees, whichever thread and context a search ran on."""
    b200 = _engine("sp_engine_b200")
    s, g = golden["searches"], golden["games"]
    assert _searches(_run(b200, net_path, "searches", s["n"], s["depth"], 1, 3)) == s["nodes"]
    # four searches alive per scheduler, the others queued: finished fibers take the next search and its slots are reused
    assert _searches(_run(b200, net_path, "searches", s["n"], s["depth"], 1, 3, 4)) == s["nodes"]
    assert _searches(_run(b200, net_path, "searches", s["n"], s["depth"], 1, 1, 5)) == s["nodes"]
    nodes, checksum, rounds, evaluations = _games(_run(b200, net_path, "games", g["n"], g["soft_nodes"], g["plies"], g["seed"], 1, 4))
    assert (nodes, checksum) == (g["nodes"], g["checksum"])
    assert rounds > 0 and evaluations / rounds > g["n"] / 4 / 2  # four schedulers of four games each, still batched


def test_reference_engine_cpu_threads_share_the_jobs(net_path, golden):
    """The CPU build's counterpart of the threaded run (plain host threads over its own evaluation): same node counts."""
    from oracle.bind import ref_isa_available

    if "avx2" not in ref_isa_available():
        pytest.skip("host cannot run the AVX2 build of the reference")
    cpu = _engine("sp_engine_cpu")
    s = golden["searches"]
    assert _searches(_run(cpu, net_path, "searches", s["n"], s["depth"], 1, 3)) == s["nodes"]
