#!/usr/bin/env python
"""bench.py -- throughput of the batched NNUE hot path (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload full|playouts]
                    [--extras all|none]

A "step" is one pass of the hot path over one batch of synthetic input.  Default workload =
BASELINE.json configs[1]: full-refresh evaluation of 1,048,576 random legal positions (first
<= 80 plies of random playouts from the start position) per GPU.  One process per GPU
(torchrun sets RANK/LOCAL_RANK/WORLD_SIZE); positions shard by rank with no data-path
collective; NCCL all-reduces only the reporting counters, the parity counts and the
max-over-ranks step time.

Prints ONE JSON line (rank 0):
  value         device-resident throughput (CUDA events on the launch stream)
  e2e           the same through the host-pointer C-ABI call, pinned host buffers (H2D of the packed
                boards and D2H of the evals inside the timed region); e2e.pageable = pageable buffers
  parity        every evaluation of the timed workload compared with the reference's own CPU code
                (oracle/_ref, else the C port) on the same positions; a mismatch exits non-zero
  roofline      the dominant kernel's achieved algorithmic bytes/s against the measured HBM peak, plus
                the on-chip figure that actually binds the cache-resident kernels
  cpu_baseline  the reference's CPU path timed on this box's host cores
  workloads     the other BASELINE configs, one short record each: playouts (configs[2]), head_sweep
                (configs[3]), slots (the NnueState drop-in path: sp_nnue_batch update + evaluate rounds)
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import threading

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "Mpositions/sec batched NNUE (bit-exact vs reference CPU path)"
UNIT = "Mpos/s"
POSITIONS_PER_GPU = 1 << 20   # full refresh: BASELINE configs[1]
PLAYOUTS_PER_GPU = 1 << 16    # incremental: BASELINE configs[2], 65,536 playouts of <= 80 plies (about 5.26 M positions)
SLOT_STATES = 1 << 16         # slots workload: one NnueState-style accumulator chain per playout
SLOT_ROUNDS = 16              # update + evaluate rounds timed per step
MAX_PLIES = 80
NET_SEED = 1234
L2_FLUSH_BYTES = 256 << 20


def log(*a):
    print(*a, file=sys.stderr, flush=True)


def measured_peak_hbm():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        with open(path) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


def measured_onchip():
    """On-chip bandwidths measured by tools/ubench.cu on a B200 (profiles/onchip_peaks.json), or None."""
    try:
        with open(os.path.join(ROOT, "profiles", "onchip_peaks.json")) as f:
            return json.load(f)
    except Exception:
        return None


def make_workload(rank: int, n_positions: int, workload: str = "full"):
    """Deterministic synthetic positions for this rank: whole games (playouts), or whole games trimmed
    to n_positions (full refresh)."""
    from stormphrax_b200 import api

    if workload == "playouts":
        return api.playouts(42 + 1000003 * rank, PLAYOUTS_PER_GPU, MAX_PLIES)
    n_games = n_positions // 78 + 64  # random playouts rarely end before ply 80
    while True:
        boards, moves, starts = api.playouts(42 + 1000003 * rank, n_games, MAX_PLIES)
        if len(boards) >= n_positions:
            break
        n_games = int(n_games * 1.1) + 1
    g = int(np.searchsorted(starts, n_positions, side="right"))  # games fully or partly inside
    starts = starts[: g + 1].copy()
    starts[-1] = n_positions
    return boards[:n_positions].copy(), moves[:n_positions].copy(), starts


class ClockSampler:
    """SM clock + throttle reasons sampled via NVML while the timed region runs."""

    def __init__(self, index: int):
        self.samples, self.reasons, self.max_mhz = [], set(), None
        self._stop = threading.Event()
        self._thread = None
        try:
            import pynvml

            pynvml.nvmlInit()
            self._nv = pynvml
            self._h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self._h, pynvml.NVML_CLOCK_SM)
        except Exception as e:  # pragma: no cover
            self._nv = None
            log("clock sampling unavailable:", e)

    def _run(self):
        nv = self._nv
        names = {
            nv.nvmlClocksThrottleReasonSwPowerCap: "sw_power_cap",
            nv.nvmlClocksThrottleReasonHwSlowdown: "hw_slowdown",
            nv.nvmlClocksThrottleReasonSwThermalSlowdown: "sw_thermal_slowdown",
            nv.nvmlClocksThrottleReasonHwThermalSlowdown: "hw_thermal_slowdown",
            nv.nvmlClocksThrottleReasonHwPowerBrakeSlowdown: "hw_power_brake",
            nv.nvmlClocksThrottleReasonApplicationsClocksSetting: "applications_clocks_setting",
        }
        while not self._stop.is_set():
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self._h, nv.NVML_CLOCK_SM))
                mask = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self._h)
                for bit, name in names.items():
                    if mask & bit:
                        self.reasons.add(name)
            except Exception:
                pass
            self._stop.wait(0.05)

    def __enter__(self):
        if self._nv:
            self._thread = threading.Thread(target=self._run, daemon=True)
            self._thread.start()
        return self

    def __exit__(self, *exc):
        self._stop.set()
        if self._thread:
            self._thread.join()

    def summary(self):
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons)}
        return {"sm_mhz": float(np.median(self.samples)), "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons)}


# ------------------------------------------------------------------ CPU side (reference arm / cpu_baseline / parity checker)

def cpu_checker():
    """The reference's own CPU code (oracle/_ref) when it can run here, else the plain-C port."""
    from oracle.bind import COracle, Reference
    from stormphrax_b200 import net as N

    image = N.synthetic(NET_SEED).image
    if Reference.available():
        ref = Reference()
        ref.load_net(image)
        return ref, "reference", f"avx{'512' if ref.isa == 'avx512' else '2'}"
    ref = COracle()
    ref.load_net(image)
    return ref, "port", "scalar C"


def cpu_arm(boards, moves, starts, workload: str, budget_s: float, threads: int | None = None, complete: bool = False):
    """Time the CPU checker on a bounded prefix of the workload using every host core (or `threads`).
    Returns (pos_per_s, info, evaluations of that prefix)."""
    ref, kind, isa = cpu_checker()
    cores = threads or os.cpu_count() or 1
    if workload == "playouts" and kind == "reference":
        probe_games = min(len(starts) - 1, 4 * cores)
        t, _ = ref.time_playouts(boards[: starts[probe_games]], moves[: starts[probe_games]], starts[: probe_games + 1], cores, 1)
        rate = starts[probe_games] / t
        games = int(min(len(starts) - 1, max(probe_games, rate * budget_s / 81)))
        n = int(starts[games])
        t, out = ref.time_playouts(boards[:n], moves[:n], starts[: games + 1], cores, 1)
        sample = f"first {games} playouts ({n} positions), incremental applyMove+applyImmediately+evaluate, 1 pass"
    else:
        probe = min(len(boards), 512 * cores)
        t, _ = ref.time_eval_once(boards[:probe], cores, 1)
        rate = probe / t
        n = int(min(len(boards), max(probe, rate * budget_s)))
        t, out = ref.time_eval_once(boards[:n], cores, 1)
        sample = f"first {n} positions of the workload, evaluateOnce, 1 pass"
    out = out[:n]
    if complete and n < len(boards):
        # the timed sample was a prefix (slow host): evaluate the rest untimed, so that parity covers every position
        if workload == "playouts" and kind == "reference":
            _, rest = ref.time_playouts(boards[n:], moves[n:], (starts[games:] - starts[games]).astype(np.uint32), cores, 1)
        else:
            _, rest = ref.time_eval_once(boards[n:], cores, 1)
        out = np.concatenate([out, rest])
    return n / t, {"kind": kind, "cores": cores, "isa": isa, "sample": sample, "seconds": round(t, 3), "positions": n}, out


def workload_config(workload: str):
    """Identical in both arms: names the workload and nothing that depends on the run."""
    return {
        "workload": ("full-refresh NNUE eval of 1,048,576 random legal positions per GPU (BASELINE configs[1])"
                     if workload == "full" else
                     "incremental NNUE eval along 65,536 random playouts of <= 80 plies per GPU (BASELINE configs[2])"),
        "positions_per_gpu": POSITIONS_PER_GPU if workload == "full" else None,
        "playouts_per_gpu": PLAYOUTS_PER_GPU if workload != "full" else None,
        "max_plies": MAX_PLIES,
        "network": f"synthetic CBNF, numpy default_rng({NET_SEED}), arch 16x704+64368 -> 1024x2 -> 32 -> 64 -> 1 x8",
        "parallelism": "positions sharded by rank, network replicated, no data-path collective",
        "l2": "L2 flushed (256 MiB write) between timed steps",
    }


def run_reference(args, rank: int, world: int):
    if rank != 0:
        return
    boards, moves, starts = make_workload(0, POSITIONS_PER_GPU, args.workload)
    budget = min(12.0, 120.0 / max(args.steps, 1))
    rates, info = [], None
    for i in range(args.warmup + args.steps):
        rate, info, _ = cpu_arm(boards, moves, starts, args.workload, budget if i >= args.warmup else 2.0)
        if i >= args.warmup:
            rates.append(rate)
    value = float(np.mean(rates)) / 1e6
    # BASELINE configs[0]'s analogue: the same reference code on ONE host thread
    one_rate, one_info, _ = cpu_arm(boards, moves, starts, args.workload, 3.0, threads=1)
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * info["seconds"], "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "int16/int8/int32", "data": "synthetic",
        "config": workload_config(args.workload),
        "cpu_baseline": {"value": value, "unit": UNIT, **{k: info[k] for k in ("cores", "kind", "sample", "isa")},
                         "single_thread": {"value": one_rate / 1e6, "unit": UNIT, "sample": one_info["sample"]}},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    if args.extras != "none" and args.workload == "full":
        # configs[2] on the CPU: the reference's incremental NnueState path over random playouts
        from stormphrax_b200 import api

        pb, pm, ps = api.playouts(42, 4096, MAX_PLIES)
        prate, pinfo, _ = cpu_arm(pb, pm, ps, "playouts", 6.0)
        line["workloads"] = {"playouts": {"value": prate / 1e6, "unit": UNIT, **{k: pinfo[k] for k in ("cores", "kind", "sample", "isa")}}}
        line["workloads"]["engine_bench"] = engine_bench()
        line["workloads"]["engine_games"] = engine_games()
    print(json.dumps(line), flush=True)


def engine_bench(depth: int = 3):
    """BASELINE configs[0]: the reference's built-in `bench` (src/bench.cpp:95-170) on ONE CPU thread -- its own search on
    its own CPU evaluation (oracle/_ref/sp_engine_cpu, built from the unmodified sources by oracle/engine/Makefile).  The
    network is the tame synthetic one (scores a search can work with); the node count is this build's determinism
    signature, not comparable with upstream's (different network)."""
    import re
    import subprocess
    import tempfile

    from stormphrax_b200 import net as N

    engine = os.path.join(ROOT, "oracle", "_ref", "sp_engine_cpu")
    if not os.path.exists(engine):
        return {"unavailable": "oracle/_ref/sp_engine_cpu not built"}
    with tempfile.TemporaryDirectory() as tmp:
        path = os.path.join(tmp, "tame.nnue")
        N.synthetic(7, tame=True).image.tofile(path)
        try:
            out = subprocess.run([engine, path, "bench", str(depth)], capture_output=True, text=True, timeout=300).stdout
            m = re.search(r"^(\d+) nodes (\d+) nps", out, re.M)
            return {"config": f"`bench` depth {depth}, 52 positions, 16 MiB hash, 1 thread (BASELINE configs[0]); synthetic tame network seed 7",
                    "nodes": int(m.group(1)), "nps": int(m.group(2)), "unit": "nodes/s", "threads": 1}
        except Exception as e:  # pragma: no cover
            return {"unavailable": f"{type(e).__name__}: {e}"}


def engine_games(games: int = 256, soft_nodes: int = 400, plies: int = 4):
    """BASELINE configs[4] in miniature on the CPU: self-play games with the reference's own datagen search (runDatagenSearch at a soft
    node limit, src/datagen/datagen.cpp:113-121, 206-260) on its own CPU evaluation, shared out to every host core
    (oracle/_ref/sp_engine_cpu `games`).  profiles/r2_engine_on_gpu.md holds the same games played on the GPU evaluator."""
    import re
    import subprocess
    import tempfile

    from stormphrax_b200 import net as N

    engine = os.path.join(ROOT, "oracle", "_ref", "sp_engine_cpu")
    if not os.path.exists(engine):
        return {"unavailable": "oracle/_ref/sp_engine_cpu not built"}
    cores = os.cpu_count() or 1
    with tempfile.TemporaryDirectory() as tmp:
        path = os.path.join(tmp, "tame.nnue")
        N.synthetic(7, tame=True).image.tofile(path)
        try:
            out = subprocess.run([engine, path, "games", str(games), str(soft_nodes), str(plies), "42", "1", str(cores)],
                                 capture_output=True, text=True, timeout=300).stdout
            m = re.search(r"nodes: (\d+) nodes ([0-9.]+) seconds (\d+) nps checksum ([0-9a-f]+)", out)
            return {"config": f"{games} self-play games x {plies} moves, datagen search with soft limit {soft_nodes} nodes, seed 42 "
                              f"(BASELINE configs[4] in miniature); synthetic tame network seed 7",
                    "nodes": int(m.group(1)), "seconds": float(m.group(2)), "nps": int(m.group(3)), "unit": "nodes/s",
                    "threads": cores, "checksum": m.group(4)}
        except Exception as e:  # pragma: no cover
            return {"unavailable": f"{type(e).__name__}: {e}"}


# ------------------------------------------------------------------ GPU side

class Gpu:
    """What every workload of the GPU arm shares: context, stream, L2 flush buffer, timing helpers."""

    def __init__(self, local_rank: int):
        import torch

        from stormphrax_b200 import api
        from stormphrax_b200 import dist as D
        from stormphrax_b200 import net as N

        self.torch, self.api, self.D = torch, api, D
        self.ctx = api.Nnue(N.synthetic(NET_SEED).image, local_rank)
        self.props = torch.cuda.get_device_properties(local_rank)
        # a real (non-default) stream: the library treats a NULL stream as "use the context's own"
        self.stream = torch.cuda.Stream()
        torch.cuda.set_stream(self.stream)
        self.s = self.stream.cuda_stream
        assert self.s != 0
        self.ctx.set_stream(self.s)
        self.flush = torch.empty(L2_FLUSH_BYTES, dtype=torch.uint8, device="cuda")

    def barrier(self):
        self.torch.cuda.synchronize()
        self.D.barrier()
        self.torch.cuda.synchronize()

    def timed(self, step, steps: int) -> float:
        """Summed milliseconds of `steps` calls, per-step CUDA-event pairs on the launch stream; L2 is flushed
        between steps, outside the pairs; barrier + synchronize on both sides; max over ranks."""
        torch = self.torch
        evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
        self.barrier()
        for a, b in evs:
            self.flush.fill_(1)
            a.record(self.stream)
            step()
            b.record(self.stream)
        self.barrier()
        return self.D.max_over_ranks(sum(a.elapsed_time(b) for a, b in evs))


def pinned(torch, array: np.ndarray):
    return torch.from_numpy(array.view(np.uint8).reshape(-1)).pin_memory()


def bench_eval(g: Gpu, workload: str, boards, starts, steps: int, warmup: int):
    """Device-resident and end-to-end timing of one evaluation workload (full refresh or playouts).
    Returns a dict with the raw measurements and the GPU's evaluations (host array)."""
    torch, ctx, s = g.torch, g.ctx, g.s
    n = len(boards)
    h_boards = pinned(torch, boards)
    h_out = torch.empty(n, dtype=torch.int32).pin_memory()
    p_boards = boards.copy()                     # pageable host memory: what sp_nnue.h promises works
    p_out = np.empty(n, dtype=np.int32)
    d_boards = h_boards.cuda()
    d_out = torch.empty(n, dtype=torch.int32, device="cuda")
    if workload == "playouts":
        starts32 = np.ascontiguousarray(starts, dtype=np.uint32)
        h_starts = pinned(torch, starts32)
        d_starts = h_starts.cuda()
        n_games = len(starts) - 1
    L = ctx._lib

    def device_step():
        if workload == "full":
            ctx.eval_full_device(d_boards, n, d_out, s)
        else:
            ctx.eval_playouts_device(d_boards, d_starts, n_games, n, d_out, s)

    def host_step(b_ptr, o_ptr):
        if workload == "full":
            ctx._check(L.sp_nnue_eval_full(ctx._h, b_ptr, n, o_ptr))
        else:
            ctx._check(L.sp_nnue_eval_playouts(ctx._h, b_ptr, h_starts.data_ptr(), n_games, o_ptr))

    for _ in range(max(warmup, 3)):
        device_step()
    ctx.sync(s)
    for _ in range(2):
        host_step(h_boards.data_ptr(), h_out.data_ptr())
    host_step(p_boards.ctypes.data, p_out.ctypes.data)
    got = d_out.cpu().numpy()
    if not (np.array_equal(h_out.numpy(), got) and np.array_equal(p_out, got)):
        raise SystemExit("bench.py: host-pointer and device-pointer paths disagree")

    launches0 = int(ctx.counters()[g.api.CTR_LAUNCHES])
    ctx.profile(True)
    ctx.profile_read()
    total_ms = g.timed(device_step, steps)
    prof = ctx.profile_read()
    ctx.profile(False)
    launches = int(ctx.counters()[g.api.CTR_LAUNCHES]) - launches0
    e2e_ms = g.timed(lambda: host_step(h_boards.data_ptr(), h_out.data_ptr()), steps)
    pageable_ms = g.timed(lambda: host_step(p_boards.ctypes.data, p_out.ctypes.data), steps)
    h2d = n * 32 + (0 if workload == "full" else len(starts) * 4)
    return {"n": n, "ms": total_ms / steps, "e2e_ms": e2e_ms / steps, "pageable_ms": pageable_ms / steps, "prof": prof, "launches": launches,
            "h2d": h2d, "d2h": n * 4, "out": got, "api": "sp_nnue_eval_full" if workload == "full" else "sp_nnue_eval_playouts"}


def roofline_of(g: Gpu, workload: str, r: dict, boards, starts, steps: int, clock_summary: dict):
    """algorithmic bytes per launch of the dominant kernel / its measured launch time."""
    api = g.api
    n = r["n"]
    peak, peak_src = measured_peak_hbm()
    counts = api.feature_counts(boards)
    prof = r["prof"]
    stats = {"mean_rows_per_perspective": {k: v / n / 2 for k, v in counts.items()}}
    if workload == "full":
        kernel = "ft_full"
        # 2*(n_psq*2048 + n_thr*1024) + 2048 (bias) + record + eval, per position
        bytes_per_pos = (counts["psq_rows"] * 2048 + (counts["threat_rows"] + counts["pawn_pair_rows"]) * 1024) / n + 2048 + 32 + 4
    else:
        kernel = "ft_games"
        # incremental formula per perspective: delta rows + read and write of the PSQ and threat accumulators
        # (2 x 2048 B each way).  Rebuilt perspectives (first board of a game, king changed bucket or side) are
        # computed by the separate `rebuilds` kernels, which read their full row lists.
        st = api.playout_stats(boards, starts)
        stats["playout_stats_per_position"] = {k: v / n for k, v in st.items()}
        bytes_per_pos = (st["psq_delta_rows"] * 2048 + st["threat_delta_rows"] * 1024 + st["updated_perspectives"] * 2 * 4096
                         + st["rebuilt_perspectives"] * 2 * 4096) / n + 32 + 4
        rebuild_bytes_per_pos = (st["rebuild_psq_rows"] * 2048 + st["rebuild_threat_rows"] * 1024 + st["rebuilt_perspectives"] * (4096 + 32)) / n
    k_ms, k_launches = prof[kernel]
    algo_bytes_per_launch = bytes_per_pos * n * steps / max(k_launches, 1)
    avg_launch_ms = k_ms / max(k_launches, 1)
    achieved = algo_bytes_per_launch / (avg_launch_ms * 1e-3) / 1e9 if avg_launch_ms else 0.0
    roofline = {
        "bound": "hbm", "kernel": kernel, "achieved": achieved, "peak": peak, "unit": "GB/s",
        "frac": achieved / peak, "traffic": None, "peak_source": peak_src,
        "algorithmic_bytes_per_position": bytes_per_pos, "launches": k_launches, "avg_launch_ms": avg_launch_ms,
        "algorithmic_bytes_per_launch": algo_bytes_per_launch,
        # fraction of the step's wall time during which this kernel runs; the other kernels overlap it on a second
        # stream, so their event spans are stretched by sharing the SMs -- the serialised ncu launch list in
        # profiles/ gives the kernel's share without overlap
        "share_of_step": (k_ms / steps) / r["ms"] if r["ms"] else None,
        "other_kernels_ms_per_launch": {k: v[0] / max(v[1], 1) for k, v in prof.items() if k != kernel and v[1]},
        "note": ("HBM does not bind this kernel: the 89 MB network is L2-resident and rows shared by neighbouring positions are "
                 "fetched once per group, so frac > 1 is expected; see `onchip` for the resource that does bind, and "
                 "workloads.head_sweep for the one kernel that streams from HBM"),
    }
    # DRAM traffic per launch from the committed ncu capture of this kernel (profiles/traffic.json)
    entry = None
    try:
        entry = json.load(open(os.path.join(ROOT, "profiles", "traffic.json"))).get(kernel)
        roofline["traffic"] = entry["dram_bytes_per_position"] * n * steps / max(k_launches, 1)
        roofline["traffic_source"] = entry["source"]
    except Exception:
        pass
    if workload == "playouts":
        rb_ms, rb_launches = prof.get("rebuilds", (0.0, 0))
        roofline["rebuilds_kernel"] = {
            "algorithmic_bytes_per_position": rebuild_bytes_per_pos,
            "achieved": rebuild_bytes_per_pos * n * steps / (rb_ms * 1e-3) / 1e9 if rb_ms else None, "unit": "GB/s",
            "ms_per_step": rb_ms / steps,
        }
    # what binds these cache-resident kernels is on-chip: measured L2 -> SM and shared-memory bandwidth (tools/ubench.cu)
    onchip = measured_onchip()
    if onchip and workload == "full" and entry and entry.get("staged_bytes_per_position"):
        # the tensor-core full refresh stages the UNION of a group's rows once (cp.async, L2 -> shared memory): bytes from the ncu capture
        cp = onchip.get("l2_to_smem_cp_async_gbs")
        staged = entry["staged_bytes_per_position"] * n * steps / (k_ms * 1e-3) / 1e9 if k_ms else 0.0
        roofline["onchip"] = {"limit": "L2 -> shared memory staging (cp.async.cg 16 B), measured peak (tools/ubench.cu)", "peak": cp, "unit": "GB/s",
                              "achieved": staged, "frac": staged / cp if cp else None,
                              "staged_bytes_per_position": entry["staged_bytes_per_position"],
                              "issue_active_pct": entry.get("issue_active_pct"),
                              "warp_instructions_per_position": entry.get("warp_instructions_per_position"),
                              "note": "neither byte stream binds: the kernel is bound by instruction issue (ncu: issue slots 69 % busy) and the "
                                      "barriers between its phases; see profiles/r2_ft_group_ncu_v3.md and profiles/r2_ft_group_timing.md",
                              "measured": onchip}
    elif onchip:
        l2 = onchip.get("l2_read_gbs")
        roofline["onchip"] = {"limit": "L2 -> SM read bandwidth, measured (tools/ubench.cu)", "peak": l2, "unit": "GB/s",
                              "frac": achieved / l2 if l2 else None, "measured": onchip}
    else:
        sm_mhz = clock_summary.get("sm_mhz") or 1965.0
        nominal = g.props.multi_processor_count * 128 * sm_mhz * 1e6 / 1e9
        roofline["onchip"] = {"limit": "L1/LSU data path, NOMINAL 128 B/clk/SM (no measured figure committed)", "peak": nominal, "unit": "GB/s",
                              "frac": achieved / nominal}
    return roofline, stats


def check_parity(g: Gpu, rank: int, world: int, got: np.ndarray, boards, moves, starts, workload: str, budget_s: float):
    """GPU evaluations vs the CPU checker on the same positions.  Rank 0 times the CPU arm alone on its whole shard
    (that run is also the cpu_baseline); the other ranks then check a prefix of theirs.  Counts are summed over ranks."""
    cpu_rate = cpu_info = None
    checked = mismatches = 0
    if rank == 0:
        cpu_rate, cpu_info, want = cpu_arm(boards, moves, starts, workload, budget_s, complete=True)
        checked, mismatches = len(want), int((got[: len(want)] != want).sum())
    g.D.barrier()
    if rank != 0:
        _, _, want = cpu_arm(boards, moves, starts, workload, 2.0, threads=max(1, (os.cpu_count() or 1) // max(world - 1, 1)))
        checked, mismatches = len(want), int((got[: len(want)] != want).sum())
    total = g.D.allreduce_counters(np.array([checked, mismatches], dtype=np.uint64))
    return cpu_rate, cpu_info, {"checked": int(total[0]), "mismatches": int(total[1]), "checked_rank0": checked if rank == 0 else None}


# ---- the other BASELINE configs as short sub-records

def extra_playouts(g: Gpu, rank: int, world: int, steps: int, boards, moves, starts):
    r = bench_eval(g, "playouts", boards, starts, steps, 3)
    rec = {"config": workload_config("playouts")["workload"], "value": world * r["n"] / (r["ms"] * 1e-3) / 1e6, "unit": UNIT,
           "ms_per_step": r["ms"], "positions_per_gpu": r["n"], "steps": steps,
           "e2e": {"value": world * r["n"] / (r["e2e_ms"] * 1e-3) / 1e6, "unit": UNIT, "pageable": world * r["n"] / (r["pageable_ms"] * 1e-3) / 1e6,
                   "h2d_bytes_per_step": world * r["h2d"], "d2h_bytes_per_step": world * r["d2h"], "api": r["api"]},
           "gpu_launches": r["launches"],
           "kernel_ms_per_step": {k: v[0] / steps for k, v in r["prof"].items() if v[1]}}
    cpu_rate, cpu_info, parity = check_parity(g, rank, world, r["out"], boards, moves, starts, "playouts", 8.0)
    rec["parity"] = parity
    if rank == 0:
        roofline, stats = roofline_of(g, "playouts", r, boards, starts, steps, {})
        rec["roofline"] = roofline
        rec["stats"] = stats
        rec["cpu_baseline"] = {"value": cpu_rate / 1e6, "unit": UNIT, **{k: cpu_info[k] for k in ("cores", "kind", "sample", "isa")}}
    return rec, parity["mismatches"]


def extra_head_sweep(g: Gpu, boards):
    """BASELINE configs[3]: the dense head alone (sp_nnue_forward_device), M = 2^10 .. 2^20 on this GPU."""
    torch, ctx, s = g.torch, g.ctx, g.s
    n = len(boards)
    d_boards = torch.from_numpy(boards.view(np.uint8).reshape(-1)).cuda()
    d_act = torch.empty(n * 1024, dtype=torch.uint8, device="cuda")
    d_bucket = torch.empty(n, dtype=torch.uint8, device="cuda")
    d_out = torch.empty(n, dtype=torch.int32, device="cuda")
    ctx.activations_device(d_boards, n, d_act, d_bucket, s)
    ctx.sync(s)
    peak, _ = measured_peak_hbm()
    rows = []
    for logm in range(10, 21, 2):
        m = min(1 << logm, n)
        for _ in range(3):
            ctx.forward_device(d_act, d_bucket, m, d_out, s)
        times = []
        for _ in range(5):
            g.flush.fill_(0)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(g.stream)
            ctx.forward_device(d_act, d_bucket, m, d_out, s)
            e1.record(g.stream)
            torch.cuda.synchronize()
            times.append(e0.elapsed_time(e1))
        us = float(np.median(times)) * 1e3
        pos_s = m / (us * 1e-6)
        row = {"M": m, "us": us, "Mpos_s": pos_s / 1e6, "int8_TOPs": pos_s * 131072 / 1e12, "GBps": pos_s * 1029 / 1e9, "hbm_frac": pos_s * 1029 / 1e9 / peak}
        # the head kernel proper (without the counting sort in front of it): CUDA events recorded by the library around that launch
        ctx.profile(True)
        ctx.profile_read()
        for _ in range(3):
            g.flush.fill_(0)
            ctx.forward_device(d_act, d_bucket, m, d_out, s)
        ctx.sync(s)
        k_ms, k_launches = ctx.profile_read().get("head_main", (0.0, 0))
        ctx.profile(False)
        if k_launches:
            k_us = k_ms / k_launches * 1e3
            row["kernel_us"] = k_us
            row["kernel_hbm_frac"] = m * 1029 / (k_us * 1e-6) / 1e9 / peak
        rows.append(row)
    return {"config": "dense head alone: u8[M][1024] activations + bucket -> i32 (BASELINE configs[3]); `us` = the whole sp_nnue_forward_device call (counting sort + "
                      "head kernel), `kernel_us` = the head kernel alone (head_umma_kernel above 2,048 positions, head_direct_kernel up to there); L2 flushed before each call",
            "bytes_per_position": 1029, "hbm_peak_gbs": peak, "per_gpu": rows}


def extra_slots(g: Gpu, rank: int, world: int, boards, starts, steps: int):
    """The NnueState drop-in path: one device accumulator chain per game, every round = ONE sp_nnue_batch_device
    submission that advances all chains by one ply (src level -> dst level) and evaluates them."""
    torch, ctx, s = g.torch, g.ctx, g.s
    L = ctx._lib
    lens = np.diff(starts.astype(np.int64))
    games = np.flatnonzero(lens > SLOT_ROUNDS + 1)[:SLOT_STATES]
    n = len(games)
    if n == 0:
        return None, 0
    first = starts[games].astype(np.int64)
    ctx.slots_reserve(2 * n)
    ids = np.arange(n, dtype=np.uint32)
    d_level = [torch.from_numpy(2 * ids + k).cuda() for k in range(2)]
    rounds = [np.ascontiguousarray(boards[first + p]) for p in range(SLOT_ROUNDS + 1)]
    d_rounds = [torch.from_numpy(b.view(np.uint8).reshape(-1)).cuda() for b in rounds]
    d_out = torch.empty(n, dtype=torch.int32, device="cuda")
    d_ref = torch.empty(n, dtype=torch.int32, device="cuda")

    def refresh():
        ctx._check(L.sp_nnue_batch_device(ctx._h, d_level[0].data_ptr(), d_rounds[0].data_ptr(), n, None, None, None, None, 0, None, None, None, 0, None, s))

    def device_step():
        for p in range(1, SLOT_ROUNDS + 1):
            ctx._check(L.sp_nnue_batch_device(ctx._h, None, None, 0, None, d_level[(p - 1) & 1].data_ptr(), d_level[p & 1].data_ptr(),
                                              d_rounds[p].data_ptr(), n, d_out.data_ptr(), None, None, 0, None, s))

    h_src = [np.ascontiguousarray(2 * ids + k) for k in range(2)]
    h_out = np.empty(n, dtype=np.int32)

    def host_step():
        for p in range(1, SLOT_ROUNDS + 1):
            ctx._check(L.sp_nnue_batch(ctx._h, None, None, 0, None, h_src[(p - 1) & 1].ctypes.data, h_src[p & 1].ctypes.data, rounds[p].ctypes.data, n,
                                       h_out.ctypes.data, None, None, 0, None))

    refresh()
    for _ in range(2):
        device_step()
    ctx.sync(s)
    # parity: the last round's evaluations against the (reference-checked) full-refresh path on the same boards
    ctx.eval_full_device(d_rounds[SLOT_ROUNDS], n, d_ref, s)
    ctx.sync(s)
    mismatches = int((d_out != d_ref).sum().item())
    ctx.profile(True)
    ctx.profile_read()
    ms = g.timed(device_step, steps) / steps
    prof = ctx.profile_read()
    ctx.profile(False)
    host_step()
    e2e_ms = g.timed(host_step, steps) / steps
    per_s = world * n * SLOT_ROUNDS / (ms * 1e-3)
    k_ms, k_launches = prof["ft_slots"]
    # per update + activation: slot read 4 KB + slot written 4 KB + activations 1 KB + two records, from HBM
    hbm_bytes = 4096 + 4096 + 1024 + 64 + 4
    peak, _ = measured_peak_hbm()
    rec = {"config": f"{n} accumulator chains per GPU x {SLOT_ROUNDS} rounds of sp_nnue_batch_device (update src level -> dst level + evaluate)",
           "value": per_s / 1e6, "unit": "M update+eval/s", "ms_per_round": ms / SLOT_ROUNDS, "steps": steps,
           "e2e": {"value": world * n * SLOT_ROUNDS / (e2e_ms * 1e-3) / 1e6, "unit": "M update+eval/s", "api": "sp_nnue_batch (pageable host arrays)",
                   "h2d_bytes_per_round": n * 40, "d2h_bytes_per_round": n * 4},
           "parity": {"checked": n, "mismatches": mismatches, "against": "sp_nnue_eval_full_device on the same boards (itself checked against the CPU reference above)"},
           "roofline": {"bound": "hbm", "kernel": "ft_slots", "hbm_bytes_per_update": hbm_bytes,
                        "achieved": hbm_bytes * n * SLOT_ROUNDS * steps / (k_ms * 1e-3) / 1e9 if k_ms else None, "peak": peak, "unit": "GB/s",
                        "frac": hbm_bytes * n * SLOT_ROUNDS * steps / (k_ms * 1e-3) / 1e9 / peak if k_ms else None,
                        "avg_launch_ms": k_ms / max(k_launches, 1)},
           "kernel_ms_per_step": {k: v[0] / steps for k, v in prof.items() if v[1]}}
    return rec, mismatches


def run_ours(args, rank: int, world: int, local_rank: int):
    import torch

    from stormphrax_b200 import api
    from stormphrax_b200 import dist as D

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device -- the evaluator has no CPU path (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local_rank)
    os.environ.setdefault("NCCL_DEBUG", "WARN")  # NCCL's version banner goes to stdout, which carries the ONE JSON line
    D.init("nccl", local_rank)
    g = Gpu(local_rank)
    ctx = g.ctx

    boards, moves, starts = make_workload(rank, POSITIONS_PER_GPU, args.workload)
    with ClockSampler(local_rank) as clocks:
        r = bench_eval(g, args.workload, boards, starts, args.steps, args.warmup)
    clock_summary = clocks.summary()
    n = r["n"]
    counters = D.allreduce_counters(ctx.counters())  # the one collective this path has: reporting counters (NCCL)
    value = world * n / (r["ms"] * 1e-3) / 1e6
    cpu_rate, cpu_info, parity = check_parity(g, rank, world, r["out"], boards, moves, starts, args.workload, 12.0)
    bad = parity["mismatches"]

    workloads = {}
    if args.extras != "none" and args.workload == "full":
        extra_steps = max(2, min(args.steps, 5))
        pb, pm, ps = make_workload(rank, 0, "playouts")
        rec, m = extra_playouts(g, rank, world, extra_steps, pb, pm, ps)
        workloads["playouts"] = rec
        bad += m
        rec, m = extra_slots(g, rank, world, pb, ps, extra_steps)
        if rec:
            workloads["slots"] = rec
            bad += m
        del pb, pm
        rec = extra_head_sweep(g, boards)
        rec["n_gpus"] = world
        workloads["head_sweep"] = rec

    if rank == 0:
        roofline, stats = roofline_of(g, args.workload, r, boards, starts, args.steps, clock_summary)
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
            "ms_per_step": r["ms"], "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "int16/int8/int32", "data": "synthetic",
            "config": workload_config(args.workload),
            "workload_stats": {"positions_this_rank": n, **stats},
            "parity": {**parity, "against": f"{cpu_info['kind']} CPU path ({cpu_info['isa']}), every position of rank 0's shard + a prefix of the other ranks'"},
            "roofline": roofline,
            "cpu_baseline": {"value": cpu_rate / 1e6, "unit": UNIT, **{k: cpu_info[k] for k in ("cores", "kind", "sample", "isa")}},
            "e2e": {"value": world * n / (r["e2e_ms"] * 1e-3) / 1e6, "unit": UNIT, "h2d_bytes_per_step": world * r["h2d"],
                    "d2h_bytes_per_step": world * r["d2h"], "ms_per_step": r["e2e_ms"], "api": r["api"], "host_memory": "pinned",
                    "pageable": {"value": world * n / (r["pageable_ms"] * 1e-3) / 1e6, "unit": UNIT, "ms_per_step": r["pageable_ms"]}},
            "gpu_launches": r["launches"],
            "clocks": clock_summary,
            "counters_allreduced": {"evals": int(counters[api.CTR_EVALS]), "full_refresh": int(counters[api.CTR_FULL_REFRESH]),
                                    "incremental": int(counters[api.CTR_INCREMENTAL]), "launches": int(counters[api.CTR_LAUNCHES])},
        }
        if workloads:
            line["workloads"] = workloads
        print(json.dumps(line), flush=True)
    ctx.close()
    if world > 1:
        import torch.distributed as dist

        dist.barrier()
        dist.destroy_process_group()
    if bad:
        raise SystemExit(f"bench.py: {bad} evaluations differ from the CPU reference")


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", choices=["ours", "reference"], default="ours")
    ap.add_argument("--workload", choices=["full", "playouts"], default="full")
    ap.add_argument("--extras", choices=["all", "none"], default="all", help="also run the short records of the other BASELINE configs")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank, world)
        return
    if world == 1 and args.gpus > 1:
        # convenience: relaunch under torchrun, one rank per GPU
        import subprocess

        cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={args.gpus}",
               "--master-addr", "127.0.0.1", "--master-port", os.environ.get("MASTER_PORT", "29533"), os.path.abspath(__file__),
               "--gpus", str(args.gpus), "--steps", str(args.steps), "--warmup", str(args.warmup), "--workload", args.workload,
               "--extras", args.extras]
        raise SystemExit(subprocess.call(cmd))
    run_ours(args, rank, world, local_rank)


if __name__ == "__main__":
    main()
