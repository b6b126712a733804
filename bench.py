#!/usr/bin/env python
"""bench.py -- throughput of the batched NNUE hot path (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload full|playouts]

A "step" is one pass of the hot path over one batch of synthetic input.  Default workload =
BASELINE.json configs[1]: full-refresh evaluation of 1,048,576 random legal positions (first
<= 80 plies of random playouts from the start position) per GPU.  One process per GPU
(torchrun sets RANK/LOCAL_RANK/WORLD_SIZE); positions shard by rank with no data-path
collective; NCCL all-reduces only the reporting counters and the max-over-ranks step time.

Prints ONE JSON line (rank 0).  `value` = device-resident throughput (CUDA events on the launch
stream), `e2e` = the same through the host-pointer C-ABI call (H2D of the packed boards and D2H
of the evals inside the timed region), `roofline` = the dominant kernel's achieved algorithmic
bytes/s against the measured HBM peak, `cpu_baseline` = the reference's CPU path on a bounded
sample of the same positions on this box's host cores.
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "Mpositions/sec batched NNUE (bit-exact vs reference CPU path)"
UNIT = "Mpos/s"
POSITIONS_PER_GPU = 1 << 20   # full refresh: BASELINE configs[1]
PLAYOUTS_PER_GPU = 1 << 16    # incremental: BASELINE configs[2], 65,536 playouts of <= 80 plies (about 5.26 M positions)
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


# ------------------------------------------------------------------ CPU side (reference arm / cpu_baseline)

def cpu_arm(boards, moves, starts, workload: str, budget_s: float, threads: int | None = None):
    """Time the reference's own CPU code (oracle/_ref) -- or the C port if it cannot run here --
    on a bounded prefix of the workload using every host core (or `threads`).  Returns (pos_per_s, info)."""
    from oracle.bind import COracle, Reference
    from stormphrax_b200 import net as N

    image = N.synthetic(NET_SEED).image
    cores = threads or os.cpu_count() or 1
    if Reference.available():
        ref = Reference()
        ref.load_net(image)
        kind, isa = "reference", f"avx{'512' if ref.isa == 'avx512' else '2'}"
    else:
        ref = COracle()
        ref.load_net(image)
        kind, isa = "port", "scalar C"
    if workload == "playouts" and kind == "reference":
        probe_games = min(len(starts) - 1, 4 * cores)
        t, _ = ref.time_playouts(boards[: starts[probe_games]], moves[: starts[probe_games]], starts[: probe_games + 1], cores, 1)
        rate = starts[probe_games] / t
        games = int(min(len(starts) - 1, max(probe_games, rate * budget_s / 81)))
        n = int(starts[games])
        t, _ = ref.time_playouts(boards[:n], moves[:n], starts[: games + 1], cores, 1)
        sample = f"first {games} playouts ({n} positions), incremental applyMove+applyImmediately+evaluate, 1 pass"
    else:
        probe = min(len(boards), 512 * cores)
        t, _ = ref.time_eval_once(boards[:probe], cores, 1)
        rate = probe / t
        n = int(min(len(boards), max(probe, rate * budget_s)))
        t, _ = ref.time_eval_once(boards[:n], cores, 1)
        sample = f"first {n} positions of the workload, evaluateOnce, 1 pass"
    return n / t, {"kind": kind, "cores": cores, "isa": isa, "sample": sample, "seconds": round(t, 3), "positions": n}


def run_reference(args, rank: int, world: int):
    if rank != 0:
        return
    boards, moves, starts = make_workload(0, POSITIONS_PER_GPU, args.workload)
    budget = min(12.0, 120.0 / max(args.steps, 1))
    rates, info = [], None
    for i in range(args.warmup + args.steps):
        rate, info = cpu_arm(boards, moves, starts, args.workload, budget if i >= args.warmup else 2.0)
        if i >= args.warmup:
            rates.append(rate)
    value = float(np.mean(rates)) / 1e6
    n_sample = info["positions"]
    # BASELINE configs[0]'s analogue: the same reference code on ONE host thread
    one_rate, one_info = cpu_arm(boards, moves, starts, args.workload, 3.0, threads=1)
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * info["seconds"], "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "int16/int8/int32", "data": "synthetic",
        "config": workload_config(args.workload, sample_positions=n_sample),
        "cpu_baseline": {"value": value, "unit": UNIT, **{k: info[k] for k in ("cores", "kind", "sample", "isa")},
                         "single_thread": {"value": one_rate / 1e6, "unit": UNIT, "sample": one_info["sample"]}},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def workload_config(workload: str, **extra):
    cfg = {
        "workload": ("full-refresh NNUE eval of 1,048,576 random legal positions per GPU (BASELINE configs[1])"
                     if workload == "full" else
                     "incremental NNUE eval along 65,536 random playouts of <= 80 plies per GPU (BASELINE configs[2])"),
        "positions_per_gpu": POSITIONS_PER_GPU if workload == "full" else None, "playouts_per_gpu": PLAYOUTS_PER_GPU if workload != "full" else None,
        "max_plies": MAX_PLIES,
        "network": f"synthetic CBNF, numpy default_rng({NET_SEED}), arch 16x704+64368 -> 1024x2 -> 32 -> 64 -> 1 x8",
        "parallelism": "positions sharded by rank, network replicated, no data-path collective",
        "l2": "L2 flushed (256 MiB write) between timed steps",
    }
    cfg.update({k: v for k, v in extra.items() if v is not None})
    return cfg


# ------------------------------------------------------------------ GPU side

def run_ours(args, rank: int, world: int, local_rank: int):
    import torch

    from stormphrax_b200 import api
    from stormphrax_b200 import dist as D
    from stormphrax_b200 import net as N

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device -- the evaluator has no CPU path (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local_rank)
    D.init("nccl", local_rank)

    boards, moves, starts = make_workload(rank, POSITIONS_PER_GPU, args.workload)
    n = len(boards)
    ctx = api.Nnue(N.synthetic(NET_SEED).image, local_rank)
    props = torch.cuda.get_device_properties(local_rank)
    # a real (non-default) stream: the library treats a NULL stream as "use the context's own"
    stream = torch.cuda.Stream()
    torch.cuda.set_stream(stream)
    s = stream.cuda_stream
    assert s != 0
    ctx.set_stream(s)

    h_boards = torch.from_numpy(boards.view(np.uint8).reshape(-1)).pin_memory()
    h_out = torch.empty(n, dtype=torch.int32).pin_memory()
    d_boards = h_boards.cuda()
    d_out = torch.empty(n, dtype=torch.int32, device="cuda")
    flush = torch.empty(L2_FLUSH_BYTES, dtype=torch.uint8, device="cuda")
    if args.workload == "playouts":
        h_starts = torch.from_numpy(starts.astype(np.uint32).view(np.int32)).pin_memory()
        d_starts = h_starts.cuda()
        n_games = len(starts) - 1

    def device_step():
        if args.workload == "full":
            ctx.eval_full_device(d_boards, n, d_out, s)
        else:
            ctx.eval_playouts_device(d_boards, d_starts, n_games, n, d_out, s)

    def host_step():
        if args.workload == "full":
            ctx._check(ctx._lib.sp_nnue_eval_full(ctx._h, h_boards.data_ptr(), n, h_out.data_ptr()))
        else:
            ctx._check(ctx._lib.sp_nnue_eval_playouts(ctx._h, h_boards.data_ptr(), h_starts.data_ptr(), n_games, h_out.data_ptr()))

    def barrier():
        torch.cuda.synchronize()
        D.barrier()
        torch.cuda.synchronize()

    def timed(step, steps):
        """Per-step CUDA-event pairs on the launch stream; L2 flushed between steps, outside the pairs."""
        evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
        barrier()
        for a, b in evs:
            flush.fill_(1)
            a.record(stream)
            step()
            b.record(stream)
        barrier()
        return D.max_over_ranks(sum(a.elapsed_time(b) for a, b in evs))

    # warm-up (also verifies results once against the CPU checker on rank 0)
    for _ in range(max(args.warmup, 3)):
        device_step()
    ctx.sync(s)
    for _ in range(2):
        host_step()
    if not np.array_equal(h_out.numpy(), d_out.cpu().numpy()):
        raise SystemExit("bench.py: host-pointer and device-pointer paths disagree")

    launches0 = int(ctx.counters()[api.CTR_LAUNCHES])
    ctx.profile(True)
    ctx.profile_read()
    with ClockSampler(local_rank) as clocks:
        total_ms = timed(device_step, args.steps)
        prof = ctx.profile_read()
        ctx.profile(False)
        launches = int(ctx.counters()[api.CTR_LAUNCHES]) - launches0
        e2e_ms = timed(host_step, args.steps)
    clock_summary = clocks.summary()

    counters = D.allreduce_counters(ctx.counters())  # the one collective this path has: reporting counters (NCCL)

    ms_per_step = total_ms / args.steps
    value = world * n / (ms_per_step * 1e-3) / 1e6
    e2e_value = world * n / (e2e_ms / args.steps * 1e-3) / 1e6

    if rank == 0:
        # roofline of the dominant kernel: algorithmic bytes per launch / measured launch time
        peak, peak_src = measured_peak_hbm()
        counts = api.feature_counts(boards)
        kernel = "ft_full" if args.workload == "full" else "ft_games"
        k_ms, k_launches = prof[kernel]
        if args.workload == "full":
            # 2*(n_psq*2048 + n_thr*1024) + 2048 (bias) + record + eval, per position
            bytes_per_pos = (counts["psq_rows"] * 2048 + (counts["threat_rows"] + counts["pawn_pair_rows"]) * 1024) / n + 2048 + 32 + 4
        else:
            # SURVEY 8(d) incremental formula per perspective: delta rows + read and write of the PSQ and
            # threat accumulators (2 x 2048 B each way).  Rebuilt perspectives (first board of a game, king
            # changed bucket or side) are computed by the separate `rebuilds` kernels, which read their full
            # row lists and write the accumulator the walker then loads.
            st = api.playout_stats(boards, starts)
            bytes_per_pos = (st["psq_delta_rows"] * 2048 + st["threat_delta_rows"] * 1024 + st["updated_perspectives"] * 2 * 4096
                             + st["rebuilt_perspectives"] * 2 * 4096) / n + 32 + 4
            rebuild_bytes_per_pos = (st["rebuild_psq_rows"] * 2048 + st["rebuild_threat_rows"] * 1024 + st["rebuilt_perspectives"] * (4096 + 32)) / n
        algo_bytes_per_launch = bytes_per_pos * n * args.steps / max(k_launches, 1)
        avg_launch_ms = k_ms / max(k_launches, 1)
        achieved = algo_bytes_per_launch / (avg_launch_ms * 1e-3) / 1e9 if avg_launch_ms else 0.0
        other_ms = sum(v[0] for k, v in prof.items() if k != kernel)
        roofline = {
            "bound": "hbm", "kernel": kernel, "achieved": achieved, "peak": peak, "unit": "GB/s",
            "frac": achieved / peak, "traffic": None, "peak_source": peak_src,
            "algorithmic_bytes_per_position": bytes_per_pos, "launches": k_launches, "avg_launch_ms": avg_launch_ms,
            # fraction of the step's wall time during which this kernel runs; the other kernels overlap it on
            # a second stream, so their event spans (below) are stretched by sharing the SMs -- the serialised
            # ncu launch list in profiles/ gives the kernel's share without overlap
            "share_of_step": (k_ms / args.steps) / ms_per_step if ms_per_step else None,
            "other_kernels_ms_per_launch": {k: v[0] / max(v[1], 1) for k, v in prof.items() if k != kernel and v[1]},
        }
        # DRAM traffic per launch from the committed ncu capture of this kernel (profiles/traffic.json)
        traffic_path = os.path.join(ROOT, "profiles", "traffic.json")
        if os.path.exists(traffic_path):
            try:
                entry = json.load(open(traffic_path)).get(kernel)
                roofline["traffic"] = entry["dram_bytes_per_position"] * n * args.steps / max(k_launches, 1)
                roofline["traffic_source"] = entry["source"]
            except Exception:
                pass
        roofline["algorithmic_bytes_per_launch"] = algo_bytes_per_launch
        if args.workload == "playouts":
            rb_ms, rb_launches = prof.get("rebuilds", (0.0, 0))
            roofline["rebuilds_kernel"] = {
                "algorithmic_bytes_per_position": rebuild_bytes_per_pos,
                "achieved": rebuild_bytes_per_pos * n * args.steps / (rb_ms * 1e-3) / 1e9 if rb_ms else None, "unit": "GB/s",
                "ms_per_step": rb_ms / args.steps,
            }
        # the resource that actually binds these cache-resident kernels: the SM's 128 B/clk L1/LSU data path
        sm_mhz = clock_summary.get("sm_mhz") or 1965.0
        onchip_peak = props.multi_processor_count * 128 * sm_mhz * 1e6 / 1e9
        roofline["onchip"] = {"limit": "L1/LSU data path, 128 B/clk/SM", "peak": onchip_peak, "unit": "GB/s", "frac": achieved / onchip_peak}
        cpu_rate, cpu_info = cpu_arm(boards, moves, starts, args.workload, 12.0)
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
            "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "int16/int8/int32", "data": "synthetic",
            "config": workload_config(
                args.workload,
                positions_this_rank=n,
                mean_rows_per_perspective={k: v / n / 2 for k, v in counts.items()},
                playout_stats_per_position=({k: v / n for k, v in st.items()} if args.workload == "playouts" else None),
            ),
            "roofline": roofline,
            "cpu_baseline": {"value": cpu_rate / 1e6, "unit": UNIT, **{k: cpu_info[k] for k in ("cores", "kind", "sample", "isa")}},
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": world * n * 32 + (0 if args.workload == "full" else world * (len(starts)) * 4),
                    "d2h_bytes_per_step": world * n * 4, "ms_per_step": e2e_ms / args.steps,
                    "api": "sp_nnue_eval_full" if args.workload == "full" else "sp_nnue_eval_playouts"},
            "gpu_launches": launches,
            "clocks": clock_summary,
            "counters_allreduced": {"evals": int(counters[api.CTR_EVALS]), "full_refresh": int(counters[api.CTR_FULL_REFRESH]),
                                    "incremental": int(counters[api.CTR_INCREMENTAL]), "launches": int(counters[api.CTR_LAUNCHES])},
        }
        print(json.dumps(line), flush=True)
    ctx.close()
    if world > 1:
        import torch.distributed as dist

        dist.barrier()
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", choices=["ours", "reference"], default="ours")
    ap.add_argument("--workload", choices=["full", "playouts"], default="full")
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
               "--gpus", str(args.gpus), "--steps", str(args.steps), "--warmup", str(args.warmup), "--workload", args.workload]
        raise SystemExit(subprocess.call(cmd))
    run_ours(args, rank, world, local_rank)


if __name__ == "__main__":
    main()
