"""Multi-GPU plumbing: one process per GPU, positions sharded by rank, network replicated.

The evaluation path has no exchange step: the only collectives are the
reporting counters (what the engine sums over threads at report time, src/thread.h:39-68) and the
max-over-ranks step time.  Backend "nccl" on GPUs, "gloo" in the CPU tests.
"""
from __future__ import annotations

import os

import numpy as np


def env_rank() -> tuple[int, int, int]:
    """(rank, world_size, local_rank) from the torchrun environment; (0, 1, 0) when absent."""
    return int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("LOCAL_RANK", "0"))


def init(backend: str | None = None, device_index: int | None = None) -> bool:
    """Join the process group if this is a multi-rank launch. Returns True when a group exists."""
    import torch
    import torch.distributed as dist

    _, world, _ = env_rank()
    if world <= 1:
        return False
    if dist.is_initialized():
        return True
    if backend is None:
        backend = "nccl" if torch.cuda.is_available() else "gloo"
    kwargs = {}
    if backend == "nccl" and device_index is not None:
        kwargs["device_id"] = torch.device("cuda", device_index)
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    dist.init_process_group(backend, **kwargs)
    return True


def shard_range(n: int, rank: int, world: int) -> tuple[int, int]:
    """Contiguous index range [lo, hi) of rank `rank` ([g N / G, (g + 1) N / G))."""
    return n * rank // world, n * (rank + 1) // world


def shard_games(game_start: np.ndarray, rank: int, world: int) -> tuple[int, int]:
    """Whole games per rank: games [lo, hi) such that ranks get near-equal position counts."""
    n_games = len(game_start) - 1
    total = int(game_start[-1])
    cuts = [int(np.searchsorted(game_start, total * r // world, side="left")) for r in range(world + 1)]
    cuts[0], cuts[-1] = 0, n_games
    return cuts[rank], cuts[rank + 1]


def _device():
    import torch
    import torch.distributed as dist

    return torch.device("cuda", torch.cuda.current_device()) if dist.get_backend() == "nccl" else torch.device("cpu")


def allreduce_counters(counters: np.ndarray) -> np.ndarray:
    """Sum the per-context uint64 counters (sp_nnue_counters) over all ranks."""
    import torch
    import torch.distributed as dist

    if not (dist.is_available() and dist.is_initialized()):
        return counters.copy()
    t = torch.from_numpy(counters.astype(np.int64)).to(_device())
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return t.cpu().numpy().astype(np.uint64)


def max_over_ranks(value: float) -> float:
    import torch
    import torch.distributed as dist

    if not (dist.is_available() and dist.is_initialized()):
        return value
    t = torch.tensor([value], dtype=torch.float64, device=_device())
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def barrier() -> None:
    import torch.distributed as dist

    if dist.is_available() and dist.is_initialized():
        dist.barrier()
