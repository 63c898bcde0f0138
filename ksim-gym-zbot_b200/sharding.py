"""Multi-GPU plumbing: environments shard trivially — one process per GPU, each rank owns a contiguous
block of environments, NO collective inside the physics step.  The only cross-rank traffic is the throughput/metric
reduction of the harness (and, in the full trainer, the PPO gradient all-reduce, which is outside this path)."""

from __future__ import annotations

import torch
import torch.distributed as dist


def shard_range(n_total: int, world: int, rank: int) -> tuple[int, int]:
    """Contiguous [lo, hi) block of environments owned by ``rank`` (remainder spread over the first ranks)."""
    if not (0 <= rank < world):
        raise ValueError("rank out of range")
    base, rem = divmod(n_total, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def rank_seed(seed: int, rank: int) -> int:
    """Independent random streams per rank (actions, pushes, resets)."""
    return seed * 1_000_003 + rank


def aggregate_throughput(units_this_rank: float, elapsed_ms_this_rank: float, device=None) -> tuple[float, float]:
    """Whole-job throughput: sum of the units all ranks processed / MAX over ranks of the device time.
    Returns (units_per_second, max_elapsed_ms).  Works without an initialised process group (single rank)."""
    t = torch.tensor([elapsed_ms_this_rank], dtype=torch.float64, device=device)
    u = torch.tensor([units_this_rank], dtype=torch.float64, device=device)
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dist.all_reduce(u, op=dist.ReduceOp.SUM)
    tmax = float(t.item())
    return float(u.item()) / (tmax * 1e-3), tmax
