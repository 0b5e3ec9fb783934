"""Embarrassingly-parallel batch sharding over GPUs.

The reference has no distributed code; every map is an independent search, so ranks own
contiguous shards of the batch and the data path needs no collective.  The only communication is
the reduction of three scalars per rank for throughput reporting (NCCL on GPUs, gloo in CPU tests).
"""
from __future__ import annotations

from typing import Optional, Tuple

import torch


def shard_range(n_items: int, rank: int, world_size: int) -> Tuple[int, int]:
    """Contiguous [begin, end) of `n_items` owned by `rank`; sizes differ by at most one and the
    union over ranks is exactly range(n_items) (ragged tails and world_size > n_items included)."""
    if world_size < 1 or not (0 <= rank < world_size):
        raise ValueError(f"bad rank/world_size {rank}/{world_size}")
    if n_items < 0:
        raise ValueError("n_items must be >= 0")
    base, extra = divmod(n_items, world_size)
    begin = rank * base + min(rank, extra)
    return begin, begin + base + (1 if rank < extra else 0)


def shard_batch(tensors, rank: int, world_size: int):
    """Slice every [B, ...] tensor to this rank's shard (views, no copies)."""
    b, e = shard_range(tensors[0].shape[0], rank, world_size)
    return tuple(t[b:e] for t in tensors)


def aggregate_throughput(maps: float, expansions: float, seconds: float, group=None,
                         device: Optional[torch.device] = None) -> Tuple[float, float, float]:
    """Whole-job (maps, expansions, seconds) = (SUM, SUM, MAX over ranks).  No-op without a process group."""
    import torch.distributed as dist

    if not (dist.is_available() and dist.is_initialized()):
        return float(maps), float(expansions), float(seconds)
    dev = device if device is not None else torch.device("cpu")
    sums = torch.tensor([maps, expansions], dtype=torch.float64, device=dev)
    mx = torch.tensor([seconds], dtype=torch.float64, device=dev)
    dist.all_reduce(sums, op=dist.ReduceOp.SUM, group=group)
    dist.all_reduce(mx, op=dist.ReduceOp.MAX, group=group)
    return float(sums[0]), float(sums[1]), float(mx[0])
