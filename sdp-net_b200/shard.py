"""Data-parallel inference plumbing: batch sharding + the two collectives that follow the path.

The forward has no cross-sample operation, so N GPUs run N independent shards
with NO collective on the hot path.  After the forward the reference's evaluation gathers what it
needs with two tiny collectives, mirrored here over `torch.distributed` (NCCL on GPUs, gloo in the
CPU tests): an all_gather of the logits and a SUM all_reduce of (correct, total)
(reference training_utilities.py:33,72-73, model_test.py:76-85).
"""
from __future__ import annotations

from typing import Optional, Tuple

import torch
import torch.distributed as dist


def shard_range(n: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous [lo, hi) slice of n samples owned by `rank` (first n % world ranks get one extra)."""
    if not (0 <= rank < world):
        raise ValueError(f"rank {rank} outside world of {world}")
    base, extra = divmod(n, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def gather_logits(local: torch.Tensor, n_total: int, group: Optional[dist.ProcessGroup] = None) -> torch.Tensor:
    """All ranks' logits in global sample order, [n_total, K]; shards may differ by one row."""
    if not dist.is_initialized() or dist.get_world_size(group) == 1:
        return local
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    rows = max(shard_range(n_total, r, world)[1] - shard_range(n_total, r, world)[0] for r in range(world))
    padded = local.new_zeros(rows, local.shape[1])
    padded[: local.shape[0]] = local
    parts = [torch.empty_like(padded) for _ in range(world)]
    dist.all_gather(parts, padded, group=group)
    out = []
    for r, p in enumerate(parts):
        lo, hi = shard_range(n_total, r, world)
        out.append(p[: hi - lo])
    del rank
    return torch.cat(out, 0)


def reduce_counts(correct: float, total: float, device=None, group: Optional[dist.ProcessGroup] = None):
    """(sum correct, sum total) over ranks -- the reference's accuracy reduction."""
    t = torch.tensor([float(correct), float(total)], dtype=torch.float64 if device is None else torch.float32,
                     device=device)
    if dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)
    return float(t[0]), float(t[1])
