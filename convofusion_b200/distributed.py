"""Data-parallel sharding of independent clips / streams.

The sampling path has no exchange step: clips (bounded generation) and streams (unbounded synthesis) are
independent units, windows of one stream are serial and stay on one rank.  One process per GPU; units are
block-partitioned over ranks, the model is replicated, and the only collective is one gather of the output
motions [n_local, T, 189] at the very end.  The reference itself is single-GPU at inference
(convofusion/config.py:92-95 forces DEVICE=[0]).
"""
from __future__ import annotations

from typing import List, Optional, Sequence, Tuple

import torch
import torch.distributed as dist


def shard_range(n_units: int, rank: int, world: int) -> Tuple[int, int]:
    """Static block partition: the first (n_units % world) ranks take one extra unit."""
    if not (0 <= rank < world):
        raise ValueError(f"rank {rank} outside world of {world}")
    base, extra = divmod(n_units, world)
    start = rank * base + min(rank, extra)
    return start, start + base + (1 if rank < extra else 0)


def batches(start: int, stop: int, batch: int) -> List[Tuple[int, int]]:
    """Split a shard into launches of at most `batch` clips (batches of 64)."""
    return [(b, min(b + batch, stop)) for b in range(start, stop, batch)]


def gather_motions(local: torch.Tensor, n_units: int, group: Optional[dist.ProcessGroup] = None) -> Optional[torch.Tensor]:
    """Gather per-rank outputs [n_local, T, F] (ragged n_local allowed) to rank 0 in unit order.
    NCCL on GPUs, gloo on CPU tensors (tests).  Returns the [n_units, T, F] tensor on rank 0, None elsewhere."""
    if not dist.is_initialized() or dist.get_world_size(group) == 1:
        return local
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    sizes = [shard_range(n_units, r, world) for r in range(world)]
    n_max = max(e - s for s, e in sizes)
    pad = local
    if local.shape[0] < n_max:   # all_gather needs equal shapes
        pad = torch.cat([local, local.new_zeros(n_max - local.shape[0], *local.shape[1:])])
    out = [torch.empty_like(pad) for _ in range(world)]
    dist.all_gather(out, pad.contiguous(), group=group)
    if rank != 0:
        return None
    return torch.cat([o[: e - s] for o, (s, e) in zip(out, sizes)])


def clip_seeds(start: int, stop: int, base_seed: int = 1234) -> List[int]:
    """Per-clip seeds `1234 + clip_id` so any sharding generates the same clips."""
    return [base_seed + i for i in range(start, stop)]
