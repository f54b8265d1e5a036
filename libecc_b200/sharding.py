"""Multi-GPU plumbing: the batch is embarrassingly parallel, so each rank takes a contiguous shard
and the only exchange is one all-gather of the fixed-size results (NCCL over NVLink on GPUs; gloo in the CPU tests).
No arithmetic here."""
from __future__ import annotations

from typing import Tuple

import torch
import torch.distributed as dist


def shard_bounds(n: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous index range [lo, hi) of rank `rank` among `world` ranks; sizes differ by at most one."""
    return (n * rank) // world, (n * (rank + 1)) // world


def gather_results(local: torch.Tensor, n_total: int, item: int, group=None) -> torch.Tensor:
    """All-gather per-rank result records (`item` bytes each, contiguous shards from shard_bounds) into the full
    [n_total * item] tensor on every rank.  Even shards use one all_gather_into_tensor; ragged ones are padded to the
    largest shard and trimmed."""
    world = dist.get_world_size(group)
    sizes = [shard_bounds(n_total, r, world) for r in range(world)]
    counts = [hi - lo for lo, hi in sizes]
    assert local.numel() == counts[dist.get_rank(group)] * item
    if len(set(counts)) == 1:
        full = torch.empty(n_total * item, dtype=local.dtype, device=local.device)
        dist.all_gather_into_tensor(full, local.contiguous().view(-1), group=group)
        return full
    mx = max(counts) * item
    padded = torch.zeros(mx, dtype=local.dtype, device=local.device)
    padded[: local.numel()] = local.view(-1)
    buf = torch.empty(world * mx, dtype=local.dtype, device=local.device)
    dist.all_gather_into_tensor(buf, padded, group=group)
    return torch.cat([buf[r * mx: r * mx + counts[r] * item] for r in range(world)])
