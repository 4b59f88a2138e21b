"""Data-parallel sharding of independent image pairs: contiguous split of the pair range over ranks, weights
replicated from the shared seed (no broadcast), no collective inside the attack loop, one all-gather of the adversarial
examples + metrics afterwards.  Works with NCCL (GPU) and gloo (CPU tests)."""
from __future__ import annotations

from typing import Tuple

import torch
import torch.distributed as dist


def shard_range(n_items: int, rank: int, world: int) -> Tuple[int, int]:
    """contiguous [lo, hi) of rank; the first (n_items % world) ranks get one extra item."""
    base, rem = divmod(n_items, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def gather_results(x_local: torch.Tensor, n_total: int) -> torch.Tensor:
    """concatenate per-rank results (possibly ragged along dim 0) on every rank, in rank order."""
    if not dist.is_available() or not dist.is_initialized() or dist.get_world_size() == 1:
        return x_local
    world, rank = dist.get_world_size(), dist.get_rank()
    sizes = [shard_range(n_total, r, world) for r in range(world)]
    mx = max(hi - lo for lo, hi in sizes)
    pad = torch.zeros((mx,) + tuple(x_local.shape[1:]), dtype=x_local.dtype, device=x_local.device)
    pad[: x_local.shape[0]] = x_local
    out = [torch.empty_like(pad) for _ in range(world)]
    dist.all_gather(out, pad)
    return torch.cat([o[: hi - lo] for o, (lo, hi) in zip(out, sizes)], 0)
