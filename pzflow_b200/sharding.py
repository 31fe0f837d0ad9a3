"""Row sharding across the GPUs of one box (one process per GPU).

Every row -- and every (galaxy, grid point) pair of a posterior -- is independent
(pzflow/bijectors.py:155-160 is row-wise throughout; pzflow/flow.py:734 normalises per
galaxy), so the path shards into contiguous row blocks with NO data-path collective:
weights are replicated per rank at plan creation.  The only collectives here are the
optional result gather and the max-over-ranks used for timing.
"""
from __future__ import annotations

from typing import Callable, Tuple

import torch
import torch.distributed as dist


def world() -> Tuple[int, int]:
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(), dist.get_world_size()
    return 0, 1


def shard_rows(n: int, rank: int, world_size: int) -> Tuple[int, int]:
    """Contiguous block [start, stop) of rank ``rank``; sizes differ by at most one row."""
    base, rem = divmod(n, world_size)
    start = rank * base + min(rank, rem)
    return start, start + base + (1 if rank < rem else 0)


def sharded_eval(fn: Callable[[torch.Tensor], torch.Tensor], rows: torch.Tensor, gather: bool = False):
    """Evaluate ``fn`` on this rank's row block; optionally all-gather the [N, ...] result."""
    rank, ws = world()
    a, b = shard_rows(rows.shape[0], rank, ws)
    local = fn(rows[a:b])
    if not gather or ws == 1:
        return local
    sizes = [shard_rows(rows.shape[0], r, ws) for r in range(ws)]
    pad = max(e - s for s, e in sizes)
    buf = local.new_zeros((pad,) + tuple(local.shape[1:]))
    buf[: local.shape[0]] = local
    parts = [torch.empty_like(buf) for _ in range(ws)]
    dist.all_gather(parts, buf)
    return torch.cat([p[: e - s] for p, (s, e) in zip(parts, sizes)], dim=0)


def max_over_ranks(value: float, device=None) -> float:
    """max of a scalar over all ranks (timing is reported as the slowest rank)."""
    rank, ws = world()
    if ws == 1:
        return float(value)
    t = torch.tensor([float(value)], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())
