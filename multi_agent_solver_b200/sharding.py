"""Host-side plumbing for one-process-per-GPU runs: how a batch of independent OCPs (or multi-agent
scenarios) is split over ranks and how per-rank results are put back in problem order.

The hot path itself has no cross-problem coupling (every OCP solve is independent, and in
the Nash strategies agents never read each other's trajectories during a solve), so sharding is a
contiguous range per rank and the only collective is a gather of results.  These helpers work with
any torch.distributed backend (nccl on the GPU box, gloo in the CPU tests).
"""
from __future__ import annotations

from typing import List, Tuple

import numpy as np


def shard_bounds(total: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous range [lo, hi) of `total` units owned by `rank`; the remainder goes to the first ranks."""
    if world <= 0 or not (0 <= rank < world):
        raise ValueError("bad rank/world")
    base, rem = divmod(total, world)
    lo = rank * base + min(rank, rem)
    hi = lo + base + (1 if rank < rem else 0)
    return lo, hi


def shard_counts(total: int, world: int) -> List[int]:
    return [shard_bounds(total, r, world)[1] - shard_bounds(total, r, world)[0] for r in range(world)]


def allgather_rows(local: np.ndarray, total: int, dist=None) -> np.ndarray:
    """All ranks pass their shard (rows = units, contiguous ranges from shard_bounds) and get the full
    array in unit order.  Shards may differ in length by one row; they are padded for the collective."""
    import torch

    if dist is None:
        import torch.distributed as dist  # type: ignore
    world = dist.get_world_size()
    counts = shard_counts(total, world)
    width = max(counts)
    pad = np.zeros((width,) + local.shape[1:], dtype=local.dtype)
    pad[: local.shape[0]] = local
    mine = torch.from_numpy(pad)
    parts = [torch.empty_like(mine) for _ in range(world)]
    dist.all_gather(parts, mine)
    return np.concatenate([parts[r].numpy()[: counts[r]] for r in range(world)], axis=0)


def ordered_total(costs: np.ndarray) -> float:
    """Sum in block order starting from 0.0, as collect_solution does (strategies/nash.hpp:28-35); an
    all-reduce would not fix the summation order, so totals are formed after the gather."""
    tot = 0.0
    for c in np.asarray(costs, dtype=np.float64).reshape(-1):
        tot += float(c)
    return tot
