"""Read-id sharding across ranks (one process per GPU) — SURVEY.md §8e.

Every read's term is independent given the walks, so each rank scores the contiguous read-id block it
holds and the only exchange is the per-set partial {sum_hi, sum_lo, floored}: 24 bytes per read set per
evaluation. The partials are ALL-GATHERED (not sum-reduced) and every rank adds them in rank order with
an error-free transformation, so the rounded total does not depend on the collective's internal order;
`torch.distributed` (NCCL over NVLink on the GPU box, gloo in CPU tests) is only the plumbing.
"""
from __future__ import annotations

from typing import Sequence, Tuple

import numpy as np


def shard_bounds(n_reads: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous read-id block of `rank` (pairs stay together because a pair shares one read id)."""
    per = (n_reads + world - 1) // world
    return min(rank * per, n_reads), min((rank + 1) * per, n_reads)


def allgather_partials(partials: np.ndarray, device=None) -> np.ndarray:
    """[n_sets*3] float64 of this rank -> [world, n_sets*3] on every rank."""
    import torch
    import torch.distributed as dist
    if not dist.is_initialized() or dist.get_world_size() == 1:
        return np.asarray(partials, dtype=np.float64)[None, :]
    t = torch.from_numpy(np.ascontiguousarray(partials, dtype=np.float64))
    if device is not None:
        t = t.to(device)
    out = [torch.empty_like(t) for _ in range(dist.get_world_size())]
    dist.all_gather(out, t)
    return torch.stack(out).cpu().numpy()


def sharded_calc_prob(pc, paths: Sequence[Sequence[int]], device=None):
    """CalcProb over all ranks' shards: local partials -> all-gather -> ordered combine (same value on every rank)."""
    part, tl = pc.calc_prob_partial(paths)
    g = allgather_partials(part, device)
    return pc.combine(g, g.shape[0], tl)
