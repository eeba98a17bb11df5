"""Multi-GPU policy of the per-scan path: replicas only.

Per-scan odometry of ONE sequence does not shard: scan k+1 needs the state and the map of scan k, and
IEKF iteration j+1 needs the 15x15 solve of iteration j (local_mapping.cpp:290-550). What shards
naturally is a batch of independent sequences: rank r owns sequence seed base + r, there is no
collective on the data path, and only the final throughput figures are reduced (max of the per-rank
device times, sum of the points).
"""
from __future__ import annotations

from typing import Sequence, Tuple


def sequence_seed(base_seed: int, rank: int) -> int:
    return int(base_seed) + int(rank)


def reduce_throughput(points: float, seconds: Sequence[float], group=None, device=None) -> Tuple[float, list]:
    """Whole-job figures: points summed over ranks, every time in `seconds` taken as the max over ranks."""
    import torch
    import torch.distributed as dist

    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return float(points), [float(s) for s in seconds]
    t = torch.tensor([float(s) for s in seconds], dtype=torch.float64, device=device)
    p = torch.tensor([float(points)], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX, group=group)
    dist.all_reduce(p, op=dist.ReduceOp.SUM, group=group)
    return float(p[0]), [float(x) for x in t]
