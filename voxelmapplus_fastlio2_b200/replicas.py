"""Multi-GPU = replicas only: a scan does not shard (every point may touch any voxel, LRU / merge
order is global, the per-iteration reduction is 27 doubles), so N GPUs run N independent trajectories, one map
per GPU, with NO collective on the data path.  The only cross-rank traffic is the measurement protocol:
a barrier around the timed region and a MAX over ranks of the device time."""
from __future__ import annotations

import os

BASE_SEED = 0xC0FFEE


def rank_env():
    return (int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("LOCAL_RANK", "0")))


def rank_seed(rank: int, base: int = BASE_SEED) -> int:
    """sequence seed of replica `rank` (seeds 0xC0FFEE + 0..7)"""
    return base + rank


def max_over_ranks(value: float, device=None) -> float:
    """MAX over ranks of a per-rank scalar (device time of the timed region)."""
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return float(value)
    t = torch.tensor([float(value)], dtype=torch.float64, device=device if device is not None else "cpu")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def whole_job_throughput(steps_per_rank: int, local_ms: float, device=None):
    """(scans/s over all replicas, ms per step): units all ranks processed / max-over-ranks time."""
    import torch.distributed as dist
    world = dist.get_world_size() if (dist.is_available() and dist.is_initialized()) else 1
    tmax = max_over_ranks(local_ms, device)
    return world * steps_per_rank / (tmax * 1e-3), tmax / steps_per_rank
