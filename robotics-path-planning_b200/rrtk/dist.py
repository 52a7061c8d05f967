"""Multi-GPU plumbing: one process per GPU, static partition of independent queries / grid rows, no
data-path collective.  torch.distributed is used only for the barrier, the max-over-ranks timing and
the final gather of per-query summaries (a few bytes per query)."""
from __future__ import annotations

import os

from .batch import shard_range


def env_world():
    return (int(os.environ.get("RANK", "0")), int(os.environ.get("LOCAL_RANK", "0")),
            int(os.environ.get("WORLD_SIZE", "1")))


def init(backend: str = "nccl", device=None):
    """Initialise torch.distributed from the torchrun environment (no-op for world size 1)."""
    import torch.distributed as dist
    rank, local, world = env_world()
    if world > 1 and not dist.is_initialized():
        kw = {}
        if backend == "nccl" and device is not None:
            kw["device_id"] = device
        dist.init_process_group(backend, **kw)
    return rank, local, world


def my_shard(n_items: int):
    rank, _, world = env_world()
    return shard_range(n_items, rank, world)


def max_over_ranks(x: float, device=None) -> float:
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return x
    t = torch.tensor([x], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def gather_summaries(summary):
    """all_gather of a per-rank [q_local, k] tensor (equal q_local on every rank) -> [world * q_local, k]
    in global query order (rank r owns the r-th contiguous block)."""
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return summary
    parts = [torch.empty_like(summary) for _ in range(dist.get_world_size())]
    dist.all_gather(parts, summary.contiguous())
    return torch.cat(parts, 0)


def best_of_replicas(cost):
    """min-cost reduce across ranks (replicas racing on the same queries): elementwise MIN of [q] costs."""
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(cost, op=dist.ReduceOp.MIN)
    return cost
