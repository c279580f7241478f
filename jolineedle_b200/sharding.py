"""Multi-GPU host logic: one process per GPU, episodes sharded by rank, no data-path collective.

Episodes are independent (each owns its image, bitmap and position; no cross-episode term in
the reward or the returns -- general_env.py:321-358, reinforce.py:196-202), so the env path
needs no exchange step.  What does cross ranks is bookkeeping only: the max-over-ranks step
time of the benchmark and the optional sum-reduction of eval metrics (the reference evaluates
on rank 0 only, supervised.py:904-909).  Both are tiny ``torch.distributed`` all-reduces (NCCL
on GPUs, gloo in the CPU tests).
"""
import os
from typing import Dict, Tuple

import torch
import torch.distributed as dist


def dist_env() -> Tuple[int, int, int]:
    """(rank, local_rank, world_size) from the torchrun environment (1-process defaults)."""
    return (int(os.environ.get("RANK", 0)), int(os.environ.get("LOCAL_RANK", 0)),
            int(os.environ.get("WORLD_SIZE", 1)))


def shard_bounds(n_items: int, rank: int, world_size: int) -> Tuple[int, int]:
    """Contiguous, balanced partition of ``n_items`` episodes: the first ``n_items % world_size``
    ranks take one extra.  Same coverage as the reference's DistributedSampler
    (reinforce.py:284-294) without its padding duplicates."""
    assert 0 <= rank < world_size
    base, extra = divmod(n_items, world_size)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def max_over_ranks(value: float, device) -> float:
    """Slowest rank's value (step time): the only number a multi-GPU bench may report."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return float(value)
    t = torch.tensor([value], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def sum_over_ranks(value: float, device) -> float:
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return float(value)
    t = torch.tensor([value], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return float(t.item())


def reduce_eval_metrics(sums: Dict[str, float], count: int, device) -> Dict[str, float]:
    """Mean of per-episode eval metrics over all ranks: one all-reduce(SUM) of [sums..., count]."""
    keys = sorted(sums)
    t = torch.tensor([sums[k] for k in keys] + [float(count)], dtype=torch.float64, device=device)
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
    total = max(float(t[-1].item()), 1.0)
    return {k: float(t[i].item()) / total for i, k in enumerate(keys)}
