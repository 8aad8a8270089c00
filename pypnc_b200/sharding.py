"""Batch sharding across the GPUs of one node: contiguous split of the state
dimension, no data-path collective, one final stats gather.  Backend-agnostic host logic
(NCCL on the GPU box, gloo in the CPU tests)."""
from __future__ import annotations

import numpy as np
import torch
import torch.distributed as dist


def shard_range(total, rank, world):
    """Contiguous [lo, hi) of `total` states owned by `rank`: GPU r gets [r*B/G, (r+1)*B/G)."""
    assert 0 <= rank < world
    lo = (total * rank) // world
    hi = (total * (rank + 1)) // world
    return lo, hi


def solver_stats(status, iters):
    """[count of status 0..4, sum of active-set iterations, number of states] as float64."""
    status = np.asarray(status)
    iters = np.asarray(iters)
    return np.array([float((status == s).sum()) for s in range(5)] + [float(iters.sum()), float(status.size)])


def gather_stats(stats, device=None):
    """Sum of every rank's stats vector (all ranks get it).  Uses the default process group if
    one is initialised, else returns the input."""
    t = torch.as_tensor(np.asarray(stats, dtype=np.float64), device=device)
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        parts = [torch.zeros_like(t) for _ in range(dist.get_world_size())]
        dist.all_gather(parts, t)
        t = torch.stack(parts).sum(0)
    return t.cpu().numpy()


def max_over_ranks(value, device=None):
    t = torch.tensor([float(value)], dtype=torch.float64, device=device)
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())
