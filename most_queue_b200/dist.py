"""Replication sharding across the GPUs of one node.

One process per GPU (torchrun).  Replications are independent, so the data path needs no
collective: rank g simulates global replication ids [g*R/G, (g+1)*R/G) -- the Philox counter
carries the GLOBAL id, so results do not depend on the number of ranks -- and only the small
aggregate vector is summed with one allreduce (NCCL over NVLink when the process group is nccl,
gloo on CPU in the tests).
"""

from __future__ import annotations

import numpy as np


def _dist():
    try:
        import torch.distributed as dist
    except Exception:  # torch absent: single process
        return None
    if dist.is_available() and dist.is_initialized():
        return dist
    return None


def world():
    d = _dist()
    if d is None:
        return 0, 1
    return d.get_rank(), d.get_world_size()


def shard(replications: int, rank: int | None = None, world_size: int | None = None):
    """(first global replication id, count) of this rank's contiguous shard."""
    if rank is None or world_size is None:
        rank, world_size = world()
    lo = replications * rank // world_size
    hi = replications * (rank + 1) // world_size
    return lo, hi - lo


def allreduce_sum(vec: np.ndarray) -> np.ndarray:
    """Sum a small fp64 vector over all ranks (identity when not distributed)."""
    d = _dist()
    if d is None or d.get_world_size() == 1:
        return vec
    import torch

    if d.get_backend() == "nccl":
        t = torch.from_numpy(np.ascontiguousarray(vec)).cuda()
        d.all_reduce(t, op=d.ReduceOp.SUM)
        return t.cpu().numpy()
    t = torch.from_numpy(np.ascontiguousarray(vec).copy())
    d.all_reduce(t, op=d.ReduceOp.SUM)
    return t.numpy()
