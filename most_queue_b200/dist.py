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


_dev_buf = {}  # (device index, length) -> (device tensor, pinned host tensor): allocated once, reused by every call


def _buffers(n: int):
    import torch

    key = (torch.cuda.current_device(), n)
    if key not in _dev_buf:
        _dev_buf[key] = (torch.empty(n, dtype=torch.float64, device="cuda"),
                         torch.empty(n, dtype=torch.float64).pin_memory())
    return _dev_buf[key]


def allreduce_sum(vec: np.ndarray) -> np.ndarray:
    """Sum a small fp64 vector over all ranks (identity when not distributed).  With NCCL the vector goes through a
    persistent device buffer and a pinned host buffer (no allocation, no pageable copy per call)."""
    d = _dist()
    if d is None or d.get_world_size() == 1:
        return vec
    import torch

    if d.get_backend() == "nccl":
        dev, host = _buffers(len(vec))
        host.numpy()[:] = vec
        dev.copy_(host, non_blocking=True)
        d.all_reduce(dev, op=d.ReduceOp.SUM)
        host.copy_(dev)
        return host.numpy().copy()
    t = torch.from_numpy(np.ascontiguousarray(vec).copy())
    d.all_reduce(t, op=d.ReduceOp.SUM)
    return t.numpy()


def common_seed(seed) -> int:
    """The Philox seed of a run.  An explicit seed is used as is (every rank was given the same one).  seed=None means
    non-deterministic like numpy.random.default_rng(None) (base_core.py:23): rank 0 draws one from os.urandom and
    broadcasts it, because a run sharded over ranks must use ONE key -- otherwise the result would depend on the
    number of ranks, and the ranges of one chain would be generated from different streams."""
    if seed is not None:
        return int(seed)
    import os

    seed = int.from_bytes(os.urandom(8), "little") & 0x7FFFFFFFFFFFFFFF
    d = _dist()
    if d is None or d.get_world_size() == 1:
        return seed
    import torch

    t = torch.tensor([seed], dtype=torch.int64)
    if d.get_backend() == "nccl":
        t = t.cuda()
    d.broadcast(t, src=0)
    return int(t.item())


def local_devices(devices=None):
    """Devices a single-process run spreads over: an explicit list, "all", or the MQ_DEVICES environment variable
    ("all" or "0,1,2"); None (the default) means one GPU, as before.  Ignored under torch.distributed, where every
    rank owns one GPU."""
    import os

    if _dist() is not None and _dist().get_world_size() > 1:
        return None
    if devices is None:
        devices = os.environ.get("MQ_DEVICES")
        if not devices:
            return None
        devices = "all" if devices == "all" else [int(x) for x in devices.split(",")]
    return devices


def run_on_devices(contexts, replications: int, fn):
    """Single-process multi-GPU: fn(ctx, rep0, count) -> aggregate vector runs on one host thread per context (the
    library call releases the GIL) on a contiguous share of the replications; the aggregates -- sums over
    replications -- are added in device order."""
    import threading

    n = len(contexts)
    out, err = [None] * n, [None] * n

    def work(i):
        lo, hi = replications * i // n, replications * (i + 1) // n
        try:
            out[i] = fn(contexts[i], lo, hi - lo) if hi > lo else None
        except BaseException as e:  # re-raised in the caller's thread
            err[i] = e

    th = [threading.Thread(target=work, args=(i,)) for i in range(n)]
    for t in th:
        t.start()
    for t in th:
        t.join()
    for e in err:
        if e is not None:
            raise e
    parts = [x for x in out if x is not None]
    total = parts[0].copy()
    for x in parts[1:]:
        total += x
    return total
