"""Host-side multi-GPU logic on CPU: two processes, gloo backend.  Covers the replication sharding
and the aggregate allreduce used by QsSim / PriorityQueueSimulator / NetworkSimulator, and the
summary all-gather + carry fold of the Lindley chain (no kernels are launched here)."""

import os
import socket
import sys

import numpy as np
import pytest
import torch.distributed as tdist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    tdist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from most_queue_b200 import dist
        from most_queue_b200.sim import lindley

        assert dist.world() == (rank, world)
        lo, n = dist.shard(1001)
        # every replication id is owned by exactly one rank
        owned = np.zeros(1001)
        owned[lo:lo + n] = 1
        total = dist.allreduce_sum(owned)
        assert np.array_equal(total, np.ones(1001))
        # aggregates add up
        agg = np.arange(10, dtype=np.float64) * (rank + 1)
        s = dist.allreduce_sum(agg)
        assert np.array_equal(s, np.arange(10) * sum(range(1, world + 1)))
        # Lindley: each rank publishes (a, b); carries are folded in rank order
        pairs = lindley._gather_pairs((float(rank), -1.5 * (rank + 1)))
        assert pairs == [(float(r), -1.5 * (r + 1)) for r in range(world)]
        carries, w_out = lindley.fold_carries(pairs, 0.0)
        q.put((rank, lo, n, carries, w_out))
    finally:
        tdist.destroy_process_group()


def test_sharding_allreduce_and_carry_fold_two_ranks():
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    out = sorted(q.get(timeout=120) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert [(o[1], o[2]) for o in out] == [(0, 500), (500, 501)]
    # carries: range 0 starts from 0; range 1 from max(a0, 0 + b0) = max(0, -1.5) = 0
    assert out[0][3] == out[1][3] == [0.0, 0.0]
    assert out[0][4] == max(1.0, 0.0 - 3.0)


def test_shard_is_a_partition_for_any_world_size():
    from most_queue_b200 import dist

    for world in (1, 2, 3, 4, 8):
        for R in (1, 7, 8, 1_000_000):
            cover = []
            for r in range(world):
                lo, n = dist.shard(R, r, world)
                cover.extend(range(lo, lo + n) if R < 100 else [(lo, n)])
            if R < 100:
                assert cover == list(range(R))
            else:
                assert sum(n for _, n in cover) == R and cover[0][0] == 0
                assert all(cover[i][0] + cover[i][1] == cover[i + 1][0] for i in range(world - 1))


def test_fold_carries_matches_sequential_lindley():
    from most_queue_b200.sim.lindley import fold_carries
    from oracle import oracle

    rng = np.random.default_rng(3)
    N = 5000
    A = rng.exponential(1.0, N + 1)
    S = rng.exponential(0.9, N)
    bounds = [0, 1200, 2500, 4100, N]
    pairs = []
    for g in range(4):
        _, o = oracle.lindley(A[bounds[g]:], S[bounds[g]:], bounds[g + 1] - bounds[g], tree=False)
        pairs.append((o.a, o.b))
    carries, _ = fold_carries(pairs)
    W, _ = oracle.lindley(A, S, N, tree=False)
    for g in range(1, 4):
        assert abs(carries[g] - W[bounds[g]]) < 1e-9
