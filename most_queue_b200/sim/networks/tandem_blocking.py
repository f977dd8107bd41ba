"""Tandem line with finite buffers and blocking after service: drop-in for the reference's
TandemBlockingSim (most_queue/sim/networks/tandem_blocking.py).

Same interface: TandemBlockingSim(seed).set_sources(arrival_rate, kendall) (:32),
set_nodes(serv_params, capacity) (:38), run(total_served, warmup_fraction=0.05) -> NetworkMeansResults(v,
mean_jobs, served) (:50-142); `throughput` and `loss_prob` are attributes after run().  The event loop runs
on the GPU (csrc/k6_tandem.cu).  Additive: `replications=`.
"""

from __future__ import annotations

import os
import time

import numpy as np

from ... import _lib, dist
from ...structs import NetworkMeansResults

Z99 = 2.5758293035489004


class TandemBlockingSim:
    def __init__(self, seed: int | None = None, replications: int = 1, device=None):
        self.seed = seed
        self.replications = int(replications)
        self.device = device
        self.arrival_rate = None
        self.capacity = None
        self.m = 0
        self._src = self._srv = None
        self.is_sources_set = self.is_nodes_set = False
        self.loss_prob = None
        self.throughput = None
        self.ci = self.agg = self.per_replication = self.timing = None

    def set_sources(self, arrival_rate: float, kendall: str = "M"):
        self.arrival_rate = arrival_rate
        self._src = _lib.make_dist(kendall, arrival_rate)
        self.is_sources_set = True

    def set_nodes(self, serv_params: list[dict], capacity: list):
        self._srv = [_lib.make_dist(p["type"], p["params"]) for p in serv_params]
        self.capacity = [None if (k is None or k == float("inf")) else int(k) for k in capacity]
        self.m = len(self.capacity)
        self.is_nodes_set = True

    def run(self, total_served: int, warmup_fraction: float = 0.05, replications: int | None = None,
            keep_replications: bool = False) -> NetworkMeansResults:
        start = time.process_time()
        if not (self.is_sources_set and self.is_nodes_set):
            raise RuntimeError("set sources and nodes before run()")
        R = int(self.replications if replications is None else replications)
        seed = dist.common_seed(self.seed)
        rep0, count = dist.shard(R)
        ctx = _lib.default_context(self.device)
        m = self.m
        if count > 0:
            agg, rec = ctx.sim_tandem(self.capacity, self._src, self._srv, total_served, warmup_fraction, count, rep0,
                                      seed, per_replication=keep_replications)
            self.timing = ctx.last_timing()
        else:
            agg, rec = np.zeros(_lib.tandem_agg_len(m)), None
        agg = dist.allreduce_sum(agg)
        self.agg, self.per_replication = agg, rec
        reps = agg[0]
        if agg[1] > 0:
            raise _lib.MqError(_lib.MQ_E_OVERFLOW, f"{int(agg[1])} replications failed")

        def mean_ci(s, s2):
            mean = s / reps
            if reps > 1:
                var = max(s2 / reps - mean * mean, 0.0) * reps / (reps - 1)
                return float(mean), float(Z99 * np.sqrt(var / reps))
            return float(mean), float("nan")

        self.throughput, thr_ci = mean_ci(agg[2], agg[3])
        self.loss_prob, loss_ci = mean_ci(agg[4], agg[5])
        v0, v_ci = mean_ci(agg[6], agg[7])
        mean_jobs, mj_ci = [], []
        for i in range(m):
            a, b = mean_ci(agg[8 + 2 * i], agg[9 + 2 * i])
            mean_jobs.append(a)
            mj_ci.append(b)
        self.ci = {"throughput": thr_ci, "loss_prob": loss_ci, "v": [v_ci], "mean_jobs": mj_ci, "level": 0.99}
        warm = int(total_served * warmup_fraction)
        return NetworkMeansResults(v=[v0], mean_jobs=mean_jobs, served=int(total_served) + warm,
                                   duration=time.process_time() - start, replications=R, ci=self.ci)


def replay(capacity, A, S, total_served, warmup_fraction=0.05, device=None):
    """Replay tier: external gaps A[i, r] and per-node service draws S[node][j, r]; per-replication results
    formed with the reference's own final operations (tandem_blocking.py:130-137), bit-exact in fp64."""
    L = _lib
    ctx = L.default_context(device)
    m = len(capacity)
    agg, rec = ctx.replay_tandem(capacity, A, S, total_served, warmup_fraction)
    flags = rec[L.TG_FLAGS]
    if np.any(flags != 0):
        bad = int(np.flatnonzero(flags)[0])
        raise L.MqError(-int(flags[bad]), f"replication {bad} failed (arrays too short)")
    reps = rec.shape[1]
    out = {"throughput": np.zeros(reps), "loss_prob": np.zeros(reps), "v0": np.zeros(reps),
           "mean_jobs": np.zeros((reps, m)), "served": rec[L.TG_SERVED].astype(np.int64),
           "a_used": rec[L.TG_AUSED].astype(np.int64),
           "s_used": np.stack([rec[L.TG_ROWS + i * L.TN_ROWS + L.TN_SUSED] for i in range(m)], 1).astype(np.int64),
           "agg": agg, "timing": ctx.last_timing()}
    for r in range(reps):
        elapsed = float(rec[L.TG_ELAPSED, r])
        dep, arr, lost = float(rec[L.TG_DEPARTURES, r]), float(rec[L.TG_ARRIVED, r]), float(rec[L.TG_LOST, r])
        thr = dep / elapsed
        mj = [float(rec[L.TG_ROWS + i * L.TN_ROWS + L.TN_AREA, r]) / elapsed for i in range(m)]
        out["throughput"][r] = thr
        out["loss_prob"][r] = lost / arr if arr else 0.0
        out["mean_jobs"][r] = mj
        out["v0"][r] = sum(mj) / thr if thr > 0 else 0.0
    return out
