"""Open network of priority nodes: drop-in for the reference's PriorityNetwork.

Mirrors most_queue/sim/networks/priority_network.py: PriorityNetwork(k_num) (:28),
set_sources(L, R) (:60), set_nodes(serv_params, n, prty, nodes_prty) (:87) and
run(job_served) -> NetworkResultsPriority(v, arrived, served, duration) (:241-276).  After run()
the object also carries v_network / w_network per class as the reference does.  The event loop
runs on the GPU (csrc/k7_prionet.cu), one replication per lane.  Additive: `seed=` (the reference
network is not seedable) and `replications=`.
"""

from __future__ import annotations

import os
import time

import numpy as np

from ... import _lib, dist
from ...structs import NetworkResultsPriority

Z99 = 2.5758293035489004


class PriorityNetwork:
    def __init__(self, k_num: int, seed=None, replications=1, queue_cap=0, device=None):
        self.k_num = int(k_num)
        self.seed = seed
        self.replications = int(replications)
        self.queue_cap = int(queue_cap)
        self.device = device
        self.L = self.R = None
        self.n_num = self.nodes = self.prty = None
        self.serv_params = self.nodes_prty = None
        self._src = self._srv = None
        self.is_sources_set = self.is_nodes_set = False
        self.v_network = [[0.0] * 3 for _ in range(self.k_num)]
        self.w_network = [[0.0] * 3 for _ in range(self.k_num)]
        self.ttek = 0
        self.served = [0] * self.k_num
        self.arrived = [0] * self.k_num
        self.results = None
        self.ci = self.agg = self.per_replication = self.timing = None

    def set_sources(self, L: list[float], R: list):
        if len(L) != self.k_num or len(R) != self.k_num:
            raise ValueError(f"need {self.k_num} arrival rates and routing matrices")
        self.L = list(L)
        self.R = np.stack([np.asarray(r, dtype=np.float64) for r in R])
        self._src = [_lib.make_dist("M", float(l)) for l in L]  # priority_network.py:82: ExpDistribution(L[k])
        self.is_sources_set = True

    def set_nodes(self, serv_params: list[list[dict]], n: list[int], prty: list[str], nodes_prty: list[list[int]]):
        self.serv_params = serv_params
        self.n_num = len(n)
        self.nodes = list(n)
        self.prty = list(prty)
        self.nodes_prty = [list(row) for row in nodes_prty]
        # node class j of node i is served with serv_params[i][nodes_prty[i][j]]  (:129-131)
        self._srv = []
        for i in range(self.n_num):
            row = []
            for j in range(self.k_num):
                s = serv_params[i][self.nodes_prty[i][j]]
                row.append(_lib.make_dist(s["type"], s["params"]))
            self._srv.append(row)
        self.is_nodes_set = True

    def _check_sources_and_nodes_is_set(self):
        if not self.is_sources_set:
            raise ValueError("Sources are not set. Please use set_sources() method.")
        if not self.is_nodes_set:
            raise ValueError("Nodes are not set. Please use set_nodes() method.")

    def run(self, job_served: int, replications: int | None = None,
            keep_replications: bool = False) -> NetworkResultsPriority:
        start = time.process_time()
        self._check_sources_and_nodes_is_set()
        R = int(self.replications if replications is None else replications)
        seed = dist.common_seed(self.seed)
        rep0, count = dist.shard(R)
        ctx = _lib.default_context(self.device)
        if count > 0:
            agg, rec = ctx.sim_prionet(self.R, self.nodes, self.prty, self.nodes_prty, self._src, self._srv,
                                       job_served, count, rep0, seed, queue_cap=self.queue_cap,
                                       per_replication=keep_replications)
            self.timing = ctx.last_timing()
        else:
            agg, rec = np.zeros(_lib.prionet_agg_len(self.k_num)), None
        agg = dist.allreduce_sum(agg)
        self._publish(agg, rec)
        self.results = NetworkResultsPriority(v=self.v_network, arrived=self.arrived, served=self.served,
                                              duration=time.process_time() - start, w=self.w_network,
                                              replications=R, ci=self.ci)
        return self.results

    def _publish(self, agg, rec):
        reps = agg[0]
        self.agg, self.per_replication = agg, rec
        if agg[2] > 0:
            raise _lib.MqError(_lib.MQ_E_OVERFLOW,
                               f"{int(agg[2])} replications overflowed a waiting line; raise queue_cap")

        def ci(s, s2):
            mean = s / reps
            if reps > 1:
                var = np.maximum(s2 / reps - mean * mean, 0.0) * reps / (reps - 1)
                return Z99 * np.sqrt(var / reps)
            return np.full_like(mean, np.nan)

        cast = (lambda x: int(round(x))) if reps == 1 else (lambda x: float(x) / reps)
        self.ttek = float(agg[1] / reps)
        self.v_network, self.w_network, self.served, self.arrived = [], [], [], []
        civ, ciw = [], []
        for k in range(self.k_num):
            b = 4 + 16 * k
            self.v_network.append([float(x) for x in agg[b:b + 3] / reps])
            self.w_network.append([float(x) for x in agg[b + 6:b + 9] / reps])
            civ.append([float(x) for x in ci(agg[b:b + 3], agg[b + 3:b + 6])])
            ciw.append([float(x) for x in ci(agg[b + 6:b + 9], agg[b + 9:b + 12])])
            self.served.append(cast(agg[b + 12]))
            self.arrived.append(cast(agg[b + 13]))
        self.ci = {"v": civ, "w": ciw, "level": 0.99}


def replay(R, channels, prty, nodes_prty, A, U, S, job_served, queue_cap=0, device=None):
    """Replay tier: per-class external gaps A[k][i, r], routing uniforms U[j, r], draws S[node][node class][j, r];
    per-replication results, bit-exact with the reference in fp64 (csrc/k7_prionet.cu)."""
    L = _lib
    ctx = L.default_context(device)
    m, K = len(channels), len(A)
    agg, rec = ctx.replay_prionet(R, channels, prty, nodes_prty, A, U, S, job_served, queue_cap=queue_cap)
    flags = rec[L.PN_G_FLAGS]
    if np.any(flags != 0):
        bad = int(np.flatnonzero(flags)[0])
        raise L.MqError(-int(flags[bad]), f"replication {bad} failed (arrays too short or waiting line overflow)")
    reps = rec.shape[1]
    cls = rec[L.PN_G_ROWS:L.PN_G_ROWS + K * L.PNC_ROWS].reshape(K, L.PNC_ROWS, reps)
    nod = rec[L.PN_G_ROWS + K * L.PNC_ROWS:].reshape(m, K, L.PNN_ROWS, reps)
    c3 = lambda f: cls[:, f:f + 3].transpose(2, 0, 1).copy()       # [rep][K][3]
    n4 = lambda f: nod[:, :, f:f + 4].transpose(3, 0, 1, 2).copy()  # [rep][m][K][4]
    ni = lambda f: nod[:, :, f].transpose(2, 0, 1).astype(np.int64)
    return {
        "ttek": rec[L.PN_G_TTEK].copy(), "u_used": rec[L.PN_G_UUSED].astype(np.int64),
        "v": c3(L.PNC_VRUN), "w": c3(L.PNC_WRUN), "v_sum": c3(L.PNC_VSUM), "w_sum": c3(L.PNC_WSUM),
        "served": cls[:, L.PNC_SERVED].T.astype(np.int64), "arrived": cls[:, L.PNC_ARRIVED].T.astype(np.int64),
        "a_used": cls[:, L.PNC_AUSED].T.astype(np.int64),
        "node_v": n4(L.PNN_VRUN), "node_w": n4(L.PNN_WRUN), "node_v_sum": n4(L.PNN_VSUM),
        "node_w_sum": n4(L.PNN_WSUM), "node_served": ni(L.PNN_SERVED), "node_taked": ni(L.PNN_TAKED),
        "s_used": ni(L.PNN_SUSED), "agg": agg, "timing": ctx.last_timing(),
    }
