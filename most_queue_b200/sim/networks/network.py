"""Open queueing network simulator: drop-in for the reference's NetworkSimulator.

Mirrors most_queue/sim/networks/network.py: NetworkSimulator().set_sources(arrival_rate, R,
source_kendall, source_params) (:53), set_nodes(serv_params, n) (:82), run(job_served) ->
NetworkResults(v, served, arrived, duration) (:196-229); after run() the object also carries
v_network / w_network and, per node, the node-level w / v moments (qs[i].w, qs[i].v).  The event
loop runs on the GPU (csrc/k5_net.cu).  Additive: `seed=` (the reference's network is not
seedable) and `replications=`.
"""

from __future__ import annotations

import os
import time
from types import SimpleNamespace

import numpy as np

from ... import _lib, dist
from ...structs import NetworkResults

Z99 = 2.5758293035489004


class NetworkSimulator:
    def __init__(self, seed=None, replications=1, queue_cap=0, device=None, devices=None):
        self.seed = seed
        self.replications = int(replications)
        self.queue_cap = int(queue_cap)
        self.device = device
        self.devices = devices  # "all" / list of GPU indices: one process, several GPUs (see QsSim)
        self.R = None
        self.n_num = None
        self.nodes = None
        self.serv_params = None
        self.arrival_rate = None
        self._src = self._srv = None
        self.is_sources_set = self.is_nodes_set = False
        self.v_network = [0.0] * 3
        self.w_network = [0.0] * 3
        self.ttek = 0
        self.served = self.arrived = self.in_sys = 0
        self.qs = []
        self.results = None
        self.ci = self.agg = self.per_replication = self.timing = None

    def set_sources(self, arrival_rate: float, R, source_kendall: str = "M", source_params=None):
        self.arrival_rate = arrival_rate
        if source_kendall == "M" and source_params is None:
            self._src = _lib.make_dist("M", arrival_rate)
        else:
            self._src = _lib.make_dist(source_kendall, source_params)
        self.R = np.asarray(R, dtype=np.float64)
        self.is_sources_set = True

    def set_nodes(self, serv_params: list[dict], n: list[int]):
        self.serv_params = serv_params
        self.nodes = list(n)
        self.n_num = len(n)
        self._srv = [_lib.make_dist(s["type"], s["params"]) for s in serv_params]
        self.is_nodes_set = True

    def _check_sources_and_nodes_is_set(self):
        if not self.is_sources_set:
            raise ValueError("Sources are not set. Please use set_sources() method.")
        if not self.is_nodes_set:
            raise ValueError("Nodes are not set. Please use set_nodes() method.")

    def run(self, job_served: int, replications: int | None = None, keep_replications: bool = False) -> NetworkResults:
        start = time.process_time()
        self._check_sources_and_nodes_is_set()
        R = int(self.replications if replications is None else replications)
        seed = dist.common_seed(self.seed)
        rep0, count = dist.shard(R)
        ctx = _lib.default_context(self.device)
        m = self.n_num
        devs = dist.local_devices(self.devices)
        if devs is not None and not keep_replications and len(_lib.device_contexts(devs)) > 1:
            # single process, several GPUs: one host thread per device, aggregates added (dist.run_on_devices)
            ctxs = _lib.device_contexts(devs)
            agg = dist.run_on_devices(ctxs, R, lambda cx, r0, n: cx.sim_net(
                self.R, self.nodes, self._src, self._srv, job_served, n, r0, seed, queue_cap=self.queue_cap)[0])
            rec = None
            self.timing = max((cx.last_timing() for cx in ctxs), key=lambda t: t["sim_ms"])
        elif count > 0:
            agg, rec = ctx.sim_net(self.R, self.nodes, self._src, self._srv, job_served, count, rep0, seed,
                                   queue_cap=self.queue_cap, per_replication=keep_replications)
            self.timing = ctx.last_timing()
        else:
            agg, rec = np.zeros(_lib.net_agg_len(m)), None
        agg = dist.allreduce_sum(agg)
        self._publish(agg, rec)
        self.results = NetworkResults(v=self.v_network, served=self.served, arrived=self.arrived,
                                      duration=time.process_time() - start, w=self.w_network, replications=R,
                                      ci=self.ci)
        return self.results

    def _publish(self, agg, rec):
        L = _lib
        reps = agg[0]
        self.agg, self.per_replication = agg, rec
        if agg[2] > 0:
            raise L.MqError(L.MQ_E_OVERFLOW, f"{int(agg[2])} replications overflowed a node's waiting line; raise queue_cap")

        def ci(s, s2):
            mean = s / reps
            if reps > 1:
                var = np.maximum(s2 / reps - mean * mean, 0.0) * reps / (reps - 1)
                return Z99 * np.sqrt(var / reps)
            return np.full_like(mean, np.nan)

        self.v_network = [float(x) for x in agg[5:8] / reps]
        self.w_network = [float(x) for x in agg[11:14] / reps]
        self.ci = {"v": [float(x) for x in ci(agg[5:8], agg[8:11])],
                   "w": [float(x) for x in ci(agg[11:14], agg[14:17])], "level": 0.99}
        one = reps == 1
        cast = (lambda x: int(round(x))) if one else (lambda x: float(x) / reps)
        self.served, self.arrived = cast(agg[3]), cast(agg[4])
        self.ttek = float(agg[1] / reps)
        self.qs = []
        for i in range(self.n_num):
            b = 20 + 16 * i
            self.qs.append(SimpleNamespace(v=[float(x) for x in agg[b:b + 4] / reps],
                                           w=[float(x) for x in agg[b + 4:b + 8] / reps],
                                           served=cast(agg[b + 8]), taked=cast(agg[b + 9]),
                                           arrived=cast(agg[b + 10]), n=self.nodes[i]))


def replay(R, channels, A, U, S, job_served, queue_cap=0, device=None):
    """Replay tier: external gaps A[i, r], routing uniforms U[j, r], node services S[node][j, r];
    per-replication results, bit-exact with the reference in fp64 (csrc/k5_net.cu)."""
    L = _lib
    ctx = L.default_context(device)
    m = len(channels)
    agg, rec = ctx.replay_net(R, channels, A, U, S, job_served, queue_cap=queue_cap)
    flags = rec[L.NG_FLAGS]
    if np.any(flags != 0):
        bad = int(np.flatnonzero(flags)[0])
        raise L.MqError(-int(flags[bad]), f"replication {bad} failed (arrays too short or waiting line overflow)")
    node = lambda f, n=1: np.stack([rec[L.NG_ROWS + i * L.NN_ROWS + f: L.NG_ROWS + i * L.NN_ROWS + f + n]
                                    for i in range(m)])
    out = {
        "ttek": rec[L.NG_TTEK].copy(), "served": rec[L.NG_SERVED].astype(np.int64),
        "arrived": rec[L.NG_ARRIVED].astype(np.int64), "in_sys": rec[L.NG_INSYS].astype(np.int64),
        "a_used": rec[L.NG_AUSED].astype(np.int64), "u_used": rec[L.NG_UUSED].astype(np.int64),
        "v": rec[L.NG_VRUN:L.NG_VRUN + 3].T.copy(), "w": rec[L.NG_WRUN:L.NG_WRUN + 3].T.copy(),
        "v_sum": rec[L.NG_VSUM:L.NG_VSUM + 3].T.copy(), "w_sum": rec[L.NG_WSUM:L.NG_WSUM + 3].T.copy(),
        "node_v": node(L.NN_VRUN, 4).transpose(2, 0, 1).copy(), "node_w": node(L.NN_WRUN, 4).transpose(2, 0, 1).copy(),
        "node_v_sum": node(L.NN_VSUM, 4).transpose(2, 0, 1).copy(),
        "node_w_sum": node(L.NN_WSUM, 4).transpose(2, 0, 1).copy(),
        "node_served": node(L.NN_SERVED)[:, 0, :].T.astype(np.int64),
        "node_taked": node(L.NN_TAKED)[:, 0, :].T.astype(np.int64),
        "node_arrived": node(L.NN_ARRIVED)[:, 0, :].T.astype(np.int64),
        "s_used": node(L.NN_SUSED)[:, 0, :].T.astype(np.int64),
        "agg": agg, "timing": ctx.last_timing(),
    }
    return out
