"""Multi-class priority queue simulator: drop-in for the reference's PriorityQueueSimulator.

Mirrors most_queue/sim/priority.py: PriorityQueueSimulator(num_of_channels, num_of_classes,
prty_type, buffer) (:28), set_sources([{"type", "params"}] * k) (:111), set_servers (:134),
run(total_served) -> PriorityResults(v, w, p, utilization) (:622-657).  Disciplines NP, PR, RS and
RW run on the GPU (csrc/k4_prio.cu); "No" is rejected (in the reference it only ever serves the
class-0 queue, priority.py:521-536).  Additive: `replications=` and `seed=`.
"""

from __future__ import annotations

import os
import time

import numpy as np

from .. import _lib, dist
from ..constants import DEFAULT_NUM_STATES
from ..random.utils.load import mean_of
from ..structs import PriorityResults

Z99 = 2.5758293035489004


class PriorityQueueSimulator:
    def __init__(self, num_of_channels, num_of_classes, prty_type="No", buffer=None, seed=None, replications=1,
                 p_bins=64, queue_cap=0, device=None, devices=None):
        self.n = int(num_of_channels)
        self.k = int(num_of_classes)
        self.prty_type = prty_type
        self.buffer = buffer
        self.seed = seed
        self.replications = int(replications)
        self.p_bins = int(p_bins)
        self.queue_cap = int(queue_cap)
        self.device = device
        self.devices = devices  # "all" / list of GPU indices: one process, several GPUs (see QsSim)
        self.num_of_states = DEFAULT_NUM_STATES
        self.w = [[0, 0, 0, 0] for _ in range(self.k)]
        self.v = [[0, 0, 0, 0] for _ in range(self.k)]
        self.p = [[0.0] * self.p_bins for _ in range(self.k)]
        self.arrived = [0] * self.k
        self.served = [0] * self.k
        self.taked = [0] * self.k
        self.dropped = [0] * self.k
        self.ttek = 0
        self.sources_params = self.servers_params = None
        self.is_set_source_params = self.is_set_server_params = False
        self._src = self._srv = None
        self.ci = self.agg = self.per_replication = self.timing = None

    def set_sources(self, sources: list[dict]):
        if len(sources) != self.k:
            raise ValueError("one source per class is required")
        self._src = [_lib.make_dist(s["type"], s["params"]) for s in sources]
        self.sources_params = sources
        self.is_set_source_params = True

    def set_servers(self, servers_params: list[dict]):
        if len(servers_params) != self.k:
            raise ValueError("one service distribution per class is required")
        self._srv = [_lib.make_dist(s["type"], s["params"]) for s in servers_params]
        self.servers_params = servers_params
        self.is_set_server_params = True

    def calc_load(self):
        """priority.py:182-280: sum of arrival rates times the sum of mean service times / (n k).
        (For "Uniform" and "D" service the reference adds 1/mean instead of the mean, :258-264;
        this mirror uses the mean.)"""
        l_sum = sum(1.0 / mean_of(s["type"], s["params"]) for s in self.sources_params)
        b_sum = sum(mean_of(s["type"], s["params"]) for s in self.servers_params)
        return l_sum * b_sum / (self.n * self.k)

    def run(self, total_served: int, replications: int | None = None, keep_replications: bool = False) -> PriorityResults:
        if not (self.is_set_source_params and self.is_set_server_params):
            raise RuntimeError("set_sources() and set_servers() must be called before run()")
        start = time.process_time()
        R = int(self.replications if replications is None else replications)
        seed = dist.common_seed(self.seed)
        rep0, count = dist.shard(R)
        ctx = _lib.default_context(self.device)
        devs = dist.local_devices(self.devices)
        if devs is not None and not keep_replications and len(_lib.device_contexts(devs)) > 1:
            # single process, several GPUs: one host thread per device, aggregates added (dist.run_on_devices)
            ctxs = _lib.device_contexts(devs)
            agg = dist.run_on_devices(ctxs, R, lambda cx, r0, n: cx.sim_prio(
                self.n, self.k, self.prty_type, self.buffer, self._src, self._srv, total_served, n, r0, seed,
                p_bins=self.p_bins, queue_cap=self.queue_cap)[0])
            rec = None
            self.timing = max((cx.last_timing() for cx in ctxs), key=lambda t: t["sim_ms"])
        elif count > 0:
            agg, rec = ctx.sim_prio(self.n, self.k, self.prty_type, self.buffer, self._src, self._srv, total_served,
                                    count, rep0, seed, p_bins=self.p_bins, queue_cap=self.queue_cap,
                                    per_replication=keep_replications)
            self.timing = ctx.last_timing()
        else:
            agg, rec = np.zeros(_lib.prio_agg_len(self.k, self.p_bins)), None
        agg = dist.allreduce_sum(agg)
        self._publish(agg, rec)
        return PriorityResults(v=self.get_v(), w=self.get_w(), p=self.get_p(), duration=time.process_time() - start,
                               utilization=self.calc_load(), replications=R, ci=self.ci)

    def _publish(self, agg, rec):
        L = _lib
        R, P = agg[0], self.p_bins
        self.agg, self.per_replication = agg, rec
        if agg[2] > 0:
            raise L.MqError(L.MQ_E_OVERFLOW, f"{int(agg[2])} replications overflowed a waiting line of the engine; "
                                             "raise queue_cap")

        def mean_ci(s, s2):
            m = s / R
            if R > 1:
                var = np.maximum(s2 / R - m * m, 0.0) * R / (R - 1)
                return m, Z99 * np.sqrt(var / R)
            return m, np.full_like(m, np.nan)

        self.ttek = float(agg[1] / R)
        self.w, self.v, self.p = [], [], []
        ci = {"w": [], "v": [], "p": [], "level": 0.99}
        one = R == 1
        cast = (lambda x: int(round(x))) if one else (lambda x: float(x) / R)
        for k in range(self.k):
            b = 4 + k * L.prio_agg(P)
            w, wci = mean_ci(agg[b:b + 4], agg[b + 4:b + 8])
            v, vci = mean_ci(agg[b + 8:b + 12], agg[b + 12:b + 16])
            p, pci = mean_ci(agg[b + 24:b + 24 + P], agg[b + 24 + P:b + 24 + 2 * P])
            self.w.append([float(x) for x in w])
            self.v.append([float(x) for x in v])
            self.p.append([float(x) for x in p])
            ci["w"].append([float(x) for x in wci])
            ci["v"].append([float(x) for x in vci])
            ci["p"].append([float(x) for x in pci])
            self.arrived[k], self.taked[k] = cast(agg[b + 16]), cast(agg[b + 17])
            self.served[k], self.dropped[k] = cast(agg[b + 18]), cast(agg[b + 19])
        self.ci = ci

    def get_p(self):
        res = []
        for k in range(self.k):
            row = list(self.p[k])
            if len(row) < self.num_of_states:
                row.extend([0.0] * (self.num_of_states - len(row)))
            res.append(row)
        return res

    def get_v(self):
        return self.v

    def get_w(self):
        return self.w


def replay(num_of_channels, num_of_classes, prty_type, A, S, total_served, buffer=None, p_bins=64, queue_cap=0,
           device=None):
    """Replay tier: per-class arrays A[k][i, r], S[k][j, r]; returns per-replication results whose
    `w`, `v` equal the reference's running-mean moments bit for bit (csrc/k4_prio.cu)."""
    L = _lib
    ctx = L.default_context(device)
    K = int(num_of_classes)
    agg, rec = ctx.replay_prio(num_of_channels, K, prty_type, buffer, A, S, total_served, p_bins=p_bins,
                               queue_cap=queue_cap)
    flags = rec[L.PG_FLAGS]
    if np.any(flags != 0):
        bad = int(np.flatnonzero(flags)[0])
        raise L.MqError(-int(flags[bad]), f"replication {bad} failed (arrays too short or waiting line overflow)")
    rows = L.prio_rows(p_bins)
    out = {"ttek": rec[L.PG_TTEK].copy(), "agg": agg, "timing": ctx.last_timing()}
    per = lambda f, n=1: np.stack([rec[L.PG_ROWS + k * rows + f: L.PG_ROWS + k * rows + f + n] for k in range(K)])
    out["w"] = per(L.PF_WRUN, 4).transpose(2, 0, 1).copy()   # [rep][class][4]
    out["v"] = per(L.PF_VRUN, 4).transpose(2, 0, 1).copy()
    out["w_sum"] = per(L.PF_WSUM, 4).transpose(2, 0, 1).copy()
    out["v_sum"] = per(L.PF_VSUM, 4).transpose(2, 0, 1).copy()
    out["p_time"] = per(L.PF_P, p_bins).transpose(2, 0, 1).copy()
    for name, f in (("arrived", L.PF_ARRIVED), ("taked", L.PF_TAKED), ("served", L.PF_SERVED),
                    ("dropped", L.PF_DROPPED), ("in_sys", L.PF_INSYS), ("a_used", L.PF_AUSED),
                    ("s_used", L.PF_SUSED)):
        out[name] = per(f)[:, 0, :].T.astype(np.int64)
    out["t_old"] = per(L.PF_TOLD)[:, 0, :].T.copy()
    return out
