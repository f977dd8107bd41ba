"""Queue with negative arrivals: drop-in for the reference's QsSimNegatives, computed on the GPU.

Mirrors most_queue/sim/negative.py: NegativeServiceType / DisasterScenario / RcsScenario (:16-54),
QsSimNegatives(num_of_channels, type_of_negatives, buffer, verbose, buffer_type, disaster_scenario, rcs_scenario)
(:62), set_positive_sources (:116), set_negative_sources (:134), set_servers, run(total_served) ->
NegativeArrivalsResults(v, w, p, v_served, v_broken, utilization, q) (:184-203), get_v_served / get_v_broken.
Implemented on the GPU (csrc/k9_negatives.cu): the DISASTER and RCS types with all four scenarios each; the RCH / RCE
types raise NotImplementedError (their reference code reads the tail of the line without removing it in the default
"list" buffer and fails on an empty line, negative.py:306-336).  Additive: `seed=`, `replications=`.
"""

from __future__ import annotations

import os
import time
from enum import Enum

import numpy as np

from .. import _lib, dist
from ..random.utils.load import calc_qs_load
from ..structs import NegativeArrivalsResults
from .base import Z99, QsSim


class NegativeServiceType(Enum):
    DISASTER = 1  # remove all customers
    RCS = 2  # remove customer in service
    RCH = 3  # remove customer at the head
    RCE = 4  # remove customer at the end


class DisasterScenario(Enum):
    CLEAR_SYSTEM = 1
    REQUEUE_ALL = 2
    RESUME_ALL = 3
    REQUEUE_ALL_NO_RESAMPLING = 4


class RcsScenario(Enum):
    REMOVE = 1
    REQUEUE = 2
    RESUME = 3
    REQUEUE_NO_RESAMPLING = 4


def _engine_type(type_of_negatives, disaster_scenario, rcs_scenario):
    """(type code, scenario code) of the C ABI."""
    name = getattr(type_of_negatives, "name", str(type_of_negatives))
    if name == "DISASTER":
        return _lib.NEG_TYPES["DISASTER"], _lib.NEG_SCENARIOS[getattr(disaster_scenario, "name", str(disaster_scenario))]
    if name == "RCS":
        return _lib.NEG_TYPES["RCS"], _lib.NEG_SCENARIOS[getattr(rcs_scenario, "name", str(rcs_scenario))]
    raise NotImplementedError(f"negatives of type {name} are not implemented on the GPU engine (available: DISASTER, RCS)")


class QsSimNegatives(QsSim):
    def __init__(self, num_of_channels, type_of_negatives=NegativeServiceType.DISASTER, buffer=None, verbose=True,
                 buffer_type="list", disaster_scenario=DisasterScenario.CLEAR_SYSTEM, rcs_scenario=RcsScenario.REMOVE,
                 seed=None, replications=1, p_bins=128, queue_cap=0, device=None):
        super().__init__(num_of_channels, buffer, verbose, buffer_type, seed=seed, replications=replications,
                         p_bins=p_bins, device=device)
        self.type_of_negatives = type_of_negatives
        self.disaster_scenario = disaster_scenario
        self.rcs_scenario = rcs_scenario
        self._type, self._scenario = _engine_type(type_of_negatives, disaster_scenario, rcs_scenario)
        self.queue_cap = int(queue_cap)
        self.v_served = [0, 0, 0, 0]
        self.v_broken = [0, 0, 0, 0]
        self.broken = self.total = 0
        self.positive_arrived = self.negative_arrived = 0
        self.positive_source_params = self.positive_source_kendall_notation = None
        self.negative_source_params = self.negative_source_kendall_notation = None
        self.is_set_positive_source_params = self.is_set_negative_source_params = False
        self._pos = self._neg = None
        self.q = 0.0

    def set_positive_sources(self, params, kendall_notation: str = "M"):
        self._pos = _lib.make_dist(kendall_notation, params)
        self.positive_source_params, self.positive_source_kendall_notation = params, kendall_notation
        self.is_set_positive_source_params = True

    def set_negative_sources(self, params, kendall_notation: str = "M"):
        self._neg = _lib.make_dist(kendall_notation, params)
        self.negative_source_params, self.negative_source_kendall_notation = params, kendall_notation
        self.is_set_negative_source_params = True

    def calc_positive_load(self):
        return calc_qs_load(self.positive_source_kendall_notation, self.positive_source_params,
                            self.server_kendall_notation, self.server_params, self.n)

    def run(self, total_served: int, replications: int | None = None,
            keep_replications: bool = False) -> NegativeArrivalsResults:
        if not (self.is_set_positive_source_params and self.is_set_negative_source_params
                and self.is_set_server_params):
            raise RuntimeError("set_positive_sources(), set_negative_sources() and set_servers() must be called "
                               "before run()")
        start = time.process_time()
        R = int(self.replications if replications is None else replications)
        if R < 1:
            raise ValueError("replications must be >= 1")
        seed = dist.common_seed(self.seed)
        rep0, count = dist.shard(R)
        ctx = _lib.default_context(self.device)
        if count > 0:
            agg, rec = ctx.sim_negatives(self.n, self.buffer, self._type, self._pos, self._neg, self._srv, total_served,
                                         count, rep0, seed, p_bins=self.p_bins, queue_cap=self.queue_cap,
                                         per_replication=keep_replications, scenario=self._scenario)
            self.timing = ctx.last_timing()
        else:
            agg, rec = np.zeros(_lib.neg_agg_len(self.p_bins)), None
        agg = dist.allreduce_sum(agg)
        self._publish_neg(agg, rec)
        self.time_spent = time.process_time() - start
        return NegativeArrivalsResults(v=self.get_v(), w=self.get_w(), p=self.get_p(), v_served=self.v_served,
                                       v_broken=self.v_broken, duration=self.time_spent,
                                       utilization=self.calc_positive_load(), q=self.q, replications=R, ci=self.ci)

    def _publish_neg(self, agg, rec):
        R, P = agg[0], self.p_bins
        self.agg, self.per_replication = agg, rec
        if agg[2] > 0:
            raise _lib.MqError(_lib.MQ_E_OVERFLOW,
                               f"{int(agg[2])} replications overflowed the waiting line; raise queue_cap")

        def mean_ci(s, s2):
            m = s / R
            if R > 1:
                var = np.maximum(s2 / R - m * m, 0.0) * R / (R - 1)
                return m, Z99 * np.sqrt(var / R)
            return m, np.full_like(m, np.nan)

        out, ci = {}, {}
        for name, b in (("w", 8), ("v", 16), ("v_served", 24), ("v_broken", 32)):
            m, h = mean_ci(agg[b:b + 4], agg[b + 4:b + 8])
            out[name], ci[name] = [float(x) for x in m], [float(x) for x in h]
        self.w, self.v, self.v_served, self.v_broken = out["w"], out["v"], out["v_served"], out["v_broken"]
        q, qci = mean_ci(agg[40:41], agg[41:42])
        p, pci = mean_ci(agg[48:48 + P], agg[48 + P:48 + 2 * P])
        self.q, self.p = float(q[0]), [float(x) for x in p]
        ci.update({"q": float(qci[0]), "p": [float(x) for x in pci], "level": 0.99})
        self.ci = ci
        cast = (lambda x: int(round(x))) if R == 1 else (lambda x: float(x) / R)
        self.served, self.broken, self.total, self.positive_arrived, self.dropped = (cast(agg[i]) for i in range(3, 8))
        self.ttek = float(agg[1] / R)

    def get_v_served(self):
        return self.v_served

    def get_v_broken(self):
        return self.v_broken


def replay(num_of_channels, buffer, type_of_negatives, streams: dict, total_served, p_bins=128, queue_cap=0,
           device=None, scenario=None):
    """Replay tier: the event loop on pre-generated draws, fp64, bit-exact with the reference.  streams maps
    positives / negatives / services (and choices for RCS) to [n, replications] arrays."""
    L = _lib
    ctx = L.default_context(device)
    tt = type_of_negatives if not isinstance(type_of_negatives, str) else NegativeServiceType[type_of_negatives]
    sc = scenario if scenario is not None else ("CLEAR_SYSTEM" if tt.name == "DISASTER" else "REMOVE")
    t, s_code = _engine_type(tt, sc, sc)
    agg, rec = ctx.replay_negatives(num_of_channels, buffer, t, streams, total_served, p_bins=p_bins,
                                    queue_cap=queue_cap, scenario=s_code)
    flags = rec[L.N_FLAGS]
    if np.any(flags != 0):
        bad = int(np.flatnonzero(flags)[0])
        raise L.MqError(-int(flags[bad]), f"replication {bad} failed (arrays too short or waiting line overflow)")
    i64 = lambda r: rec[r].astype(np.int64)
    m4 = lambda r: rec[r:r + 4].T.copy()
    return {
        "w": m4(L.N_WRUN), "v": m4(L.N_VRUN), "v_served": m4(L.N_VSRUN), "v_broken": m4(L.N_VBRUN),
        "w_sum": m4(L.N_WSUM), "v_sum": m4(L.N_VSUM), "vs_sum": m4(L.N_VSSUM), "vb_sum": m4(L.N_VBSUM),
        "p_time": rec[L.N_P:L.N_P + p_bins].T.copy(), "p_over": rec[L.N_POVER].copy(), "ttek": rec[L.N_TTEK].copy(),
        "positive_arrived": i64(L.N_POS_ARRIVED), "negative_arrived": i64(L.N_NEG_ARRIVED), "taked": i64(L.N_TAKED),
        "served": i64(L.N_SERVED), "broken": i64(L.N_BROKEN), "total": i64(L.N_TOTAL), "dropped": i64(L.N_DROPPED),
        "in_sys": i64(L.N_INSYS), "queued": i64(L.N_QUEUED),
        "used": {k: i64(L.N_USED + i) for i, k in enumerate(L.NEG_STREAMS)},
        "agg": agg, "timing": ctx.last_timing(),
    }
