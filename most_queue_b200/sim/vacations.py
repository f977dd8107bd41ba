"""Queue with warm-up, cold (vacation) and cold-delay periods, and the N-policy queue: drop-ins for the
reference's VacationQueueingSystemSimulator and NPolicyQueueSim, computed on the GPU.

Mirrors most_queue/sim/vacations.py: VacationQueueingSystemSimulator(num_of_channels, buffer, verbose,
buffer_type, is_service_on_warm_up, is_multiple_vacations) (:17), set_warm (:52), set_cold (:79),
set_cold_delay (:89), run(total_served) -> VacationResults(v, w, p, utilization, warmup_prob, cold_prob)
(:303-319), get_warmup_prob / get_cold_prob / get_cold_delay_prob (:321-338), and NPolicyQueueSim(num_of_channels,
big_n) (:320-364).  The event loop runs in csrc/k8_vacation.cu, one replication per lane.  Additive: `seed=` and
`replications=` as in QsSim.
"""

from __future__ import annotations

import os
import time

import numpy as np

from .. import _lib, dist
from ..structs import VacationResults
from .base import Z99, QsSim


class VacationQueueingSystemSimulator(QsSim):
    def __init__(self, num_of_channels, buffer=None, verbose=True, buffer_type="list", is_service_on_warm_up=False,
                 is_multiple_vacations=False, seed=None, replications=1, p_bins=128, queue_cap=0, device=None):
        super().__init__(num_of_channels, buffer, verbose, buffer_type, seed=seed, replications=replications,
                         p_bins=p_bins, device=device)
        self.is_service_on_warm_up = bool(is_service_on_warm_up)
        self.is_multiple_vacations = bool(is_multiple_vacations)
        self.queue_cap = int(queue_cap)
        self.big_n = 0
        self._warm = self._cold = self._delay = None
        self.warmup_prob = self.cold_prob = self.cold_delay_prob = 0.0
        self.warm_after_cold_starts = 0

    def set_warm(self, params, kendall_notation):
        if not self.is_set_server_params:
            raise ValueError("Server parameters are not set. Please call set_servers() first.")  # vacations.py:61
        self._warm = _lib.make_dist(kendall_notation, params)

    def set_cold(self, params, kendall_notation):
        self._cold = _lib.make_dist(kendall_notation, params)

    def set_cold_delay(self, params, kendall_notation):
        if self._cold is None and not self.big_n:
            raise ValueError("You must first set the cooling time. Use the set_cold() method.")  # vacations.py:99
        self._delay = _lib.make_dist(kendall_notation, params)

    def _dists(self):
        d = {"arrivals": self._src, "services": self._srv, "cold": self._cold, "delay": self._delay}
        if self._warm is not None:
            d["warm"] = self._warm  # the system-level warm-up period (vacations.py:64)
            if self.is_service_on_warm_up:
                d["warm_services"] = self._warm  # and the law on every server (vacations.py:66-76)
        return d

    def run(self, total_served: int, replications: int | None = None, keep_replications: bool = False) -> VacationResults:
        if not (self.is_set_source_params and self.is_set_server_params):
            raise RuntimeError("set_sources() and set_servers() must be called before run()")
        start = time.process_time()
        R = int(self.replications if replications is None else replications)
        if R < 1:
            raise ValueError("replications must be >= 1")
        seed = dist.common_seed(self.seed)
        rep0, count = dist.shard(R)
        ctx = _lib.default_context(self.device)
        if count > 0:
            agg, rec = ctx.sim_vacation(self.n, self.buffer, self._dists(), total_served, count, rep0, seed,
                                        service_on_warm_up=self.is_service_on_warm_up,
                                        multiple_vacations=self.is_multiple_vacations, big_n=self.big_n,
                                        p_bins=self.p_bins, queue_cap=self.queue_cap,
                                        per_replication=keep_replications)
            self.timing = ctx.last_timing()
        else:
            agg, rec = np.zeros(_lib.vac_agg_len(self.p_bins)), None
        agg = dist.allreduce_sum(agg)
        self._publish_vac(agg, rec)
        self.time_spent = time.process_time() - start
        return VacationResults(v=self.get_v(), w=self.get_w(), p=self.get_p(), duration=self.time_spent,
                               utilization=self.calc_load(), replications=R, ci=self.ci,
                               warmup_prob=self.get_warmup_prob(), cold_prob=self.get_cold_prob(),
                               cold_delay_prob=self.get_cold_delay_prob())

    def _publish_vac(self, agg, rec):
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

        w, wci = mean_ci(agg[8:12], agg[12:16])
        v, vci = mean_ci(agg[16:20], agg[20:24])
        ph, phci = mean_ci(agg[24:27], agg[27:30])
        p, pci = mean_ci(agg[32:32 + P], agg[32 + P:32 + 2 * P])
        self.w, self.v, self.p = [float(x) for x in w], [float(x) for x in v], [float(x) for x in p]
        self.warmup_prob, self.cold_prob, self.cold_delay_prob = (float(x) for x in ph)
        self.ci = {"w": [float(x) for x in wci], "v": [float(x) for x in vci], "p": [float(x) for x in pci],
                   "phases": [float(x) for x in phci], "level": 0.99}
        cast = (lambda x: int(round(x))) if R == 1 else (lambda x: float(x) / R)
        self.served, self.arrived, self.dropped, self.taked = (cast(agg[i]) for i in (3, 4, 5, 6))
        self.ttek = float(agg[1] / R)

    def get_warmup_prob(self):
        return self.warmup_prob

    def get_cold_prob(self):
        return self.cold_prob

    def get_cold_delay_prob(self):
        return self.cold_delay_prob


class NPolicyQueueSim(VacationQueueingSystemSimulator):
    """The server switches off when the system empties and resumes when N jobs have accumulated
    (vacations.py:320-364)."""

    def __init__(self, num_of_channels: int, big_n: int, **kwargs):
        super().__init__(num_of_channels, **kwargs)
        if big_n < 1:
            raise ValueError(f"N must be >= 1, got {big_n}")
        self.big_n = int(big_n)


def replay(num_of_channels, buffer, streams: dict, total_served, is_service_on_warm_up=False,
           is_multiple_vacations=False, big_n=0, p_bins=128, queue_cap=0, device=None):
    """Replay tier: the event loop on pre-generated draws, fp64, bit-exact with the reference.  streams maps
    arrivals / services / warm_services / warm / cold / delay to [n, replications] arrays (absent = law not set)."""
    L = _lib
    ctx = L.default_context(device)
    agg, rec = ctx.replay_vacation(num_of_channels, buffer, streams, total_served,
                                   service_on_warm_up=is_service_on_warm_up, multiple_vacations=is_multiple_vacations,
                                   big_n=big_n, p_bins=p_bins, queue_cap=queue_cap)
    flags = rec[L.V_FLAGS]
    if np.any(flags != 0):
        bad = int(np.flatnonzero(flags)[0])
        raise L.MqError(-int(flags[bad]), f"replication {bad} failed (arrays too short or waiting line overflow)")
    i64 = lambda r: rec[r].astype(np.int64)
    return {
        "w": rec[L.V_WRUN:L.V_WRUN + 4].T.copy(), "v": rec[L.V_VRUN:L.V_VRUN + 4].T.copy(),
        "w_sum": rec[L.V_WSUM:L.V_WSUM + 4].T.copy(), "v_sum": rec[L.V_VSUM:L.V_VSUM + 4].T.copy(),
        "busy": rec[L.V_BUSYRUN:L.V_BUSYRUN + 3].T.copy(), "busy_n": i64(L.V_BUSYN),
        "p_time": rec[L.V_P:L.V_P + p_bins].T.copy(), "p_over": rec[L.V_POVER].copy(), "ttek": rec[L.V_TTEK].copy(),
        "arrived": i64(L.V_ARRIVED), "taked": i64(L.V_TAKED), "served": i64(L.V_SERVED), "dropped": i64(L.V_DROPPED),
        "in_sys": i64(L.V_INSYS), "queued": i64(L.V_QUEUED),
        "warm_prob": rec[L.V_WARMPROB].copy(), "cold_prob": rec[L.V_COLDPROB].copy(),
        "delay_prob": rec[L.V_DELAYPROB].copy(), "starts": rec[L.V_STARTS:L.V_STARTS + 3].T.astype(np.int64),
        "warm_after_cold": i64(L.V_WARM_AFTER_COLD),
        "used": {k: i64(L.V_USED + i) for i, k in enumerate(L.VAC_STREAMS)},
        "agg": agg, "timing": ctx.last_timing(),
    }
