"""Size-based single-server queue (SJF, PSJF, SRPT, ...): drop-in for the reference's SizeBasedQsSim on the GPU.

Mirrors most_queue/sim/size_based.py: SizeBasedQsSim(num_of_channels=1, discipline, buffer, verbose, buffer_type,
track_slowdown) (:72), set_sources / set_servers (:121), set_predictor (:105), run(total_served) -> QueueResults,
get_slowdown (:108).  Job sizes are drawn at arrival; the waiting line is ordered by (rank, arrival time, creation
order); PSJF, SRPT, PSPJF and SPRPT pre-empt the job in service when a strictly better job arrives.  SPJF / PSPJF /
SPRPT rank by the PREDICTED size (size_based.py:128-153); the predictors of sim/utils/predictor.py are built into the
engine (perfect, exponential noise, log-normal noise); any other predictor is a Python callback and is rejected.  The event loop runs in csrc/k10_size.cu.  Additive: `seed=`, `replications=`;
`get_slowdown()` returns the mean of sojourn / size over completed jobs (the reference returns the list of samples).
"""

from __future__ import annotations

import os
import time

import numpy as np

from .. import _lib, dist
from ..structs import QueueResults
from .base import Z99, QsSim

__all__ = ["SizeBasedQsSim", "PerfectSimPredictor", "ExpNoiseSimPredictor", "LognormalNoiseSimPredictor", "replay"]


class PerfectSimPredictor:
    """Ideal predictions: Y = X (size_based.py:33-38)."""

    def predict(self, size: float, _rng) -> float:
        return float(size)


class ExpNoiseSimPredictor:
    """Y | X = x ~ Exp(mean x) (sim/utils/predictor.py:14-24); drawn on the GPU."""


class LognormalNoiseSimPredictor:
    """Y | X = x ~ LogNormal(log x, sigma) (sim/utils/predictor.py:27-43); drawn on the GPU."""

    def __init__(self, sigma: float = 0.5) -> None:
        if sigma <= 0:
            raise ValueError("sigma must be positive")
        self.sigma = float(sigma)


class SizeBasedQsSim(QsSim):
    def __init__(self, num_of_channels: int = 1, discipline: str = "SRPT", buffer=None, verbose=True,
                 buffer_type="list", track_slowdown=False, seed=None, replications=1, p_bins=128, queue_cap=0,
                 device=None):
        if num_of_channels != 1:
            raise ValueError("SizeBasedQsSim currently supports only num_of_channels=1.")  # size_based.py:82
        valid = ("FCFS", "SJF", "PSJF", "SRPT", "SPJF", "PSPJF", "SPRPT")
        if discipline not in valid:
            raise ValueError(f"Unknown discipline {discipline!r}; expected one of {valid}.")
        super().__init__(num_of_channels, buffer, verbose, buffer_type, seed=seed, replications=replications,
                         p_bins=p_bins, device=device)
        self._discipline = discipline
        self._track_slowdown = bool(track_slowdown)
        self.queue_cap = int(queue_cap)
        self.slowdown = 0.0
        self.preemptions = 0
        self._pred_kind, self._pred_sigma = "perfect", 0.5

    def set_predictor(self, predictor) -> None:
        """Accepts the reference's predictor objects too (matched by class name)."""
        name = "PerfectSimPredictor" if predictor is None else type(predictor).__name__
        if name == "PerfectSimPredictor":
            self._pred_kind = "perfect"
        elif name == "ExpNoiseSimPredictor":
            self._pred_kind = "exp"
        elif name == "LognormalNoiseSimPredictor":
            self._pred_kind, self._pred_sigma = "lognormal", float(predictor.sigma)
        else:
            raise NotImplementedError("custom size predictors are Python callbacks; the GPU engine implements the "
                                      "perfect, exponential-noise and log-normal-noise predictors")

    def get_slowdown(self) -> float:
        """Mean of sojourn / size over completed jobs (0 for FCFS, which draws no sizes)."""
        return self.slowdown

    def run(self, total_served: int, replications: int | None = None, keep_replications: bool = False) -> QueueResults:
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
            agg, rec = ctx.sim_size(_lib.SIZE_DISCIPLINES[self._discipline], self.buffer, self._src, self._srv,
                                    total_served, count, rep0, seed, p_bins=self.p_bins, queue_cap=self.queue_cap,
                                    per_replication=keep_replications, predictor=self._pred_kind,
                                    sigma=self._pred_sigma)
            self.timing = ctx.last_timing()
        else:
            agg, rec = np.zeros(_lib.size_agg_len(self.p_bins)), None
        agg = dist.allreduce_sum(agg)
        self._publish_size(agg, rec)
        self.time_spent = time.process_time() - start
        return QueueResults(v=self.get_v(), w=self.get_w(), p=self.get_p(), duration=self.time_spent,
                            utilization=self.calc_load(), replications=R, ci=self.ci)

    def _publish_size(self, agg, rec):
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
        sd, sdci = mean_ci(agg[24:25], agg[25:26])
        p, pci = mean_ci(agg[32:32 + P], agg[32 + P:32 + 2 * P])
        self.w, self.v, self.p = [float(x) for x in w], [float(x) for x in v], [float(x) for x in p]
        self.slowdown = float(sd[0])
        self.ci = {"w": [float(x) for x in wci], "v": [float(x) for x in vci], "p": [float(x) for x in pci],
                   "slowdown": float(sdci[0]), "level": 0.99}
        cast = (lambda x: int(round(x))) if R == 1 else (lambda x: float(x) / R)
        self.served, self.arrived, self.dropped, self.taked, self.preemptions = (cast(agg[i]) for i in range(3, 8))
        self.ttek = float(agg[1] / R)


def replay(discipline, buffer, A, S, total_served, p_bins=128, queue_cap=0, device=None, Y=None):
    """Replay tier: gaps A[i, r] and job sizes S[j, r] in arrival order (FCFS: services in start order), optional
    predicted sizes Y[j, r]; fp64, bit-exact with the reference (csrc/k10_size.cu)."""
    L = _lib
    ctx = L.default_context(device)
    agg, rec = ctx.replay_size(L.SIZE_DISCIPLINES[discipline], buffer, A, S, total_served, p_bins=p_bins,
                               queue_cap=queue_cap, Y=Y)
    flags = rec[L.Z_FLAGS]
    if np.any(flags != 0):
        bad = int(np.flatnonzero(flags)[0])
        raise L.MqError(-int(flags[bad]), f"replication {bad} failed (arrays too short or waiting line overflow)")
    i64 = lambda r: rec[r].astype(np.int64)
    m4 = lambda r: rec[r:r + 4].T.copy()
    return {
        "w": m4(L.Z_WRUN), "v": m4(L.Z_VRUN), "w_sum": m4(L.Z_WSUM), "v_sum": m4(L.Z_VSUM),
        "busy": rec[L.Z_BUSYRUN:L.Z_BUSYRUN + 3].T.copy(), "busy_n": i64(L.Z_BUSYN),
        "slowdown_sum": rec[L.Z_SLOWDOWN:L.Z_SLOWDOWN + 2].T.copy(),
        "p_time": rec[L.Z_P:L.Z_P + p_bins].T.copy(), "p_over": rec[L.Z_POVER].copy(), "ttek": rec[L.Z_TTEK].copy(),
        "arrived": i64(L.Z_ARRIVED), "taked": i64(L.Z_TAKED), "served": i64(L.Z_SERVED), "dropped": i64(L.Z_DROPPED),
        "in_sys": i64(L.Z_INSYS), "queued": i64(L.Z_QUEUED), "preemptions": i64(L.Z_PREEMPTIONS),
        "a_used": i64(L.Z_AUSED), "s_used": i64(L.Z_SUSED), "y_used": i64(L.Z_YUSED), "agg": agg, "timing": ctx.last_timing(),
    }
