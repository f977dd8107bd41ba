"""GI/G/c/r FIFO simulator: drop-in for the reference's QsSim, computed on the GPU.

Mirrors most_queue/sim/base.py: QsSim(num_of_channels, buffer, verbose, buffer_type, seed)
(:29), set_sources (:95), set_servers (:114), run(total_served) (:488) -> QueueResults, and the
attributes callers read afterwards (.w .v .get_p() .dropped .arrived .served .taked .ttek).
Additive: `replications=` (constructor or run()) simulates that many independent replications
in one kernel launch -- one warp lane each -- and returns the mean of the per-replication
estimates (the averaging the reference's scripts do by hand,
misc/run_mgn_with_h2_delay_cold_warm.py:52-87) plus 99 % confidence half-widths.

The event loop itself lives in csrc/k2_fifo_gen.cuh; there is no Python or CPU simulation path.
"""

from __future__ import annotations

import os
import time

import numpy as np

from .. import _lib, dist
from ..constants import DEFAULT_NUM_STATES
from ..random.utils.load import calc_qs_load
from ..structs import QueueResults

Z99 = 2.5758293035489004  # two-sided 99 % normal quantile


class QsSim:
    """Queueing system GI/G/n/r and GI/G/n (FIFO), batched over replications on one B200."""

    def __init__(self, num_of_channels, buffer=None, verbose=True, buffer_type="list", seed=None,
                 replications=1, p_bins=128, device=None, warmup=0, v_at_start=False, precision="f32",
                 busy=False, devices=None):
        if buffer_type not in ("list", "deque"):
            raise ValueError("Unknown queue type")  # base.py:73
        self.n = int(num_of_channels)
        self.buffer = buffer
        self.verbose = verbose
        self.seed = seed
        self.replications = int(replications)
        self.p_bins = int(p_bins)
        self.device = device
        # additive: devices="all" (or a list of indices, or MQ_DEVICES=all in the environment) spreads the replications
        # of ONE run() call over several GPUs from this single process (one host thread per GPU inside the library,
        # mq_sim_fifo_multi); under torchrun every rank keeps its one GPU
        self.devices = devices
        # additive: discard the statistics of the first `warmup` jobs that start service (the reference
        # starts empty and keeps everything, base.py:488-528, which biases every replication by O(1/N))
        self.warmup = int(warmup)
        # additive: sample v for every job that started service before the stop (no right-censoring)
        self.v_at_start = bool(v_at_start)
        # additive: "f32" = the production kernel (fp32 clocks relative to an epoch, MUFU samplers); "f64" = the same
        # loop on the same Philox words in fp64 (MQ_FIFO_FP64), the precision of the reference's Python floats
        if precision not in ("f32", "f64"):
            raise ValueError("precision must be 'f32' or 'f64'")
        self.precision = precision
        # busy-period moments (QsSim.busy, base.py:274-281) cost a few instructions per event: they are gathered when
        # busy=True, or on first access of `.busy` after a run (the run is repeated with the same seed)
        self.want_busy = bool(busy)
        self._busy = None
        self._last_run = None
        self.busy_moments = 0

        self.num_of_states = DEFAULT_NUM_STATES
        self.w = [0, 0, 0, 0]
        self.v = [0, 0, 0, 0]
        self.p = [0.0] * self.p_bins
        self.taked = self.served = self.in_sys = self.arrived = self.dropped = 0
        self.ttek = 0
        self.time_spent = 0
        self.ci = None
        self.agg = None
        self.per_replication = None
        self.timing = None

        self.source_params = self.source_kendall_notation = None
        self.server_params = self.server_kendall_notation = None
        self.is_set_source_params = self.is_set_server_params = False
        self._src = self._srv = None

    # ------------------------------------------------------------------ configuration
    def set_sources(self, params, kendall_notation: str = "M"):
        self._src = _lib.make_dist(kendall_notation, params)
        self.source_params, self.source_kendall_notation = params, kendall_notation
        self.is_set_source_params = True

    def set_servers(self, params, kendall_notation: str = "M"):
        self._srv = _lib.make_dist(kendall_notation, params)
        self.server_params, self.server_kendall_notation = params, kendall_notation
        self.is_set_server_params = True

    def calc_load(self) -> float:
        return calc_qs_load(self.source_kendall_notation, self.source_params,
                            self.server_kendall_notation, self.server_params, self.n)

    # ------------------------------------------------------------------ run
    def run(self, total_served: int, replications: int | None = None, keep_replications: bool = False) -> QueueResults:
        if not (self.is_set_source_params and self.is_set_server_params):
            raise RuntimeError("set_sources() and set_servers() must be called before run()")
        start = time.process_time()
        R = int(self.replications if replications is None else replications)
        if R < 1:
            raise ValueError("replications must be >= 1")
        seed = dist.common_seed(self.seed)  # seed=None: drawn on rank 0 and broadcast
        self._busy = None
        self._last_run = (int(total_served), R, seed)
        agg, rec = self._launch(total_served, R, seed, keep_replications, self.want_busy)
        self._publish(agg, rec)
        self.time_spent = time.process_time() - start
        return QueueResults(v=self.get_v(), w=self.get_w(), p=self.get_p(), duration=self.time_spent,
                            utilization=self.calc_load(), replications=R, ci=self.ci)

    def _launch(self, total_served, R, seed, keep_replications, busy):
        devs = dist.local_devices(self.devices)
        if devs is not None and not keep_replications:
            ctxs = _lib.device_contexts(devs)
            if len(ctxs) > 1:
                agg, _ = ctxs[0].sim_fifo_multi(ctxs[1:], self.n, self.buffer, self._src, self._srv, total_served, R, 0,
                                                seed, p_bins=self.p_bins, warmup=self.warmup, v_at_start=self.v_at_start,
                                                fp64=self.precision == "f64", busy=busy)
                self.timing = ctxs[0].last_timing()
                return agg, None
        rep0, count = dist.shard(R)
        ctx = _lib.default_context(self.device)
        if count > 0:
            agg, rec = ctx.sim_fifo(self.n, self.buffer, self._src, self._srv, total_served, count, rep0, seed,
                                    p_bins=self.p_bins, per_replication=keep_replications, warmup=self.warmup,
                                    v_at_start=self.v_at_start, fp64=self.precision == "f64", busy=busy)
            self.timing = ctx.last_timing()
        else:
            agg, rec = np.zeros(_lib.agg_len(self.p_bins) + (_lib.AGG_BUSY_LEN if busy else 0)), None
        return dist.allreduce_sum(agg), rec

    @property
    def busy(self):
        """Raw moments 1..3 of the busy period (QsSim.busy, base.py:274-281), mean over replications."""
        if self._busy is None and self._last_run is not None:
            n, R, seed = self._last_run
            agg, _ = self._launch(n, R, seed, False, True)
            self._set_busy(agg)
        return self._busy if self._busy is not None else [0, 0, 0]

    def _set_busy(self, agg):
        b = agg[_lib.agg_len(self.p_bins):]
        R = agg[_lib.A_REPS]
        self._busy = [float(x) / R for x in b[0:3]]
        self.busy_moments = int(round(b[3])) if R == 1 else float(b[3]) / R

    def _publish(self, agg: np.ndarray, rec):
        L = _lib
        R = agg[L.A_REPS]
        P = self.p_bins
        self.agg = agg
        self.per_replication = rec
        if len(agg) > L.agg_len(P):
            self._set_busy(agg)
        if agg[L.A_FLAGGED] > 0:
            raise _lib.MqError(_lib.MQ_E_OVERFLOW, f"{int(agg[L.A_FLAGGED])} replications hit a library limit")

        def mean_ci(s, s2):
            m = s / R
            if R > 1:
                var = np.maximum(s2 / R - m * m, 0.0) * R / (R - 1)
                return m, Z99 * np.sqrt(var / R)
            return m, np.full_like(m, np.nan)

        w, wci = mean_ci(agg[L.A_W:L.A_W + 4], agg[L.A_W2:L.A_W2 + 4])
        v, vci = mean_ci(agg[L.A_V:L.A_V + 4], agg[L.A_V2:L.A_V2 + 4])
        p, pci = mean_ci(agg[L.A_P:L.A_P + P], agg[L.A_P + P:L.A_P + 2 * P])
        self.w, self.v = [float(x) for x in w], [float(x) for x in v]
        self.p = [float(x) for x in p]
        self.p_over = float(agg[L.A_POVER] / R)
        self.ci = {"w": [float(x) for x in wci], "v": [float(x) for x in vci], "p": [float(x) for x in pci],
                   "level": 0.99}
        one = R == 1
        cast = (lambda x: int(round(x))) if one else (lambda x: float(x) / R)
        self.arrived, self.taked = cast(agg[L.A_ARRIVED]), cast(agg[L.A_TAKED])
        self.served, self.dropped = cast(agg[L.A_SERVED]), cast(agg[L.A_DROPPED])
        self.ttek = float(agg[L.A_TTEK] / R)
        self.total_served_all = int(round(agg[L.A_SERVED]))

    # ------------------------------------------------------------------ accessors (base.py:566-591)
    def get_p(self) -> list[float]:
        """p[j]: probability of exactly j jobs in the system; padded with zeros to the reference's
        list length (constants.py:6) so that indexing code written for it keeps working."""
        res = list(self.p)
        if len(res) < self.num_of_states:
            res.extend([0.0] * (self.num_of_states - len(res)))
        return res

    def get_v(self) -> list[float]:
        return self.v

    def get_w(self) -> list[float]:
        return self.w

    def __str__(self, is_short=False):
        buffer_str = f"/{self.buffer}" if self.buffer is not None else ""
        res = f"Queueing system {self.source_kendall_notation}/{self.server_kendall_notation}/{self.n}{buffer_str}\n"
        res += f"Load: {self.calc_load():4.3f}\n"
        res += f"Current Time {self.ttek:8.3f}\n"
        res += "Sojourn moments:\n\t" + "\t".join(f"{self.v[i]:8.4f}" for i in range(3)) + "\n"
        res += "Wait moments:\n\t" + "\t".join(f"{self.w[i]:8.4f}" for i in range(3)) + "\n"
        if not is_short:
            res += "Stationary prob:\n\t" + "\t".join(f"{self.p[i]:6.5f}" for i in range(min(10, len(self.p)))) + "\n"
            res += f"Arrived: {self.arrived}\n"
            if self.buffer is not None:
                res += f"Dropped: {self.dropped}\n"
            res += f"Taken: {self.taked}\nServed: {self.served}\n"
        return res


def replay(num_of_channels: int, buffer, A, S, total_served: int, p_bins: int = 128, device=None):
    """Replay tier: run QsSim's event loop on pre-generated variates, fp64, bit-exact.

    A[i, r] is the i-th interarrival draw of replication r (A[0] is the draw set_sources makes,
    base.py:112), S[j, r] the j-th service draw (servers.py:88).  Returns a dict of
    per-replication arrays whose `w` and `v` rows equal the reference's running-mean moments
    bit for bit (csrc/k1_fifo_replay.cuh).
    """
    L = _lib
    ctx = L.default_context(device)
    agg, rec = ctx.replay_fifo(num_of_channels, buffer, A, S, total_served, p_bins=p_bins)
    base = L.rec_rows(p_bins)
    flags = rec[L.F_FLAGS]
    if np.any(flags != 0):
        bad = int(np.flatnonzero(flags)[0])
        code = -int(flags[bad])
        if code == L.MQ_E_INPUT_SHORT:
            raise L.MqError(code, f"replication {bad}: variate arrays ran out before {total_served} departures")
        raise L.MqError(code, f"replication {bad} failed")
    ttek = rec[L.F_TTEK]
    out = {
        "w": rec[base + L.RF_WRUN: base + L.RF_WRUN + 4].T.copy(),
        "v": rec[base + L.RF_VRUN: base + L.RF_VRUN + 4].T.copy(),
        "busy": rec[base + L.RF_BUSYRUN: base + L.RF_BUSYRUN + 3].T.copy(),
        "busy_n": rec[base + L.RF_BUSYN].astype(np.int64),
        "w_sum": rec[L.F_W:L.F_W + 4].T.copy(), "v_sum": rec[L.F_V:L.F_V + 4].T.copy(),
        "p_time": rec[L.F_P:L.F_P + p_bins].T.copy(), "p_over": rec[L.F_POVER].copy(),
        "ttek": ttek.copy(),
        "arrived": rec[L.F_ARRIVED].astype(np.int64), "taked": rec[L.F_TAKED].astype(np.int64),
        "served": rec[L.F_SERVED].astype(np.int64), "dropped": rec[L.F_DROPPED].astype(np.int64),
        "in_sys": rec[base + L.RF_INSYS].astype(np.int64), "queued": rec[base + L.RF_QUEUED].astype(np.int64),
        "a_used": rec[base + L.RF_AUSED].astype(np.int64), "s_used": rec[base + L.RF_SUSED].astype(np.int64),
        "agg": agg, "timing": ctx.last_timing(),
    }
    out["p"] = out["p_time"] / ttek[:, None]
    return out


def replay_jobs(num_of_channels: int, A, S, jobs: int | None = None, want_samples: bool = False, device=None):
    """Job-indexed replay tier for an infinite buffer (csrc/k1b_fifo_jobs.cu): jobs 0..jobs-1 of every replication in
    arrival order.  A[n, r] is the gap before job n of replication r, S[n, r] its service time.  Every per-job wait and
    sojourn time equals the sample QsSim produces for that job bit for bit (base.py:230-236, 269-272, 300-303,
    servers.py:88-90); the moments are over the first `jobs` ARRIVALS (QsSim.run stops at the jobs-th departure).
    Returns a dict: w, v [R, 4] raw moments per replication, w_sum, v_sum, waited, t_last, d_max, and W, V [jobs, R]
    when want_samples."""
    L = _lib
    ctx = L.default_context(device)
    agg, rec, W, V = ctx.replay_fifo_jobs(num_of_channels, A, S, jobs, want_samples)
    n = rec[L.JB_JOBS]
    out = {"w_sum": rec[L.JB_W:L.JB_W + 4].T.copy(), "v_sum": rec[L.JB_V:L.JB_V + 4].T.copy(),
           "w": (rec[L.JB_W:L.JB_W + 4] / n).T.copy(), "v": (rec[L.JB_V:L.JB_V + 4] / n).T.copy(),
           "jobs": n.astype(np.int64), "t_last": rec[L.JB_TLAST].copy(), "d_max": rec[L.JB_DMAX].copy(),
           "waited": rec[L.JB_WAITED].astype(np.int64), "agg": agg, "timing": ctx.last_timing()}
    if want_samples:
        out["W"], out["V"] = W, V
    return out
