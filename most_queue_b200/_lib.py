"""ctypes binding of libmqb200.so (the C ABI in include/mqb200.h).

The library is the only compute path.  If it is missing the import of any simulator fails
loudly with instructions to build it; there is no CPU fallback.
"""

from __future__ import annotations

import ctypes as C
import os
import threading

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
# MQB200_LIB: another build of the same library (A/B measurements of kernel variants, tools/ab_variants.sh)
LIB_PATH = os.environ.get("MQB200_LIB") or os.path.join(HERE, "libmqb200.so")

# distribution kinds (mqb200.h MQ_M ...)
KINDS = {"M": 0, "H": 1, "E": 2, "Gamma": 3, "Pa": 4, "D": 5, "C": 6, "Uniform": 7, "Norm": 8, "Weibull": 9}

# error codes
MQ_E_INVALID, MQ_E_CUDA, MQ_E_NODEVICE, MQ_E_INPUT_SHORT, MQ_E_UNSUPPORTED, MQ_E_OVERFLOW = -1, -2, -3, -4, -5, -6

# record rows / aggregate layout (mqb200.h)
F_W, F_V, F_ARRIVED, F_TAKED, F_SERVED, F_DROPPED, F_TTEK, F_FLAGS, F_POVER, F_P = 0, 4, 8, 9, 10, 11, 12, 13, 14, 16
A_REPS, A_W, A_W2, A_V, A_V2 = 0, 1, 5, 9, 13
A_ARRIVED, A_TAKED, A_SERVED, A_DROPPED, A_TTEK, A_WPOOL, A_VPOOL, A_FLAGGED, A_POVER, A_P = 17, 18, 19, 20, 21, 22, 26, 30, 31, 32
RF_WRUN, RF_VRUN, RF_BUSYRUN, RF_BUSYN, RF_INSYS, RF_QUEUED, RF_AUSED, RF_SUSED, REPLAY_EXTRA = 0, 4, 8, 11, 12, 13, 14, 15, 16
MAX_PBINS = 1024
FIFO_V_AT_START, FIFO_FP64, FIFO_BUSY = 1, 2, 4  # mq_fifo_cfg.flags
FB_BUSY, FB_BUSYN, FIFO_BUSY_ROWS, AGG_BUSY_LEN = 0, 3, 4, 8


def rec_rows(p_bins: int) -> int:
    return 16 + p_bins


def agg_len(p_bins: int) -> int:
    return 32 + 2 * p_bins


def replay_rows(p_bins: int) -> int:
    return rec_rows(p_bins) + REPLAY_EXTRA


class MqDist(C.Structure):
    _fields_ = [("kind", C.c_int32), ("reserved", C.c_int32), ("p", C.c_double * 4)]


class MqFifoCfg(C.Structure):
    _fields_ = [
        ("c", C.c_int32), ("p_bins", C.c_int32), ("buffer", C.c_int64),
        ("src", MqDist), ("srv", MqDist),
        ("total_served", C.c_int64), ("replications", C.c_int64), ("rep0", C.c_int64),
        ("seed", C.c_uint64), ("flags", C.c_int32), ("warmup", C.c_int32),
    ]


class MqJobsCfg(C.Structure):
    _fields_ = [("c", C.c_int32), ("device_ptrs", C.c_int32), ("jobs", C.c_int64), ("replications", C.c_int64),
                ("flags", C.c_int32), ("reserved", C.c_int32)]


JB_W, JB_V, JB_JOBS, JB_TLAST, JB_DMAX, JB_WAITED, JB_ROWS, JB_AGG_LEN = 0, 4, 8, 9, 10, 11, 12, 24


class MqReplayCfg(C.Structure):
    _fields_ = [
        ("c", C.c_int32), ("p_bins", C.c_int32), ("buffer", C.c_int64), ("total_served", C.c_int64),
        ("replications", C.c_int64), ("n_a", C.c_int64), ("n_s", C.c_int64),
        ("device_ptrs", C.c_int32), ("flags", C.c_int32),
    ]


MAX_CLASSES = 8
PRTY = {"NP": 0, "PR": 1, "RS": 2, "RW": 3}


def _prty_code(name):
    if name not in PRTY:
        raise ValueError(f"prty must be one of {sorted(PRTY)}, got {name!r}")
    return PRTY[name]
PG_TTEK, PG_FLAGS, PG_ROWS = 0, 1, 2
PF_WSUM, PF_VSUM, PF_WRUN, PF_VRUN = 0, 4, 8, 12
PF_ARRIVED, PF_TAKED, PF_SERVED, PF_DROPPED, PF_INSYS, PF_QUEUED, PF_TOLD, PF_POVER, PF_AUSED, PF_SUSED, PF_P = (
    16, 17, 18, 19, 20, 21, 22, 23, 24, 25, 26)


def prio_rows(p_bins: int) -> int:
    return 26 + p_bins


def prio_rec_rows(K: int, p_bins: int) -> int:
    return 2 + K * prio_rows(p_bins)


def prio_agg(p_bins: int) -> int:
    return 24 + 2 * p_bins


def prio_agg_len(K: int, p_bins: int) -> int:
    return 4 + K * prio_agg(p_bins)


class MqPrioCfg(C.Structure):
    _fields_ = [
        ("c", C.c_int32), ("K", C.c_int32), ("prty_type", C.c_int32), ("p_bins", C.c_int32),
        ("buffer", C.c_int64), ("total_served", C.c_int64), ("replications", C.c_int64), ("rep0", C.c_int64),
        ("seed", C.c_uint64), ("queue_cap", C.c_int32), ("flags", C.c_int32),
        ("src", MqDist * MAX_CLASSES), ("srv", MqDist * MAX_CLASSES),
    ]


NET_MAX_NODES = 16
NG_TTEK, NG_FLAGS, NG_SERVED, NG_ARRIVED, NG_INSYS, NG_AUSED, NG_UUSED = 0, 1, 2, 3, 4, 5, 6
NG_VSUM, NG_WSUM, NG_VRUN, NG_WRUN, NG_ROWS = 8, 11, 14, 17, 20
NN_VSUM, NN_WSUM, NN_VRUN, NN_WRUN, NN_ARRIVED, NN_TAKED, NN_SERVED, NN_SUSED, NN_INSYS, NN_QUEUED, NN_ROWS = (
    0, 4, 8, 12, 16, 17, 18, 19, 20, 21, 22)


def net_rec_rows(m: int) -> int:
    return NG_ROWS + m * NN_ROWS


def net_agg_len(m: int) -> int:
    return 20 + 16 * m


class MqNetCfg(C.Structure):
    _fields_ = [
        ("m", C.c_int32), ("flags", C.c_int32), ("channels", C.c_int32 * NET_MAX_NODES),
        ("R", C.c_double * ((NET_MAX_NODES + 1) * (NET_MAX_NODES + 1))),
        ("src", MqDist), ("srv", MqDist * NET_MAX_NODES),
        ("job_served", C.c_int64), ("replications", C.c_int64), ("rep0", C.c_int64), ("seed", C.c_uint64),
        ("queue_cap", C.c_int32), ("reserved", C.c_int32),
    ]


S_W, S_V, S_SUMA, S_SUMS, S_TAKED, S_SERVED, S_A, S_B, S_WIN, S_WOUT, S_TAILRAW, S_TAILV, SCAN_LEN = (
    0, 4, 8, 9, 10, 11, 12, 13, 14, 15, 16, 17, 56)
S_LEVELS, SCAN_NLEVEL = 20, 33


class MqScanCfg(C.Structure):
    _fields_ = [
        ("src", MqDist), ("srv", MqDist), ("jobs", C.c_int64), ("job0", C.c_int64), ("seed", C.c_uint64),
        ("chain", C.c_uint32), ("device_ptrs", C.c_int32), ("last_range", C.c_int32), ("flags", C.c_int32),
        ("jobs_ahead", C.c_int64),
    ]


TG_ELAPSED, TG_FLAGS, TG_ARRIVED, TG_LOST, TG_DEPARTURES, TG_SERVED, TG_AREASUM, TG_TEND, TG_T0, TG_AUSED, TG_ROWS = (
    0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 10)
TN_AREA, TN_SUSED, TN_ROWS = 0, 1, 2


def tandem_rec_rows(m: int) -> int:
    return TG_ROWS + m * TN_ROWS


def tandem_agg_len(m: int) -> int:
    return 8 + 2 * m


class MqTandemCfg(C.Structure):
    _fields_ = [
        ("m", C.c_int32), ("flags", C.c_int32), ("capacity", C.c_int64 * NET_MAX_NODES),
        ("src", MqDist), ("srv", MqDist * NET_MAX_NODES),
        ("total_served", C.c_int64), ("warmup_fraction", C.c_double),
        ("replications", C.c_int64), ("rep0", C.c_int64), ("seed", C.c_uint64),
    ]


PN_MAX_NODES, PN_MAX_CLASSES, PN_MAX_SERVERS = 8, 4, 32


class MqPrionetCfg(C.Structure):
    _fields_ = [
        ("m", C.c_int32), ("K", C.c_int32), ("channels", C.c_int32 * PN_MAX_NODES),
        ("prty", C.c_int32 * PN_MAX_NODES), ("nodes_prty", C.c_int32 * (PN_MAX_NODES * PN_MAX_CLASSES)),
        ("R", C.c_double * (PN_MAX_CLASSES * (PN_MAX_NODES + 1) * (PN_MAX_NODES + 1))),
        ("src", MqDist * PN_MAX_CLASSES), ("srv", MqDist * (PN_MAX_NODES * PN_MAX_CLASSES)),
        ("job_served", C.c_int64), ("replications", C.c_int64), ("rep0", C.c_int64), ("seed", C.c_uint64),
        ("queue_cap", C.c_int32), ("flags", C.c_int32),
    ]


# record / aggregate rows of the priority network (include/mqb200.h MQ_PN*)
PNC_VSUM, PNC_WSUM, PNC_VRUN, PNC_WRUN, PNC_SERVED, PNC_ARRIVED, PNC_AUSED, PNC_ROWS = 0, 3, 6, 9, 12, 13, 14, 16
PNN_VSUM, PNN_WSUM, PNN_VRUN, PNN_WRUN, PNN_SERVED, PNN_TAKED, PNN_SUSED, PNN_ROWS = 0, 4, 8, 12, 16, 17, 18, 20
PN_G_TTEK, PN_G_FLAGS, PN_G_UUSED, PN_G_ROWS = 0, 1, 2, 4


def prionet_rec_rows(m, K):
    return PN_G_ROWS + K * PNC_ROWS + m * K * PNN_ROWS


def prionet_agg_len(K):
    return 4 + 16 * K


VAC_STREAMS = ("arrivals", "services", "warm_services", "warm", "cold", "delay")


class MqVacCfg(C.Structure):
    _fields_ = [
        ("c", C.c_int32), ("p_bins", C.c_int32), ("buffer", C.c_int64), ("dist", MqDist * 6), ("set", C.c_int32 * 6),
        ("service_on_warm_up", C.c_int32), ("multiple_vacations", C.c_int32), ("big_n", C.c_int64),
        ("total_served", C.c_int64), ("replications", C.c_int64), ("rep0", C.c_int64), ("seed", C.c_uint64),
        ("queue_cap", C.c_int32), ("flags", C.c_int32),
    ]


# record rows of the vacation kernel (include/mqb200.h MQ_V_*)
(V_TTEK, V_FLAGS, V_ARRIVED, V_TAKED, V_SERVED, V_DROPPED, V_INSYS, V_QUEUED, V_BUSYN, V_POVER, V_WSUM, V_VSUM,
 V_WRUN, V_VRUN, V_BUSYRUN, V_WARMPROB, V_COLDPROB, V_DELAYPROB, V_STARTS, V_WARM_AFTER_COLD, V_USED, V_P) = (
    0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 14, 18, 22, 26, 29, 30, 31, 32, 35, 36, 44)


def vac_rec_rows(p_bins):
    return 44 + int(p_bins)


def vac_agg_len(p_bins):
    return 32 + 2 * int(p_bins)


NEG_STREAMS = ("positives", "negatives", "services", "choices")
NEG_TYPES = {"DISASTER": 1, "RCS": 2}
NEG_SCENARIOS = {"CLEAR_SYSTEM": 1, "REMOVE": 1, "REQUEUE_ALL": 2, "REQUEUE": 2, "RESUME_ALL": 3, "RESUME": 3,
                 "REQUEUE_ALL_NO_RESAMPLING": 4, "REQUEUE_NO_RESAMPLING": 4}


class MqNegCfg(C.Structure):
    _fields_ = [
        ("c", C.c_int32), ("p_bins", C.c_int32), ("buffer", C.c_int64), ("type", C.c_int32), ("queue_cap", C.c_int32),
        ("scenario", C.c_int32), ("reserved", C.c_int32), ("positive", MqDist), ("negative", MqDist), ("service", MqDist),
        ("total_served", C.c_int64), ("replications", C.c_int64), ("rep0", C.c_int64), ("seed", C.c_uint64),
    ]


# record rows of the negative-arrivals kernel (include/mqb200.h MQ_N_*)
(N_TTEK, N_FLAGS, N_POS_ARRIVED, N_NEG_ARRIVED, N_TAKED, N_SERVED, N_BROKEN, N_TOTAL, N_DROPPED, N_INSYS, N_QUEUED,
 N_POVER, N_WSUM, N_VSUM, N_VSSUM, N_VBSUM, N_WRUN, N_VRUN, N_VSRUN, N_VBRUN, N_USED, N_P) = (
    0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11, 12, 16, 20, 24, 28, 32, 36, 40, 44, 48)


def neg_rec_rows(p_bins):
    return 48 + int(p_bins)


def neg_agg_len(p_bins):
    return 48 + 2 * int(p_bins)


SIZE_DISCIPLINES = {"FCFS": 0, "SJF": 1, "PSJF": 2, "SRPT": 3, "SPJF": 4, "PSPJF": 5, "SPRPT": 6}
PRED_KINDS = {"perfect": 0, "exp": 1, "lognormal": 2}


class MqSizeCfg(C.Structure):
    _fields_ = [
        ("discipline", C.c_int32), ("p_bins", C.c_int32), ("buffer", C.c_int64), ("src", MqDist), ("srv", MqDist),
        ("total_served", C.c_int64), ("replications", C.c_int64), ("rep0", C.c_int64), ("seed", C.c_uint64),
        ("queue_cap", C.c_int32), ("pred_kind", C.c_int32), ("pred_sigma", C.c_double),
    ]


# record rows of the size-based kernel (include/mqb200.h MQ_Z_*)
(Z_TTEK, Z_FLAGS, Z_ARRIVED, Z_TAKED, Z_SERVED, Z_DROPPED, Z_INSYS, Z_QUEUED, Z_BUSYN, Z_POVER, Z_PREEMPTIONS, Z_AUSED,
 Z_SUSED, Z_YUSED, Z_SLOWDOWN, Z_WSUM, Z_VSUM, Z_WRUN, Z_VRUN, Z_BUSYRUN, Z_P) = (
    0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11, 12, 13, 14, 16, 20, 24, 28, 32, 36)


def size_rec_rows(p_bins):
    return 36 + int(p_bins)


def size_agg_len(p_bins):
    return 32 + 2 * int(p_bins)


class MqError(RuntimeError):
    def __init__(self, code: int, msg: str):
        super().__init__(f"libmqb200 error {code}: {msg}")
        self.code = code


_lib = None
_lock = threading.Lock()


def load():
    """Load libmqb200.so; raise if it has not been built (no fallback path exists)."""
    global _lib
    with _lock:
        if _lib is not None:
            return _lib
        if not os.path.exists(LIB_PATH):
            raise ImportError(
                f"{LIB_PATH} is missing. Build it with `python -m most_queue_b200.build` "
                "(needs nvcc; the CUDA library is the only compute path, there is no CPU fallback)."
            )
        L = C.CDLL(LIB_PATH)
        dp = C.POINTER(C.c_double)
        vp = C.c_void_p
        L.mq_version.restype = C.c_int
        L.mq_device_count.restype = C.c_int
        L.mq_last_error.restype = C.c_char_p
        L.mq_create.restype = C.c_int
        L.mq_create.argtypes = [C.POINTER(vp), C.c_int]
        L.mq_destroy.restype = C.c_int
        L.mq_destroy.argtypes = [vp]
        L.mq_sim_fifo.restype = C.c_int
        L.mq_sim_fifo.argtypes = [vp, C.POINTER(MqFifoCfg), dp, dp]
        L.mq_replay_fifo.restype = C.c_int
        L.mq_replay_fifo.argtypes = [vp, C.POINTER(MqReplayCfg), vp, vp, dp, vp]
        L.mq_sim_prio.restype = C.c_int
        L.mq_sim_prio.argtypes = [vp, C.POINTER(MqPrioCfg), dp, dp]
        L.mq_replay_prio.restype = C.c_int
        L.mq_replay_prio.argtypes = [vp, C.POINTER(MqPrioCfg), C.POINTER(vp), C.POINTER(C.c_int64),
                                     C.POINTER(vp), C.POINTER(C.c_int64), dp, dp]
        L.mq_sim_net.restype = C.c_int
        L.mq_sim_net.argtypes = [vp, C.POINTER(MqNetCfg), dp, dp]
        L.mq_replay_net.restype = C.c_int
        L.mq_replay_net.argtypes = [vp, C.POINTER(MqNetCfg), vp, C.c_int64, vp, C.c_int64, C.POINTER(vp),
                                    C.POINTER(C.c_int64), dp, dp]
        L.mq_scan_gg1_summary.restype = C.c_int
        L.mq_scan_gg1_summary.argtypes = [vp, C.POINTER(MqScanCfg), vp, vp, dp]
        L.mq_scan_gg1_apply.restype = C.c_int
        L.mq_scan_gg1_apply.argtypes = [vp, C.POINTER(MqScanCfg), vp, vp, C.c_double, dp, vp]
        L.mq_scan_gg1.restype = C.c_int
        L.mq_scan_gg1.argtypes = [vp, C.POINTER(MqScanCfg), vp, vp, dp, vp]
        L.mq_sim_tandem.restype = C.c_int
        L.mq_sim_tandem.argtypes = [vp, C.POINTER(MqTandemCfg), dp, dp]
        L.mq_replay_tandem.restype = C.c_int
        L.mq_replay_tandem.argtypes = [vp, C.POINTER(MqTandemCfg), vp, C.c_int64, C.POINTER(vp),
                                       C.POINTER(C.c_int64), dp, dp]
        L.mq_sim_prionet.restype = C.c_int
        L.mq_sim_prionet.argtypes = [vp, C.POINTER(MqPrionetCfg), dp, dp]
        L.mq_replay_prionet.restype = C.c_int
        L.mq_replay_prionet.argtypes = [vp, C.POINTER(MqPrionetCfg), C.POINTER(vp), C.POINTER(C.c_int64), vp,
                                        C.c_int64, C.POINTER(vp), C.POINTER(C.c_int64), dp, dp]
        L.mq_sim_vacation.restype = C.c_int
        L.mq_sim_vacation.argtypes = [vp, C.POINTER(MqVacCfg), dp, dp]
        L.mq_replay_vacation.restype = C.c_int
        L.mq_replay_vacation.argtypes = [vp, C.POINTER(MqVacCfg), C.POINTER(vp), C.POINTER(C.c_int64), dp, dp]
        L.mq_sim_negatives.restype = C.c_int
        L.mq_sim_negatives.argtypes = [vp, C.POINTER(MqNegCfg), dp, dp]
        L.mq_replay_negatives.restype = C.c_int
        L.mq_replay_negatives.argtypes = [vp, C.POINTER(MqNegCfg), C.POINTER(vp), C.POINTER(C.c_int64), dp, dp]
        L.mq_sim_size.restype = C.c_int
        L.mq_sim_size.argtypes = [vp, C.POINTER(MqSizeCfg), dp, dp]
        L.mq_replay_size.restype = C.c_int
        L.mq_replay_size.argtypes = [vp, C.POINTER(MqSizeCfg), vp, C.c_int64, vp, C.c_int64, vp, C.c_int64, dp, dp]
        L.mq_sim_fifo_multi.restype = C.c_int
        L.mq_sim_fifo_multi.argtypes = [C.POINTER(vp), C.c_int, C.POINTER(MqFifoCfg), dp]
        L.mq_replay_fifo_jobs.restype = C.c_int
        L.mq_replay_fifo_jobs.argtypes = [vp, C.POINTER(MqJobsCfg), vp, vp, dp, vp, vp, vp]
        L.mq_last_timing.restype = C.c_int
        L.mq_last_timing.argtypes = [vp, dp, dp, dp]
        L.mq_last_launches.restype = C.c_int
        L.mq_last_launches.argtypes = [vp]
        L.mq_debug_philox2x32.restype = C.c_int
        L.mq_debug_philox2x32.argtypes = [vp, C.c_uint32, C.c_uint32, C.c_uint32, C.POINTER(C.c_uint32)]
        L.mq_debug_variates.restype = C.c_int
        L.mq_debug_variates.argtypes = [vp, C.POINTER(MqDist), C.c_uint64, C.c_uint32, C.c_int, C.c_int64,
                                        C.POINTER(C.c_float)]
        _lib = L
        return _lib


def check(rc: int):
    if rc != 0:
        msg = load().mq_last_error().decode("utf-8", "replace")
        if rc == MQ_E_INVALID:
            raise ValueError(msg)
        raise MqError(rc, msg)


def _get(params, name):
    if isinstance(params, dict):
        return params[name]
    return getattr(params, name)


def make_dist(kendall_notation: str, params, weibull_ok: bool = False) -> MqDist:
    """(Kendall code, reference-style params) -> mq_dist.

    Unknown codes raise ValueError exactly like create_distribution (random/utils/create.py:76).
    """
    # create_distribution knows no Kendall code for Weibull (create.py:48-77); only the Weibull class itself
    # (random/distributions.py) reaches the sampler
    if kendall_notation not in KINDS or (kendall_notation == "Weibull" and not weibull_ok):
        raise ValueError(
            "Incorrect distribution type specified. Supported: M, H, E, Gamma, C, Pa, Uniform, Norm, D"
        )
    d = MqDist()
    d.kind = KINDS[kendall_notation]
    k = kendall_notation
    if k in ("M", "D"):
        vals = [float(params)]
    elif k in ("H", "C"):
        p1, m1, m2 = _get(params, "p1"), _get(params, "mu1"), _get(params, "mu2")
        if any(isinstance(x, complex) and abs(x.imag) > 0 for x in (p1, m1, m2)):
            raise ValueError("complex H2/Cox parameters cannot be sampled")
        vals = [float(getattr(p1, "real", p1)), float(getattr(m1, "real", m1)), float(getattr(m2, "real", m2))]
    elif k == "E":
        vals = [float(_get(params, "r")), float(_get(params, "mu"))]
    elif k == "Gamma":
        vals = [float(_get(params, "mu")), float(_get(params, "alpha"))]
    elif k == "Pa":
        vals = [float(_get(params, "alpha")), float(_get(params, "K"))]
    elif k == "Uniform":
        vals = [float(_get(params, "mean")), float(_get(params, "half_interval"))]
    elif k == "Norm":
        vals = [float(_get(params, "mean")), float(_get(params, "std_dev"))]
    elif k == "Weibull":
        vals = [float(_get(params, "k")), float(_get(params, "W"))]
    else:  # pragma: no cover
        raise ValueError(k)
    for i, v in enumerate(vals):
        d.p[i] = v
    return d


class Context:
    """One mq_ctx: one GPU, one stream.  Not re-entrant."""

    def __init__(self, device: int | None = None):
        self._L = load()
        if device is None:
            device = int(os.environ.get("LOCAL_RANK", "0"))
            n = self._L.mq_device_count()
            if n > 0:
                device %= n
        self.device = device
        self._h = C.c_void_p()
        check(self._L.mq_create(C.byref(self._h), int(device)))

    def close(self):
        if getattr(self, "_h", None) and self._h.value:
            self._L.mq_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ---- K2 ----
    def sim_fifo(self, c, buffer, src: MqDist, srv: MqDist, total_served, replications, rep0, seed,
                 p_bins=128, per_replication=False, warmup=0, v_at_start=False, fp64=False, busy=False):
        """busy=True: the aggregate has AGG_BUSY_LEN more entries behind agg_len(p_bins) and the record
        FIFO_BUSY_ROWS more rows behind rec_rows(p_bins) (mqb200.h MQ_FIFO_BUSY)."""
        cfg = MqFifoCfg()
        cfg.c, cfg.p_bins = int(c), int(p_bins)
        cfg.buffer = -1 if buffer is None else int(buffer)
        cfg.src, cfg.srv = src, srv
        cfg.total_served, cfg.replications, cfg.rep0 = int(total_served), int(replications), int(rep0)
        cfg.seed = int(seed) & 0xFFFFFFFFFFFFFFFF
        cfg.warmup = int(warmup)
        cfg.flags = (FIFO_V_AT_START if v_at_start else 0) | (FIFO_FP64 if fp64 else 0) | (FIFO_BUSY if busy else 0)
        agg = np.zeros(agg_len(p_bins) + (AGG_BUSY_LEN if busy else 0))
        rows = rec_rows(p_bins) + (FIFO_BUSY_ROWS if busy else 0)
        rec = np.empty((rows, int(replications))) if per_replication else None
        check(self._L.mq_sim_fifo(
            self._h, C.byref(cfg), agg.ctypes.data_as(C.POINTER(C.c_double)),
            rec.ctypes.data_as(C.POINTER(C.c_double)) if rec is not None else None))
        return agg, rec

    def sim_fifo_multi(self, others, c, buffer, src, srv, total_served, replications, rep0, seed, p_bins=128,
                       warmup=0, v_at_start=False, fp64=False, busy=False):
        """mq_sim_fifo_multi: this context and `others` (one per GPU) share the replications of one call."""
        cfg = MqFifoCfg()
        cfg.c, cfg.p_bins = int(c), int(p_bins)
        cfg.buffer = -1 if buffer is None else int(buffer)
        cfg.src, cfg.srv = src, srv
        cfg.total_served, cfg.replications, cfg.rep0 = int(total_served), int(replications), int(rep0)
        cfg.seed = int(seed) & 0xFFFFFFFFFFFFFFFF
        cfg.warmup = int(warmup)
        cfg.flags = (FIFO_V_AT_START if v_at_start else 0) | (FIFO_FP64 if fp64 else 0) | (FIFO_BUSY if busy else 0)
        agg = np.zeros(agg_len(p_bins) + (AGG_BUSY_LEN if busy else 0))
        ctxs = [self] + list(others)
        arr = (C.c_void_p * len(ctxs))(*[x._h for x in ctxs])
        check(self._L.mq_sim_fifo_multi(arr, len(ctxs), C.byref(cfg), agg.ctypes.data_as(C.POINTER(C.c_double))))
        return agg, None

    # ---- K1 ----
    def replay_fifo(self, c, buffer, A: np.ndarray, S: np.ndarray, total_served, p_bins=128, want_agg=True):
        """A: [n_a, R], S: [n_s, R] float64 host arrays (replication-minor)."""
        A = np.ascontiguousarray(A, dtype=np.float64)
        S = np.ascontiguousarray(S, dtype=np.float64)
        if A.ndim != 2 or S.ndim != 2 or A.shape[1] != S.shape[1]:
            raise ValueError("A and S must be 2-D [draw][replication] with equal replication counts")
        R = A.shape[1]
        cfg = MqReplayCfg()
        cfg.c, cfg.p_bins = int(c), int(p_bins)
        cfg.buffer = -1 if buffer is None else int(buffer)
        cfg.total_served, cfg.replications = int(total_served), R
        cfg.n_a, cfg.n_s, cfg.device_ptrs = A.shape[0], S.shape[0], 0
        rec = np.empty((replay_rows(p_bins), R))
        agg = np.zeros(agg_len(p_bins)) if want_agg else None
        check(self._L.mq_replay_fifo(
            self._h, C.byref(cfg), A.ctypes.data, S.ctypes.data,
            agg.ctypes.data_as(C.POINTER(C.c_double)) if want_agg else None, rec.ctypes.data))
        return agg, rec

    def replay_fifo_device(self, c, buffer, a_ptr: int, s_ptr: int, n_a, n_s, replications, total_served,
                           rec_ptr: int, p_bins=128):
        """Device-pointer form (inputs already resident in HBM); rec_ptr: replay_rows x R doubles."""
        cfg = MqReplayCfg()
        cfg.c, cfg.p_bins = int(c), int(p_bins)
        cfg.buffer = -1 if buffer is None else int(buffer)
        cfg.total_served, cfg.replications = int(total_served), int(replications)
        cfg.n_a, cfg.n_s, cfg.device_ptrs = int(n_a), int(n_s), 1
        agg = np.zeros(agg_len(p_bins))
        check(self._L.mq_replay_fifo(self._h, C.byref(cfg), a_ptr, s_ptr,
                                     agg.ctypes.data_as(C.POINTER(C.c_double)), rec_ptr))
        return agg

    # ---- K1b ----
    def replay_fifo_jobs(self, c, A, S, jobs=None, want_samples=False):
        """Job-indexed replay, infinite buffer (mq_replay_fifo_jobs).  A, S: [jobs, R] float64 host arrays.
        Returns (agg, rec, W, V); W, V are None unless want_samples."""
        A = np.ascontiguousarray(A, dtype=np.float64)
        S = np.ascontiguousarray(S, dtype=np.float64)
        if A.ndim != 2 or S.ndim != 2 or A.shape[1] != S.shape[1]:
            raise ValueError("A and S must be 2-D [job][replication] with equal replication counts")
        n = int(min(A.shape[0], S.shape[0]) if jobs is None else jobs)
        if n < 1 or n > A.shape[0] or n > S.shape[0]:
            raise ValueError("jobs must be in 1..rows of A and S")
        R = A.shape[1]
        A, S = np.ascontiguousarray(A[:n]), np.ascontiguousarray(S[:n])
        cfg = MqJobsCfg()
        cfg.c, cfg.device_ptrs, cfg.jobs, cfg.replications = int(c), 0, n, R
        agg, rec = np.zeros(JB_AGG_LEN), np.empty((JB_ROWS, R))
        W = np.empty((n, R)) if want_samples else None
        V = np.empty((n, R)) if want_samples else None
        check(self._L.mq_replay_fifo_jobs(self._h, C.byref(cfg), A.ctypes.data, S.ctypes.data,
                                          agg.ctypes.data_as(C.POINTER(C.c_double)), rec.ctypes.data,
                                          W.ctypes.data if want_samples else None, V.ctypes.data if want_samples else None))
        return agg, rec, W, V

    def replay_fifo_jobs_device(self, c, a_ptr: int, s_ptr: int, jobs, replications, rec_ptr: int = 0):
        """Device-pointer form (arrays resident in HBM); rec_ptr: optional JB_ROWS x R doubles on the device."""
        cfg = MqJobsCfg()
        cfg.c, cfg.device_ptrs, cfg.jobs, cfg.replications = int(c), 1, int(jobs), int(replications)
        agg = np.zeros(JB_AGG_LEN)
        check(self._L.mq_replay_fifo_jobs(self._h, C.byref(cfg), a_ptr, s_ptr, agg.ctypes.data_as(C.POINTER(C.c_double)),
                                          rec_ptr or None, None, None))
        return agg

    # ---- K4 ----
    def _prio_cfg(self, c, K, prty, buffer, total_served, replications, rep0, seed, p_bins, queue_cap):
        if prty not in PRTY:
            raise ValueError(f"prty_type {prty!r} is not supported by the GPU engine (supported: NP, PR, RS, RW)")
        cfg = MqPrioCfg()
        cfg.c, cfg.K, cfg.prty_type, cfg.p_bins = int(c), int(K), PRTY[prty], int(p_bins)
        cfg.buffer = 0 if not buffer else int(buffer)  # priority.py:373: None or 0 -> infinite
        cfg.total_served, cfg.replications, cfg.rep0 = int(total_served), int(replications), int(rep0)
        cfg.seed = int(seed) & 0xFFFFFFFFFFFFFFFF
        cfg.queue_cap = int(queue_cap)
        return cfg

    def sim_prio(self, c, K, prty, buffer, src, srv, total_served, replications, rep0, seed, p_bins=64,
                 queue_cap=0, per_replication=False):
        cfg = self._prio_cfg(c, K, prty, buffer, total_served, replications, rep0, seed, p_bins, queue_cap)
        for k in range(K):
            cfg.src[k], cfg.srv[k] = src[k], srv[k]
        agg = np.zeros(prio_agg_len(K, p_bins))
        rec = np.empty((prio_rec_rows(K, p_bins), int(replications))) if per_replication else None
        check(self._L.mq_sim_prio(self._h, C.byref(cfg), agg.ctypes.data_as(C.POINTER(C.c_double)),
                                  rec.ctypes.data_as(C.POINTER(C.c_double)) if rec is not None else None))
        return agg, rec

    def replay_prio(self, c, K, prty, buffer, A, S, total_served, p_bins=64, queue_cap=0):
        """A[k], S[k]: per-class [n, R] float64 host arrays."""
        A = [np.ascontiguousarray(a, dtype=np.float64) for a in A]
        S = [np.ascontiguousarray(x, dtype=np.float64) for x in S]
        R = A[0].shape[1]
        cfg = self._prio_cfg(c, K, prty, buffer, total_served, R, 0, 0, p_bins, queue_cap)
        pa = (C.c_void_p * K)(*[a.ctypes.data for a in A])
        ps = (C.c_void_p * K)(*[x.ctypes.data for x in S])
        na = (C.c_int64 * K)(*[a.shape[0] for a in A])
        ns = (C.c_int64 * K)(*[x.shape[0] for x in S])
        agg = np.zeros(prio_agg_len(K, p_bins))
        rec = np.empty((prio_rec_rows(K, p_bins), R))
        check(self._L.mq_replay_prio(self._h, C.byref(cfg), pa, na, ps, ns,
                                     agg.ctypes.data_as(C.POINTER(C.c_double)),
                                     rec.ctypes.data_as(C.POINTER(C.c_double))))
        return agg, rec

    # ---- K5 ----
    def _net_cfg(self, R, channels, job_served, replications, rep0, seed, queue_cap):
        m = len(channels)
        Rm = np.asarray(R, dtype=np.float64)
        if Rm.shape != (m + 1, m + 1):
            raise ValueError(f"routing matrix must be {(m + 1)} x {(m + 1)} for {m} nodes")
        if m > NET_MAX_NODES:
            raise MqError(MQ_E_UNSUPPORTED, f"at most {NET_MAX_NODES} nodes")
        cfg = MqNetCfg()
        cfg.m = m
        for i, c in enumerate(channels):
            cfg.channels[i] = int(c)
        flat = Rm.reshape(-1)
        for i in range(flat.size):
            cfg.R[i] = float(flat[i])
        cfg.job_served, cfg.replications, cfg.rep0 = int(job_served), int(replications), int(rep0)
        cfg.seed = int(seed) & 0xFFFFFFFFFFFFFFFF
        cfg.queue_cap = int(queue_cap)
        return cfg

    def sim_net(self, R, channels, src, srv, job_served, replications, rep0, seed, queue_cap=0,
                per_replication=False):
        cfg = self._net_cfg(R, channels, job_served, replications, rep0, seed, queue_cap)
        cfg.src = src
        for i, d in enumerate(srv):
            cfg.srv[i] = d
        m = len(channels)
        agg = np.zeros(net_agg_len(m))
        rec = np.empty((net_rec_rows(m), int(replications))) if per_replication else None
        check(self._L.mq_sim_net(self._h, C.byref(cfg), agg.ctypes.data_as(C.POINTER(C.c_double)),
                                 rec.ctypes.data_as(C.POINTER(C.c_double)) if rec is not None else None))
        return agg, rec

    def replay_net(self, R, channels, A, U, S, job_served, queue_cap=0):
        """A: [n_a, reps] external gaps, U: [n_u, reps] routing uniforms, S[i]: [n, reps] node services."""
        A = np.ascontiguousarray(A, dtype=np.float64)
        U = np.ascontiguousarray(U, dtype=np.float64)
        S = [np.ascontiguousarray(x, dtype=np.float64) for x in S]
        reps = A.shape[1]
        m = len(channels)
        cfg = self._net_cfg(R, channels, job_served, reps, 0, 0, queue_cap)
        ps = (C.c_void_p * m)(*[x.ctypes.data for x in S])
        ns = (C.c_int64 * m)(*[x.shape[0] for x in S])
        agg = np.zeros(net_agg_len(m))
        rec = np.empty((net_rec_rows(m), reps))
        check(self._L.mq_replay_net(self._h, C.byref(cfg), A.ctypes.data, A.shape[0], U.ctypes.data, U.shape[0],
                                    ps, ns, agg.ctypes.data_as(C.POINTER(C.c_double)),
                                    rec.ctypes.data_as(C.POINTER(C.c_double))))
        return agg, rec

    # ---- K3 ----
    def _scan_cfg(self, src, srv, jobs, job0, seed, chain, device_ptrs, last_range, want_p=False, jobs_ahead=0,
                  keep=False):
        cfg = MqScanCfg()
        cfg.flags = (1 if want_p else 0) | (2 if keep else 0)  # MQ_SCAN_WANT_P | MQ_SCAN_KEEP
        cfg.jobs_ahead = int(jobs_ahead)
        if src is not None:
            cfg.src, cfg.srv = src, srv
        cfg.jobs, cfg.job0 = int(jobs), int(job0)
        cfg.seed = int(seed) & 0xFFFFFFFFFFFFFFFF
        cfg.chain, cfg.device_ptrs, cfg.last_range = int(chain), int(device_ptrs), int(last_range)
        return cfg

    def scan_summary(self, jobs, job0=0, src=None, srv=None, seed=0, chain=0, A=None, S=None, device_ptrs=False,
                     keep=False, want_p=False, jobs_ahead=0):
        """(a, b) of the range: W_out = max(a, W_in + b).  A, S: numpy arrays or device pointers (ints).
        keep=True (MQ_SCAN_KEEP): the next scan_apply of the same range on this context walks from what this call left
        on the device (pass the same want_p / jobs_ahead)."""
        cfg = self._scan_cfg(src, srv, jobs, job0, seed, chain, device_ptrs, 0, want_p, jobs_ahead, keep)
        pa, ps, _keep = _scan_ptrs(A, S, jobs, device_ptrs)
        ab = (C.c_double * 2)()
        check(self._L.mq_scan_gg1_summary(self._h, C.byref(cfg), pa, ps, ab))
        return float(ab[0]), float(ab[1])

    def scan_apply(self, jobs, w_in, job0=0, src=None, srv=None, seed=0, chain=0, A=None, S=None,
                   device_ptrs=False, last_range=True, want_w=False, want_p=False, jobs_ahead=0):
        cfg = self._scan_cfg(src, srv, jobs, job0, seed, chain, device_ptrs, last_range, want_p, jobs_ahead)
        pa, ps, _keep = _scan_ptrs(A, S, jobs + (jobs_ahead if want_p else 0), device_ptrs)
        stats = np.zeros(SCAN_LEN)
        W = np.empty(int(jobs)) if want_w else None
        check(self._L.mq_scan_gg1_apply(self._h, C.byref(cfg), pa, ps, float(w_in),
                                        stats.ctypes.data_as(C.POINTER(C.c_double)),
                                        W.ctypes.data if W is not None else None))
        return stats, W

    # ---- K6 ----
    def _tandem_cfg(self, capacity, total_served, warmup_fraction, replications, rep0, seed):
        m = len(capacity)
        if m > NET_MAX_NODES:
            raise MqError(MQ_E_UNSUPPORTED, f"at most {NET_MAX_NODES} nodes")
        cfg = MqTandemCfg()
        cfg.m = m
        for i, k in enumerate(capacity):
            cfg.capacity[i] = -1 if (k is None or k == float("inf")) else int(k)
        cfg.total_served, cfg.warmup_fraction = int(total_served), float(warmup_fraction)
        cfg.replications, cfg.rep0 = int(replications), int(rep0)
        cfg.seed = int(seed) & 0xFFFFFFFFFFFFFFFF
        return cfg

    def sim_tandem(self, capacity, src, srv, total_served, warmup_fraction, replications, rep0, seed,
                   per_replication=False):
        cfg = self._tandem_cfg(capacity, total_served, warmup_fraction, replications, rep0, seed)
        cfg.src = src
        for i, d in enumerate(srv):
            cfg.srv[i] = d
        m = len(capacity)
        agg = np.zeros(tandem_agg_len(m))
        rec = np.empty((tandem_rec_rows(m), int(replications))) if per_replication else None
        check(self._L.mq_sim_tandem(self._h, C.byref(cfg), agg.ctypes.data_as(C.POINTER(C.c_double)),
                                    rec.ctypes.data_as(C.POINTER(C.c_double)) if rec is not None else None))
        return agg, rec

    def replay_tandem(self, capacity, A, S, total_served, warmup_fraction=0.05):
        A = np.ascontiguousarray(A, dtype=np.float64)
        S = [np.ascontiguousarray(x, dtype=np.float64) for x in S]
        reps, m = A.shape[1], len(capacity)
        cfg = self._tandem_cfg(capacity, total_served, warmup_fraction, reps, 0, 0)
        ps = (C.c_void_p * m)(*[x.ctypes.data for x in S])
        ns = (C.c_int64 * m)(*[x.shape[0] for x in S])
        agg = np.zeros(tandem_agg_len(m))
        rec = np.empty((tandem_rec_rows(m), reps))
        check(self._L.mq_replay_tandem(self._h, C.byref(cfg), A.ctypes.data, A.shape[0], ps, ns,
                                       agg.ctypes.data_as(C.POINTER(C.c_double)),
                                       rec.ctypes.data_as(C.POINTER(C.c_double))))
        return agg, rec

    # ---- K7 ----
    def _prionet_cfg(self, R, channels, prty, nodes_prty, job_served, replications, rep0, seed, queue_cap):
        m = len(channels)
        Rm = np.asarray(R, dtype=np.float64)
        if Rm.ndim != 3 or Rm.shape[1:] != (m + 1, m + 1):
            raise ValueError(f"routing matrices must be K x {m + 1} x {m + 1} for {m} nodes")
        K = Rm.shape[0]
        if m > PN_MAX_NODES or K > PN_MAX_CLASSES:
            raise MqError(MQ_E_UNSUPPORTED, f"at most {PN_MAX_NODES} nodes and {PN_MAX_CLASSES} classes")
        cfg = MqPrionetCfg()
        cfg.m, cfg.K = m, K
        for i in range(m):
            cfg.channels[i] = int(channels[i])
            cfg.prty[i] = _prty_code(prty[i])
            for k in range(K):
                cfg.nodes_prty[i * K + k] = int(nodes_prty[i][k])
        flat = Rm.reshape(-1)
        for i in range(flat.size):
            cfg.R[i] = float(flat[i])
        cfg.job_served, cfg.replications, cfg.rep0 = int(job_served), int(replications), int(rep0)
        cfg.seed = int(seed) & 0xFFFFFFFFFFFFFFFF
        cfg.queue_cap = int(queue_cap)
        return cfg

    def sim_prionet(self, R, channels, prty, nodes_prty, src, srv, job_served, replications, rep0, seed,
                    queue_cap=0, per_replication=False):
        """src[k]: MqDist of class k's external flow; srv[i][j]: MqDist of NODE class j at node i."""
        cfg = self._prionet_cfg(R, channels, prty, nodes_prty, job_served, replications, rep0, seed, queue_cap)
        m, K = cfg.m, cfg.K
        for k in range(K):
            cfg.src[k] = src[k]
        for i in range(m):
            for j in range(K):
                cfg.srv[i * K + j] = srv[i][j]
        agg = np.zeros(prionet_agg_len(K))
        rec = np.empty((prionet_rec_rows(m, K), int(replications))) if per_replication else None
        check(self._L.mq_sim_prionet(self._h, C.byref(cfg), agg.ctypes.data_as(C.POINTER(C.c_double)),
                                     rec.ctypes.data_as(C.POINTER(C.c_double)) if rec is not None else None))
        return agg, rec

    def replay_prionet(self, R, channels, prty, nodes_prty, A, U, S, job_served, queue_cap=0):
        """A[k]: [n, reps] gaps of class k, U: [n_u, reps] routing uniforms, S[i][j]: [n, reps] draws of node i /
        node class j (may be empty)."""
        A = [np.ascontiguousarray(a, dtype=np.float64) for a in A]
        U = np.ascontiguousarray(U, dtype=np.float64)
        reps = U.shape[1]
        S = [np.ascontiguousarray(x, dtype=np.float64).reshape(-1, reps) for row in S for x in row]
        cfg = self._prionet_cfg(R, channels, prty, nodes_prty, job_served, reps, 0, 0, queue_cap)
        m, K = cfg.m, cfg.K
        pa = (C.c_void_p * K)(*[a.ctypes.data for a in A])
        na = (C.c_int64 * K)(*[a.shape[0] for a in A])
        ps = (C.c_void_p * (m * K))(*[x.ctypes.data if x.shape[0] else None for x in S])
        ns = (C.c_int64 * (m * K))(*[x.shape[0] for x in S])
        agg = np.zeros(prionet_agg_len(K))
        rec = np.empty((prionet_rec_rows(m, K), reps))
        check(self._L.mq_replay_prionet(self._h, C.byref(cfg), pa, na, U.ctypes.data, U.shape[0], ps, ns,
                                        agg.ctypes.data_as(C.POINTER(C.c_double)),
                                        rec.ctypes.data_as(C.POINTER(C.c_double))))
        return agg, rec

    # ---- K8 ----
    def _vac_cfg(self, c, buffer, is_set, service_on_warm_up, multiple_vacations, big_n, total_served, replications,
                 rep0, seed, p_bins, queue_cap):
        cfg = MqVacCfg()
        cfg.c, cfg.p_bins = int(c), int(p_bins)
        cfg.buffer = -1 if buffer is None else int(buffer)
        for i, k in enumerate(VAC_STREAMS):
            cfg.set[i] = int(bool(is_set.get(k)))
        cfg.service_on_warm_up, cfg.multiple_vacations = int(bool(service_on_warm_up)), int(bool(multiple_vacations))
        cfg.big_n = int(big_n or 0)
        cfg.total_served, cfg.replications, cfg.rep0 = int(total_served), int(replications), int(rep0)
        cfg.seed = int(seed) & 0xFFFFFFFFFFFFFFFF
        cfg.queue_cap = int(queue_cap)
        return cfg

    def sim_vacation(self, c, buffer, dists: dict, total_served, replications, rep0, seed, service_on_warm_up=False,
                     multiple_vacations=False, big_n=0, p_bins=128, queue_cap=0, per_replication=False):
        """dists: MqDist under the names of VAC_STREAMS (absent / None = that law is not set)."""
        cfg = self._vac_cfg(c, buffer, {k: dists.get(k) is not None for k in VAC_STREAMS}, service_on_warm_up,
                            multiple_vacations, big_n, total_served, replications, rep0, seed, p_bins, queue_cap)
        for i, k in enumerate(VAC_STREAMS):
            if dists.get(k) is not None:
                cfg.dist[i] = dists[k]
        agg = np.zeros(vac_agg_len(p_bins))
        rec = np.empty((vac_rec_rows(p_bins), int(replications))) if per_replication else None
        check(self._L.mq_sim_vacation(self._h, C.byref(cfg), agg.ctypes.data_as(C.POINTER(C.c_double)),
                                      rec.ctypes.data_as(C.POINTER(C.c_double)) if rec is not None else None))
        return agg, rec

    def replay_vacation(self, c, buffer, arrays: dict, total_served, service_on_warm_up=False,
                        multiple_vacations=False, big_n=0, p_bins=128, queue_cap=0):
        """arrays: [n, reps] draws under the names of VAC_STREAMS (absent = that law is not set)."""
        X = {k: np.ascontiguousarray(v, dtype=np.float64) for k, v in arrays.items() if v is not None}
        reps = X["arrivals"].shape[1]
        cfg = self._vac_cfg(c, buffer, {k: k in X for k in VAC_STREAMS}, service_on_warm_up, multiple_vacations,
                            big_n, total_served, reps, 0, 0, p_bins, queue_cap)
        ptr = (C.c_void_p * 6)(*[X[k].ctypes.data if k in X else None for k in VAC_STREAMS])
        n = (C.c_int64 * 6)(*[X[k].shape[0] if k in X else 0 for k in VAC_STREAMS])
        agg = np.zeros(vac_agg_len(p_bins))
        rec = np.empty((vac_rec_rows(p_bins), reps))
        check(self._L.mq_replay_vacation(self._h, C.byref(cfg), ptr, n, agg.ctypes.data_as(C.POINTER(C.c_double)),
                                         rec.ctypes.data_as(C.POINTER(C.c_double))))
        return agg, rec

    # ---- K9 ----
    def _neg_cfg(self, c, buffer, neg_type, total_served, replications, rep0, seed, p_bins, queue_cap, scenario=1):
        cfg = MqNegCfg()
        cfg.scenario = int(scenario)
        cfg.c, cfg.p_bins = int(c), int(p_bins)
        cfg.buffer = -1 if buffer is None else int(buffer)
        cfg.type = int(neg_type)
        cfg.queue_cap = int(queue_cap)
        cfg.total_served, cfg.replications, cfg.rep0 = int(total_served), int(replications), int(rep0)
        cfg.seed = int(seed) & 0xFFFFFFFFFFFFFFFF
        return cfg

    def sim_negatives(self, c, buffer, neg_type, positive, negative, service, total_served, replications, rep0, seed,
                      p_bins=128, queue_cap=0, per_replication=False, scenario=1):
        cfg = self._neg_cfg(c, buffer, neg_type, total_served, replications, rep0, seed, p_bins, queue_cap, scenario)
        cfg.positive, cfg.negative, cfg.service = positive, negative, service
        agg = np.zeros(neg_agg_len(p_bins))
        rec = np.empty((neg_rec_rows(p_bins), int(replications))) if per_replication else None
        check(self._L.mq_sim_negatives(self._h, C.byref(cfg), agg.ctypes.data_as(C.POINTER(C.c_double)),
                                       rec.ctypes.data_as(C.POINTER(C.c_double)) if rec is not None else None))
        return agg, rec

    def replay_negatives(self, c, buffer, neg_type, arrays: dict, total_served, p_bins=128, queue_cap=0, scenario=1):
        """arrays: [n, reps] draws under the names of NEG_STREAMS (choices only for RCS)."""
        X = {k: np.ascontiguousarray(v, dtype=np.float64) for k, v in arrays.items() if v is not None}
        reps = X["positives"].shape[1]
        cfg = self._neg_cfg(c, buffer, neg_type, total_served, reps, 0, 0, p_bins, queue_cap, scenario)
        ptr = (C.c_void_p * 4)(*[X[k].ctypes.data if k in X else None for k in NEG_STREAMS])
        n = (C.c_int64 * 4)(*[X[k].shape[0] if k in X else 0 for k in NEG_STREAMS])
        agg = np.zeros(neg_agg_len(p_bins))
        rec = np.empty((neg_rec_rows(p_bins), reps))
        check(self._L.mq_replay_negatives(self._h, C.byref(cfg), ptr, n, agg.ctypes.data_as(C.POINTER(C.c_double)),
                                          rec.ctypes.data_as(C.POINTER(C.c_double))))
        return agg, rec

    # ---- K10 ----
    def _size_cfg(self, discipline, buffer, total_served, replications, rep0, seed, p_bins, queue_cap):
        cfg = MqSizeCfg()
        cfg.discipline, cfg.p_bins = int(discipline), int(p_bins)
        cfg.buffer = -1 if buffer is None else int(buffer)
        cfg.total_served, cfg.replications, cfg.rep0 = int(total_served), int(replications), int(rep0)
        cfg.seed = int(seed) & 0xFFFFFFFFFFFFFFFF
        cfg.queue_cap = int(queue_cap)
        return cfg

    def sim_size(self, discipline, buffer, src, srv, total_served, replications, rep0, seed, p_bins=128, queue_cap=0,
                 per_replication=False, predictor="perfect", sigma=0.5):
        cfg = self._size_cfg(discipline, buffer, total_served, replications, rep0, seed, p_bins, queue_cap)
        cfg.src, cfg.srv = src, srv
        cfg.pred_kind, cfg.pred_sigma = PRED_KINDS[predictor], float(sigma)
        agg = np.zeros(size_agg_len(p_bins))
        rec = np.empty((size_rec_rows(p_bins), int(replications))) if per_replication else None
        check(self._L.mq_sim_size(self._h, C.byref(cfg), agg.ctypes.data_as(C.POINTER(C.c_double)),
                                  rec.ctypes.data_as(C.POINTER(C.c_double)) if rec is not None else None))
        return agg, rec

    def replay_size(self, discipline, buffer, A, S, total_served, p_bins=128, queue_cap=0, Y=None):
        A = np.ascontiguousarray(A, dtype=np.float64)
        S = np.ascontiguousarray(S, dtype=np.float64)
        Y = None if Y is None else np.ascontiguousarray(Y, dtype=np.float64)
        reps = A.shape[1]
        cfg = self._size_cfg(discipline, buffer, total_served, reps, 0, 0, p_bins, queue_cap)
        agg = np.zeros(size_agg_len(p_bins))
        rec = np.empty((size_rec_rows(p_bins), reps))
        check(self._L.mq_replay_size(self._h, C.byref(cfg), A.ctypes.data, A.shape[0], S.ctypes.data, S.shape[0],
                                     Y.ctypes.data if Y is not None else None, Y.shape[0] if Y is not None else 0,
                                     agg.ctypes.data_as(C.POINTER(C.c_double)),
                                     rec.ctypes.data_as(C.POINTER(C.c_double))))
        return agg, rec

    def last_timing(self):
        a, b, c = C.c_double(), C.c_double(), C.c_double()
        check(self._L.mq_last_timing(self._h, C.byref(a), C.byref(b), C.byref(c)))
        return {"sim_ms": a.value, "reduce_ms": b.value, "total_ms": c.value,
                "launches": self._L.mq_last_launches(self._h)}

    # ---- debug / known-answer ----
    def philox2x32(self, c0, c1, key):
        out = (C.c_uint32 * 2)()
        check(self._L.mq_debug_philox2x32(self._h, c0, c1, key, out))
        return int(out[0]), int(out[1])

    def variates(self, dist: MqDist, seed, rep, which, n):
        out = np.empty(int(n), dtype=np.float32)
        check(self._L.mq_debug_variates(self._h, C.byref(dist), int(seed) & 0xFFFFFFFFFFFFFFFF, int(rep),
                                        int(which), int(n), out.ctypes.data_as(C.POINTER(C.c_float))))
        return out


def _scan_ptrs(A, S, jobs, device_ptrs):
    if A is None and S is None:
        return None, None, None
    if device_ptrs:
        return int(A), int(S), None
    A = np.ascontiguousarray(A, dtype=np.float64)
    S = np.ascontiguousarray(S, dtype=np.float64)
    if A.shape[0] < jobs + 1 or S.shape[0] < jobs:
        raise ValueError("replay form needs jobs + 1 gaps and jobs service times")
    return A.ctypes.data, S.ctypes.data, (A, S)


_default_ctx: dict[int, Context] = {}


def default_context(device: int | None = None) -> Context:
    key = -1 if device is None else int(device)
    if key not in _default_ctx:
        _default_ctx[key] = Context(device)
    return _default_ctx[key]


def device_contexts(devices) -> list:
    """Contexts of several GPUs for single-process use: devices = "all" or a list of device indices."""
    if devices == "all":
        devices = list(range(load().mq_device_count()))
    return [default_context(int(d)) for d in devices]


def seed_to_key(seed: int) -> int:
    """64-bit seed -> 32-bit Philox key (splitmix64 finaliser; mirrors mq_api.cu seed_to_key)."""
    m = 0xFFFFFFFFFFFFFFFF
    s = (int(seed) + 0x9E3779B97F4A7C15) & m
    s = ((s ^ (s >> 30)) * 0xBF58476D1CE4E5B9) & m
    s = ((s ^ (s >> 27)) * 0x94D049BB133111EB) & m
    s ^= s >> 31
    return (s ^ (s >> 32)) & 0xFFFFFFFF
