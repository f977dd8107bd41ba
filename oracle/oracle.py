"""ctypes wrapper around the CPU oracle (oracle/libmqoracle.so).

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and the
cpu_baseline / --impl reference legs of bench.py.  The product package
(most_queue_b200) never imports this module.
"""

from __future__ import annotations

import ctypes as C
import os
import subprocess
from dataclasses import dataclass

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "libmqoracle.so")

KINDS = {"M": 0, "H": 1, "E": 2, "Gamma": 3, "Pa": 4, "D": 5, "C": 6, "Uniform": 7, "Norm": 8, "Weibull": 9}


class MqoDist(C.Structure):
    _fields_ = [("kind", C.c_int), ("p", C.c_double * 4)]


class MqoFifoOut(C.Structure):
    _fields_ = [
        ("w_sum", C.c_double * 4),
        ("v_sum", C.c_double * 4),
        ("w_run", C.c_double * 4),
        ("v_run", C.c_double * 4),
        ("busy_run", C.c_double * 3),
        ("busy_n", C.c_int64),
        ("arrived", C.c_int64),
        ("taked", C.c_int64),
        ("served", C.c_int64),
        ("dropped", C.c_int64),
        ("in_sys", C.c_int64),
        ("queued", C.c_int64),
        ("ttek", C.c_double),
        ("p_over", C.c_double),
        ("a_used", C.c_int64),
        ("s_used", C.c_int64),
        ("w_count", C.c_int64),
        ("v_count", C.c_int64),
        ("t_obs", C.c_double),
    ]


class MqoSrc(C.Structure):
    _fields_ = [
        ("generate", C.c_int), ("arr", C.POINTER(C.c_double)), ("n", C.c_int64), ("stride", C.c_int64),
        ("dist", MqoDist), ("seed", C.c_uint32), ("sid", C.c_uint32), ("rep", C.c_uint32),
        ("which", C.c_int), ("pos", C.c_int64),
    ]


class MqoPrioClassOut(C.Structure):
    _fields_ = [
        ("w_run", C.c_double * 4), ("v_run", C.c_double * 4), ("w_sum", C.c_double * 4), ("v_sum", C.c_double * 4),
        ("arrived", C.c_int64), ("taked", C.c_int64), ("served", C.c_int64), ("dropped", C.c_int64),
        ("in_sys", C.c_int64), ("queued", C.c_int64), ("t_old", C.c_double), ("p_over", C.c_double),
    ]


PRTY = {"NP": 0, "PR": 1, "RS": 2, "RW": 3}


class MqoNetNodeOut(C.Structure):
    _fields_ = [("v_run", C.c_double * 4), ("w_run", C.c_double * 4), ("v_sum", C.c_double * 4),
                ("w_sum", C.c_double * 4), ("arrived", C.c_int64), ("taked", C.c_int64), ("served", C.c_int64),
                ("in_sys", C.c_int64), ("queued", C.c_int64)]


class MqoNetOut(C.Structure):
    _fields_ = [("v_run", C.c_double * 3), ("w_run", C.c_double * 3), ("v_sum", C.c_double * 3),
                ("w_sum", C.c_double * 3), ("served", C.c_int64), ("arrived", C.c_int64), ("in_sys", C.c_int64),
                ("ttek", C.c_double)]


class MqoLindleyOut(C.Structure):
    _fields_ = [("w_sum", C.c_double * 4), ("v_sum", C.c_double * 4), ("taked", C.c_int64), ("served", C.c_int64),
                ("sum_a", C.c_double), ("sum_s", C.c_double), ("w_last_raw", C.c_double), ("a", C.c_double),
                ("b", C.c_double)]


class MqoTandemOut(C.Structure):
    _fields_ = [("throughput", C.c_double), ("loss_prob", C.c_double), ("v0", C.c_double),
                ("mean_jobs", C.c_double * 16), ("area", C.c_double * 16), ("t_end", C.c_double),
                ("t0", C.c_double), ("elapsed", C.c_double), ("served", C.c_int64), ("arrived", C.c_int64),
                ("lost", C.c_int64), ("departures", C.c_int64)]


class MqoVacCfg(C.Structure):
    _fields_ = [("c", C.c_int), ("buffer", C.c_int64), ("warm_set", C.c_int), ("cold_set", C.c_int),
                ("delay_set", C.c_int), ("service_on_warm_up", C.c_int), ("multiple_vacations", C.c_int),
                ("big_n", C.c_int64)]


class MqoVacOut(C.Structure):
    _fields_ = [("w_sum", C.c_double * 4), ("v_sum", C.c_double * 4), ("w_run", C.c_double * 4),
                ("v_run", C.c_double * 4), ("busy_run", C.c_double * 3), ("busy_n", C.c_int64),
                ("arrived", C.c_int64), ("taked", C.c_int64), ("served", C.c_int64), ("dropped", C.c_int64),
                ("in_sys", C.c_int64), ("queued", C.c_int64), ("ttek", C.c_double), ("p_over", C.c_double),
                ("warm_prob", C.c_double), ("cold_prob", C.c_double), ("delay_prob", C.c_double),
                ("warm_starts", C.c_int64), ("cold_starts", C.c_int64), ("delay_starts", C.c_int64),
                ("warm_after_cold", C.c_int64)]


class MqoNegOut(C.Structure):
    _fields_ = [(n, C.c_double * 4) for n in ("w_sum", "v_sum", "vs_sum", "vb_sum", "w_run", "v_run", "vs_run", "vb_run")] + [
        (n, C.c_int64) for n in ("positive_arrived", "negative_arrived", "taked", "served", "broken", "total", "dropped",
                                 "in_sys", "queued")] + [("ttek", C.c_double), ("p_over", C.c_double)]


class MqoSizeOut(C.Structure):
    _fields_ = [("w_sum", C.c_double * 4), ("v_sum", C.c_double * 4), ("w_run", C.c_double * 4),
                ("v_run", C.c_double * 4), ("busy_run", C.c_double * 3), ("slowdown_sum", C.c_double * 2)] + [
        (n, C.c_int64) for n in ("busy_n", "arrived", "taked", "served", "dropped", "in_sys", "queued", "preemptions")] + [
        ("ttek", C.c_double), ("p_over", C.c_double)]


class MqoPnClassOut(C.Structure):
    _fields_ = [("v_run", C.c_double * 3), ("w_run", C.c_double * 3), ("v_sum", C.c_double * 3),
                ("w_sum", C.c_double * 3), ("served", C.c_int64), ("arrived", C.c_int64), ("in_sys", C.c_int64)]


class MqoPnNodeClassOut(C.Structure):
    _fields_ = [("v_run", C.c_double * 4), ("w_run", C.c_double * 4), ("v_sum", C.c_double * 4),
                ("w_sum", C.c_double * 4), ("arrived", C.c_int64), ("taked", C.c_int64), ("served", C.c_int64),
                ("in_sys", C.c_int64), ("queued", C.c_int64)]


def build(force: bool = False) -> str:
    """Compile the oracle with the committed Makefile (gcc only)."""
    srcs = [os.path.join(_HERE, f) for f in os.listdir(_HERE) if f.endswith((".c", ".h"))]
    if force or not os.path.exists(_LIB_PATH) or any(
        os.path.getmtime(s) > os.path.getmtime(_LIB_PATH) for s in srcs
    ):
        subprocess.run(["make", "-C", _HERE, "-s", "CC=gcc"], check=True)
    return _LIB_PATH


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(_LIB_PATH)
        dp = C.POINTER(C.c_double)
        L.mqo_fifo_replay.restype = C.c_int
        L.mqo_fifo_replay.argtypes = [
            C.c_int, C.c_int64, dp, C.c_int64, C.c_int64, dp, C.c_int64, C.c_int64, C.c_int64,
            dp, C.c_int64, C.POINTER(MqoFifoOut), dp, C.c_int64, dp, C.c_int64,
        ]
        L.mqo_fifo_generate.restype = C.c_int
        L.mqo_fifo_generate.argtypes = [
            C.c_int, C.c_int64, C.POINTER(MqoDist), C.POINTER(MqoDist), C.c_int64, C.c_uint32,
            C.c_int64, C.c_int64, dp, C.c_int64, C.POINTER(MqoFifoOut), C.c_int,
        ]
        L.mqo_fifo_generate_warm.restype = C.c_int
        L.mqo_fifo_generate_warm.argtypes = [
            C.c_int, C.c_int64, C.POINTER(MqoDist), C.POINTER(MqoDist), C.c_int64, C.c_int64, C.c_uint32,
            C.c_int64, C.c_int64, dp, C.c_int64, C.POINTER(MqoFifoOut), C.c_int,
        ]
        L.mqo_philox2x32_10.restype = None
        L.mqo_philox2x32_10.argtypes = [C.c_uint32, C.c_uint32, C.c_uint32, C.POINTER(C.c_uint32)]
        L.mqo_philox4x32_10.restype = None
        L.mqo_philox4x32_10.argtypes = [C.POINTER(C.c_uint32)] * 3
        L.mqo_variate.restype = C.c_double
        L.mqo_variate.argtypes = [C.POINTER(MqoDist), C.c_uint32, C.c_uint32, C.c_uint32, C.c_int]
        L.mqo_variate_sid.restype = C.c_double
        L.mqo_variate_sid.argtypes = [C.POINTER(MqoDist), C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32, C.c_int]
        L.mqo_prio_run.restype = C.c_int
        L.mqo_prio_run.argtypes = [C.c_int, C.c_int, C.c_int, C.c_int64, C.POINTER(MqoSrc), C.POINTER(MqoSrc),
                                   C.c_int64, dp, C.c_int64, C.POINTER(MqoPrioClassOut), dp]
        L.mqo_net_run.restype = C.c_int
        L.mqo_net_run.argtypes = [C.c_int, C.POINTER(C.c_int), dp, C.POINTER(MqoSrc), C.POINTER(MqoSrc),
                                  C.POINTER(MqoSrc), C.c_int64, C.POINTER(MqoNetOut), C.POINTER(MqoNetNodeOut),
                                  dp, dp, C.c_int64]
        L.mqo_lindley_seq.restype = C.c_int
        L.mqo_lindley_seq.argtypes = [dp, dp, C.c_int64, C.c_double, dp, C.POINTER(MqoLindleyOut)]
        L.mqo_lindley_tree.restype = C.c_int
        L.mqo_lindley_tree.argtypes = [dp, dp, C.c_int64, C.c_double, C.c_int, dp, C.POINTER(MqoLindleyOut)]
        L.mqo_tandem_run.restype = C.c_int
        L.mqo_tandem_run.argtypes = [C.c_int, C.POINTER(C.c_int64), C.POINTER(MqoSrc), C.POINTER(MqoSrc),
                                     C.c_int64, C.c_double, C.POINTER(MqoTandemOut)]
        L.mqo_prionet_run.restype = C.c_int
        L.mqo_prionet_run.argtypes = [C.c_int, C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_int), C.POINTER(C.c_int), dp,
                                      C.POINTER(MqoSrc), C.POINTER(MqoSrc), C.POINTER(MqoSrc), C.c_int64,
                                      C.POINTER(MqoPnClassOut), C.POINTER(MqoPnNodeClassOut), dp]
        L.mqo_version.restype = C.c_char_p
        _lib = L
    return _lib


def _dp(a: np.ndarray):
    return a.ctypes.data_as(C.POINTER(C.c_double))


def make_dist(kind: str, params) -> MqoDist:
    """(kendall code, reference-style params) -> oracle struct.

    Param order follows most_queue/random/utils/params.py: H2Params(mu1, mu2, p1) etc.
    """
    d = MqoDist()
    d.kind = KINDS[kind]
    if kind in ("M", "D"):
        vals = [float(params)]
    elif kind in ("H", "C"):
        vals = [float(params.p1), float(params.mu1), float(params.mu2)]
    elif kind == "E":
        vals = [float(params.r), float(params.mu)]
    elif kind == "Gamma":
        vals = [float(params.mu), float(params.alpha)]
    elif kind == "Pa":
        vals = [float(params.alpha), float(params.K)]
    elif kind == "Uniform":
        vals = [float(params.mean), float(params.half_interval)]
    elif kind == "Norm":
        vals = [float(params.mean), float(params.std_dev)]
    elif kind == "Weibull":
        vals = [float(params.k), float(params.W)]
    else:
        raise ValueError(kind)
    for i, v in enumerate(vals):
        d.p[i] = v
    return d


@dataclass
class FifoResult:
    w_sum: np.ndarray
    v_sum: np.ndarray
    w: np.ndarray  # reference running-mean moments
    v: np.ndarray
    busy: np.ndarray
    busy_n: int
    arrived: int
    taked: int
    served: int
    dropped: int
    in_sys: int
    queued: int
    ttek: float
    p_time: np.ndarray  # time-in-state sums
    p_over: float
    a_used: int
    s_used: int
    w_count: int = 0
    v_count: int = 0
    t_obs: float = 0.0
    w_samples: np.ndarray | None = None
    v_samples: np.ndarray | None = None

    @property
    def p(self) -> np.ndarray:
        return self.p_time / self.ttek


def _unpack(o: MqoFifoOut, p: np.ndarray, ws=None, vs=None) -> FifoResult:
    return FifoResult(
        w_sum=np.array(o.w_sum[:]), v_sum=np.array(o.v_sum[:]),
        w=np.array(o.w_run[:]), v=np.array(o.v_run[:]), busy=np.array(o.busy_run[:]),
        busy_n=o.busy_n, arrived=o.arrived, taked=o.taked, served=o.served, dropped=o.dropped,
        in_sys=o.in_sys, queued=o.queued, ttek=o.ttek, p_time=p, p_over=o.p_over,
        a_used=o.a_used, s_used=o.s_used, w_count=o.w_count, v_count=o.v_count, t_obs=o.t_obs,
        w_samples=ws, v_samples=vs,
    )


def fifo_replay(c: int, buffer, A: np.ndarray, S: np.ndarray, total_served: int,
                p_bins: int = 64, samples: bool = False) -> FifoResult:
    """QsSim.run() on pre-generated interarrival (A) and service (S) arrays."""
    A = np.ascontiguousarray(A, dtype=np.float64)
    S = np.ascontiguousarray(S, dtype=np.float64)
    p = np.zeros(p_bins)
    out = MqoFifoOut()
    ws = np.zeros(len(S)) if samples else None
    vs = np.zeros(int(total_served)) if samples else None
    rc = lib().mqo_fifo_replay(
        c, -1 if buffer is None else int(buffer), _dp(A), len(A), 1, _dp(S), len(S), 1,
        int(total_served), _dp(p), p_bins, C.byref(out),
        _dp(ws) if samples else None, len(ws) if samples else 0,
        _dp(vs) if samples else None, len(vs) if samples else 0,
    )
    if rc != 0:
        raise RuntimeError(f"oracle fifo_replay failed rc={rc} (input arrays too short?)")
    if samples:
        ws = ws[: out.taked]
        vs = vs[: out.served]
    return _unpack(out, p, ws, vs)


def fifo_jobs(c: int, A: np.ndarray, S: np.ndarray, n: int | None = None):
    """Job-indexed replay (infinite buffer): dict with per-job `w`, `v` and the sums of the kernel's association."""
    A = np.ascontiguousarray(A, dtype=np.float64)
    S = np.ascontiguousarray(S, dtype=np.float64)
    n = int(min(len(A), len(S)) if n is None else n)
    w, v = np.zeros(n), np.zeros(n)
    ws, vs = np.zeros(4), np.zeros(4)
    tl, dm, nw = C.c_double(), C.c_double(), C.c_int64()
    L = lib()
    L.mqo_fifo_jobs.restype = C.c_int
    L.mqo_fifo_jobs.argtypes = [C.c_int, C.POINTER(C.c_double), C.c_int64, C.POINTER(C.c_double), C.c_int64, C.c_int64,
                                C.POINTER(C.c_double), C.POINTER(C.c_double), C.POINTER(C.c_double),
                                C.POINTER(C.c_double), C.POINTER(C.c_double), C.POINTER(C.c_double),
                                C.POINTER(C.c_int64)]
    rc = L.mqo_fifo_jobs(c, _dp(A), 1, _dp(S), 1, n, _dp(w), _dp(v), _dp(ws), _dp(vs), C.byref(tl), C.byref(dm), C.byref(nw))
    if rc != 0:
        raise RuntimeError(f"oracle fifo_jobs failed rc={rc}")
    return {"w": w, "v": v, "w_sum": ws, "v_sum": vs, "t_last": tl.value, "d_max": dm.value, "waited": nw.value}


def fifo_generate(c: int, buffer, src, srv, total_served: int, seed: int, rep0: int, reps: int,
                  p_bins: int = 64, threads: int = 1, warmup: int = 0) -> list[FifoResult]:
    """Generator-mode oracle: src/srv are (kendall code, params) tuples."""
    ds, dv = make_dist(*src), make_dist(*srv)
    p = np.zeros((reps, p_bins))
    outs = (MqoFifoOut * reps)()
    rc = lib().mqo_fifo_generate_warm(
        c, -1 if buffer is None else int(buffer), C.byref(ds), C.byref(dv), int(total_served), int(warmup),
        C.c_uint32(seed & 0xFFFFFFFF), int(rep0), int(reps), _dp(p), p_bins, outs, int(threads),
    )
    if rc != 0:
        raise RuntimeError(f"oracle fifo_generate failed rc={rc}")
    return [_unpack(outs[r], p[r]) for r in range(reps)]


def philox2x32(c0: int, c1: int, key: int) -> tuple[int, int]:
    out = (C.c_uint32 * 2)()
    lib().mqo_philox2x32_10(c0, c1, key, out)
    return int(out[0]), int(out[1])


def philox4x32(ctr, key) -> tuple[int, ...]:
    out = (C.c_uint32 * 4)()
    lib().mqo_philox4x32_10((C.c_uint32 * 4)(*ctr), (C.c_uint32 * 2)(*key), out)
    return tuple(int(x) for x in out)


def variate(src_or_srv, seed: int, rep: int, i: int, which: int) -> float:
    d = make_dist(*src_or_srv)
    return float(lib().mqo_variate(C.byref(d), seed & 0xFFFFFFFF, rep, i, which))


# --------------------------------------------------------------------------- priority
def _replay_src(arr: np.ndarray) -> MqoSrc:
    s = MqoSrc()
    s.generate = 0
    s.arr = _dp(arr)
    s.n, s.stride, s.pos = len(arr), 1, 0
    return s


def _gen_src(dist, seed: int, sid: int, rep: int, which: int) -> MqoSrc:
    s = MqoSrc()
    s.generate = 1
    s.dist = make_dist(*dist)
    s.seed, s.sid, s.rep, s.which, s.pos = seed & 0xFFFFFFFF, sid, rep, which, 0
    return s


@dataclass
class PrioResult:
    w: np.ndarray  # [K][4] running-mean moments
    v: np.ndarray
    w_sum: np.ndarray
    v_sum: np.ndarray
    arrived: np.ndarray
    taked: np.ndarray
    served: np.ndarray
    dropped: np.ndarray
    in_sys: np.ndarray
    queued: np.ndarray
    t_old: np.ndarray
    p_time: np.ndarray  # [K][p_bins]
    p_over: np.ndarray
    ttek: float
    a_used: np.ndarray
    s_used: np.ndarray


def _prio_run(c, K, prty, buffer, arrivals, services, total_served, p_bins) -> PrioResult:
    A = (MqoSrc * K)(*arrivals)
    S = (MqoSrc * K)(*services)
    p = np.zeros((K, p_bins))
    outs = (MqoPrioClassOut * K)()
    ttek = C.c_double()
    rc = lib().mqo_prio_run(c, K, PRTY[prty], -1 if buffer is None else int(buffer), A, S, int(total_served),
                            _dp(p), p_bins, outs, C.byref(ttek))
    if rc != 0:
        raise RuntimeError(f"oracle prio run failed rc={rc}")
    f = lambda name: np.array([getattr(outs[k], name) for k in range(K)])
    g = lambda name: np.array([list(getattr(outs[k], name)) for k in range(K)])
    return PrioResult(w=g("w_run"), v=g("v_run"), w_sum=g("w_sum"), v_sum=g("v_sum"), arrived=f("arrived"),
                      taked=f("taked"), served=f("served"), dropped=f("dropped"), in_sys=f("in_sys"),
                      queued=f("queued"), t_old=f("t_old"), p_time=p, p_over=f("p_over"), ttek=ttek.value,
                      a_used=np.array([A[k].pos for k in range(K)]), s_used=np.array([S[k].pos for k in range(K)]))


def prio_replay(c, K, prty, A, S, total_served, buffer=None, p_bins=64) -> PrioResult:
    """PriorityQueueSimulator.run() on per-class pre-generated arrays A[k], S[k]."""
    A = [np.ascontiguousarray(a, dtype=np.float64) for a in A]
    S = [np.ascontiguousarray(x, dtype=np.float64) for x in S]
    return _prio_run(c, K, prty, buffer, [_replay_src(a) for a in A], [_replay_src(x) for x in S],
                     total_served, p_bins)


def prio_generate(c, K, prty, sources, servers, total_served, seed, rep, buffer=None, p_bins=64) -> PrioResult:
    """Generator mode: class k draws gaps from word 0 and services from word 1 of stream sid=k."""
    return _prio_run(c, K, prty, buffer, [_gen_src(sources[k], seed, k, rep, 0) for k in range(K)],
                     [_gen_src(servers[k], seed, k, rep, 1) for k in range(K)], total_served, p_bins)


# --------------------------------------------------------------------------- network
@dataclass
class NetResult:
    v: np.ndarray  # v_network running moments (3)
    w: np.ndarray
    v_sum: np.ndarray
    w_sum: np.ndarray
    served: int
    arrived: int
    in_sys: int
    ttek: float
    node_v: np.ndarray  # [m][4]
    node_w: np.ndarray
    node_v_sum: np.ndarray
    node_w_sum: np.ndarray
    node_served: np.ndarray
    node_taked: np.ndarray
    node_arrived: np.ndarray
    a_used: int
    u_used: int
    s_used: np.ndarray
    v_samples: np.ndarray | None = None
    w_samples: np.ndarray | None = None


def _net_run(R, channels, arrivals, routing, services, job_served, samples) -> NetResult:
    m = len(channels)
    Rm = np.ascontiguousarray(np.asarray(R, dtype=np.float64))
    assert Rm.shape == (m + 1, m + 1)
    ch = (C.c_int * m)(*[int(x) for x in channels])
    S = (MqoSrc * m)(*services)
    out = MqoNetOut()
    nodes = (MqoNetNodeOut * m)()
    vs = np.zeros(int(job_served)) if samples else None
    ws = np.zeros(int(job_served)) if samples else None
    rc = lib().mqo_net_run(m, ch, _dp(Rm), C.byref(arrivals), C.byref(routing), S, int(job_served),
                           C.byref(out), nodes, _dp(vs) if samples else None, _dp(ws) if samples else None,
                           int(job_served) if samples else 0)
    if rc != 0:
        raise RuntimeError(f"oracle net run failed rc={rc}")
    g = lambda name: np.array([list(getattr(nodes[i], name)) for i in range(m)])
    f = lambda name: np.array([getattr(nodes[i], name) for i in range(m)])
    return NetResult(v=np.array(out.v_run[:]), w=np.array(out.w_run[:]), v_sum=np.array(out.v_sum[:]),
                     w_sum=np.array(out.w_sum[:]), served=out.served, arrived=out.arrived, in_sys=out.in_sys,
                     ttek=out.ttek, node_v=g("v_run"), node_w=g("w_run"), node_v_sum=g("v_sum"),
                     node_w_sum=g("w_sum"), node_served=f("served"), node_taked=f("taked"),
                     node_arrived=f("arrived"), a_used=arrivals.pos, u_used=routing.pos,
                     s_used=np.array([S[i].pos for i in range(m)]), v_samples=vs, w_samples=ws)


def net_replay(R, channels, A, U, S, job_served, samples=False) -> NetResult:
    """NetworkSimulator.run() on pre-generated external gaps A, routing uniforms U, node services S[i]."""
    A = np.ascontiguousarray(A, dtype=np.float64)
    U = np.ascontiguousarray(U, dtype=np.float64)
    S = [np.ascontiguousarray(x, dtype=np.float64) for x in S]
    return _net_run(R, channels, _replay_src(A), _replay_src(U), [_replay_src(x) for x in S], job_served, samples)


def net_generate(R, channels, source, servers, job_served, seed, rep) -> NetResult:
    """Generator mode: node i's services = word 1 of stream sid=i; external gaps = word 0 of stream
    sid=m; routing uniforms = word 1 of stream sid=m mapped by ("Uniform", mean .5, half .5)."""
    m = len(channels)

    class U01:
        mean, half_interval = 0.5, 0.5

    return _net_run(R, channels, _gen_src(source, seed, m, rep, 0), _gen_src(("Uniform", U01), seed, m, rep, 1),
                    [_gen_src(servers[i], seed, i, rep, 1) for i in range(m)], job_served, False)


# --------------------------------------------------------------------------- Lindley chain
def lindley(A, S, N=None, w_in=0.0, tree=True, L=16):
    """W[0..N) of the GI/G/1 chain: sequentially (tree=False) or with the scan kernel's association
    tree (tree=True).  Returns (W, MqoLindleyOut)."""
    A = np.ascontiguousarray(A, dtype=np.float64)
    S = np.ascontiguousarray(S, dtype=np.float64)
    N = int(len(S) if N is None else N)
    assert len(A) >= N + 1
    W = np.zeros(N)
    out = MqoLindleyOut()
    if tree:
        rc = lib().mqo_lindley_tree(_dp(A), _dp(S), N, float(w_in), int(L), _dp(W), C.byref(out))
    else:
        rc = lib().mqo_lindley_seq(_dp(A), _dp(S), N, float(w_in), _dp(W), C.byref(out))
    if rc != 0:
        raise RuntimeError(f"oracle lindley failed rc={rc}")
    return W, out


# --------------------------------------------------------------------------- tandem with blocking
def _tandem_run(capacity, arrivals, services, total_served, warmup_fraction):
    m = len(capacity)
    cap = (C.c_int64 * m)(*[-1 if (k is None or k == float("inf")) else int(k) for k in capacity])
    S = (MqoSrc * m)(*services)
    out = MqoTandemOut()
    rc = lib().mqo_tandem_run(m, cap, C.byref(arrivals), S, int(total_served), float(warmup_fraction), C.byref(out))
    if rc != 0:
        raise RuntimeError(f"oracle tandem run failed rc={rc}")
    out.a_used = arrivals.pos
    out.s_used = [S[i].pos for i in range(m)]
    out.m = m
    return out


def tandem_replay(capacity, A, S, total_served, warmup_fraction=0.05):
    """TandemBlockingSim.run() on pre-generated external gaps A and per-node service draws S[i]."""
    A = np.ascontiguousarray(A, dtype=np.float64)
    S = [np.ascontiguousarray(x, dtype=np.float64) for x in S]
    return _tandem_run(capacity, _replay_src(A), [_replay_src(x) for x in S], total_served, warmup_fraction)


def tandem_generate(capacity, source, servers, total_served, seed, rep, warmup_fraction=0.05):
    m = len(capacity)
    return _tandem_run(capacity, _gen_src(source, seed, m, rep, 0),
                       [_gen_src(servers[i], seed, i, rep, 1) for i in range(m)], total_served, warmup_fraction)


# --------------------------------------------------------------------------- priority network
@dataclass
class PrioNetResult:
    v: np.ndarray        # [K][3] v_network running moments
    w: np.ndarray
    v_sum: np.ndarray
    w_sum: np.ndarray
    served: np.ndarray
    arrived: np.ndarray
    ttek: float
    node_v: np.ndarray   # [m][K][4]
    node_w: np.ndarray
    node_served: np.ndarray
    node_taked: np.ndarray
    a_used: np.ndarray
    u_used: int
    s_used: np.ndarray   # [m][K]


def _prionet_run(R, channels, prty, nodes_prty, arrivals, routing, services, job_served) -> PrioNetResult:
    m, K = len(channels), len(arrivals)
    Rm = np.ascontiguousarray(np.asarray(R, dtype=np.float64))
    assert Rm.shape == (K, m + 1, m + 1)
    ch = (C.c_int * m)(*[int(x) for x in channels])
    pt = (C.c_int * m)(*[PRTY[x] for x in prty])
    npr = (C.c_int * (m * K))(*[int(nodes_prty[i][k]) for i in range(m) for k in range(K)])
    A = (MqoSrc * K)(*arrivals)
    S = (MqoSrc * (m * K))(*[services[i][j] for i in range(m) for j in range(K)])
    oc = (MqoPnClassOut * K)()
    on = (MqoPnNodeClassOut * (m * K))()
    ttek = C.c_double()
    rc = lib().mqo_prionet_run(m, K, ch, pt, npr, _dp(Rm), A, C.byref(routing), S, int(job_served), oc, on,
                               C.byref(ttek))
    if rc != 0:
        raise RuntimeError(f"oracle prionet run failed rc={rc}")
    gc = lambda n: np.array([list(getattr(oc[k], n)) for k in range(K)])
    gn = lambda n: np.array([[list(getattr(on[i * K + k], n)) for k in range(K)] for i in range(m)])
    fn = lambda n: np.array([[getattr(on[i * K + k], n) for k in range(K)] for i in range(m)])
    return PrioNetResult(v=gc("v_run"), w=gc("w_run"), v_sum=gc("v_sum"), w_sum=gc("w_sum"),
                         served=np.array([oc[k].served for k in range(K)]),
                         arrived=np.array([oc[k].arrived for k in range(K)]), ttek=ttek.value,
                         node_v=gn("v_run"), node_w=gn("w_run"), node_served=fn("served"), node_taked=fn("taked"),
                         a_used=np.array([A[k].pos for k in range(K)]), u_used=routing.pos,
                         s_used=np.array([[S[i * K + k].pos for k in range(K)] for i in range(m)]))


def prionet_replay(R, channels, prty, nodes_prty, A, U, S, job_served) -> PrioNetResult:
    """PriorityNetwork.run(): A[k] external gaps of class k, U routing uniforms, S[node][node_class] draws."""
    A = [np.ascontiguousarray(a, dtype=np.float64) for a in A]
    U = np.ascontiguousarray(U, dtype=np.float64)
    S = [[np.ascontiguousarray(x, dtype=np.float64) for x in row] for row in S]
    return _prionet_run(R, channels, prty, nodes_prty, [_replay_src(a) for a in A], _replay_src(U),
                        [[_replay_src(x) for x in row] for row in S], job_served)


def prionet_generate(R, channels, prty, nodes_prty, sources, servers, job_served, seed, rep) -> PrioNetResult:
    """Generator mode: stream sid = i*K + j serves node i / node class j (word 1), sid = m*K + k gives class k's
    external gaps (word 0), sid = m*K + K the routing uniforms (word 1).  servers[i][j] is the law of NODE class j."""
    m, K = len(channels), len(sources)

    class U01:
        mean, half_interval = 0.5, 0.5

    return _prionet_run(R, channels, prty, nodes_prty, [_gen_src(sources[k], seed, m * K + k, rep, 0) for k in range(K)],
                        _gen_src(("Uniform", U01), seed, m * K + K, rep, 1),
                        [[_gen_src(servers[i][j], seed, i * K + j, rep, 1) for j in range(K)] for i in range(m)],
                        job_served)


# --------------------------------------------------------------------------- vacations / N-policy
@dataclass
class VacResult:
    w: np.ndarray
    v: np.ndarray
    w_sum: np.ndarray
    v_sum: np.ndarray
    busy: np.ndarray
    p_time: np.ndarray
    ttek: float
    arrived: int
    taked: int
    served: int
    dropped: int
    in_sys: int
    queued: int
    busy_n: int
    p_over: float
    warm_prob: float   # time in completed warm-up periods (QsPhase.prob, not yet divided by ttek)
    cold_prob: float
    delay_prob: float
    starts: tuple      # (warm, cold, delay) starts_times
    warm_after_cold: int
    used: dict         # draws consumed per stream


def _vac_run(c, buffer, srcs, total_served, service_on_warm_up, multiple_vacations, big_n, p_bins):
    """srcs: dict with MqoSrc (or None) under arrivals, services, warm_services, warm, cold, delay."""
    cfg = MqoVacCfg(c=int(c), buffer=-1 if buffer is None else int(buffer),
                    warm_set=int(srcs.get("warm") is not None or srcs.get("warm_services") is not None),
                    cold_set=int(srcs.get("cold") is not None), delay_set=int(srcs.get("delay") is not None),
                    service_on_warm_up=int(bool(service_on_warm_up)), multiple_vacations=int(bool(multiple_vacations)),
                    big_n=int(big_n or 0))
    ref = lambda k: C.byref(srcs[k]) if srcs.get(k) is not None else None
    p = np.zeros(int(p_bins))
    out = MqoVacOut()
    L = lib()
    L.mqo_vac_run.restype = C.c_int
    rc = L.mqo_vac_run(C.byref(cfg), ref("arrivals"), ref("services"), ref("warm_services"), ref("warm"), ref("cold"),
                       ref("delay"), C.c_int64(int(total_served)), _dp(p), C.c_int64(int(p_bins)), C.byref(out))
    if rc != 0:
        raise RuntimeError(f"oracle vacation run failed rc={rc}")
    return VacResult(w=np.array(out.w_run), v=np.array(out.v_run), w_sum=np.array(out.w_sum), v_sum=np.array(out.v_sum),
                     busy=np.array(out.busy_run), p_time=p, ttek=out.ttek, arrived=out.arrived, taked=out.taked,
                     served=out.served, dropped=out.dropped, in_sys=out.in_sys, queued=out.queued, busy_n=out.busy_n,
                     p_over=out.p_over, warm_prob=out.warm_prob, cold_prob=out.cold_prob, delay_prob=out.delay_prob,
                     starts=(out.warm_starts, out.cold_starts, out.delay_starts), warm_after_cold=out.warm_after_cold,
                     used={k: (v.pos if v is not None else 0) for k, v in srcs.items()})


VAC_STREAMS = ("arrivals", "services", "warm_services", "warm", "cold", "delay")


def vac_replay(c, buffer, arrays: dict, total_served, service_on_warm_up=False, multiple_vacations=False, big_n=0,
               p_bins=128) -> VacResult:
    """VacationQueueingSystemSimulator.run() / NPolicyQueueSim.run() on pre-generated draws.  arrays maps the stream
    names of VAC_STREAMS to 1-D arrays (absent = that law is not set)."""
    keep = {k: np.ascontiguousarray(v, dtype=np.float64) for k, v in arrays.items() if v is not None}
    srcs = {k: (_replay_src(keep[k]) if k in keep else None) for k in VAC_STREAMS}
    return _vac_run(c, buffer, srcs, total_served, service_on_warm_up, multiple_vacations, big_n, p_bins)


def vac_generate(c, buffer, dists: dict, total_served, seed, rep, service_on_warm_up=False, multiple_vacations=False,
                 big_n=0, p_bins=128) -> VacResult:
    """Generator mode.  Streams: arrivals = word 0 and services = word 1 of stream sid 0 (as the FIFO kernel),
    warm_services = word 1 of sid 1, warm / cold / delay = word 0 of sid 2 / 3 / 4."""
    where = {"arrivals": (0, 0), "services": (0, 1), "warm_services": (1, 1), "warm": (2, 0), "cold": (3, 0),
             "delay": (4, 0)}
    srcs = {k: (_gen_src(dists[k], seed, where[k][0], rep, where[k][1]) if dists.get(k) is not None else None)
            for k in VAC_STREAMS}
    return _vac_run(c, buffer, srcs, total_served, service_on_warm_up, multiple_vacations, big_n, p_bins)


# --------------------------------------------------------------------------- negative arrivals
NEG_TYPES = {"DISASTER": 1, "RCS": 2}


@dataclass
class NegResult:
    w: np.ndarray
    v: np.ndarray
    v_served: np.ndarray
    v_broken: np.ndarray
    w_sum: np.ndarray
    v_sum: np.ndarray
    vs_sum: np.ndarray
    vb_sum: np.ndarray
    p_time: np.ndarray
    p_over: float
    ttek: float
    counts: dict
    used: dict


NEG_SCENARIOS = {"CLEAR_SYSTEM": 1, "REMOVE": 1, "REQUEUE_ALL": 2, "REQUEUE": 2, "RESUME_ALL": 3, "RESUME": 3,
                 "REQUEUE_ALL_NO_RESAMPLING": 4, "REQUEUE_NO_RESAMPLING": 4}


def _neg_run(c, buffer, neg_type, srcs, total_served, p_bins, scenario=1):
    p = np.zeros(int(p_bins))
    out = MqoNegOut()
    ref = lambda k: C.byref(srcs[k]) if srcs.get(k) is not None else None
    scenario = NEG_SCENARIOS.get(scenario, scenario)
    rc = lib().mqo_neg_run(int(c), C.c_int64(-1 if buffer is None else int(buffer)), NEG_TYPES[neg_type],
                           int(scenario), ref("positives"), ref("negatives"), ref("services"), ref("choices"),
                           C.c_int64(int(total_served)), _dp(p), C.c_int64(int(p_bins)), C.byref(out))
    if rc != 0:
        raise RuntimeError(f"oracle negative-arrivals run failed rc={rc}")
    names = ("positive_arrived", "negative_arrived", "taked", "served", "broken", "total", "dropped", "in_sys", "queued")
    return NegResult(w=np.array(out.w_run), v=np.array(out.v_run), v_served=np.array(out.vs_run),
                     v_broken=np.array(out.vb_run), w_sum=np.array(out.w_sum), v_sum=np.array(out.v_sum),
                     vs_sum=np.array(out.vs_sum), vb_sum=np.array(out.vb_sum), p_time=p, p_over=out.p_over,
                     ttek=out.ttek, counts={n: getattr(out, n) for n in names},
                     used={k: (v.pos if v is not None else 0) for k, v in srcs.items()})


NEG_STREAMS = ("positives", "negatives", "services", "choices")


def neg_replay(c, buffer, neg_type, arrays: dict, total_served, p_bins=128, scenario=1) -> NegResult:
    """QsSimNegatives.run() on pre-generated draws; arrays maps NEG_STREAMS names to 1-D arrays."""
    keep = {k: np.ascontiguousarray(v, dtype=np.float64) for k, v in arrays.items() if v is not None}
    return _neg_run(c, buffer, neg_type, {k: (_replay_src(keep[k]) if k in keep else None) for k in NEG_STREAMS},
                    total_served, p_bins, scenario)


def neg_generate(c, buffer, neg_type, positive, negative, service, total_served, seed, rep, p_bins=128,
                 scenario=1) -> NegResult:
    """Generator mode: positive gaps = word 0 and services = word 1 of stream sid 0 (as the FIFO kernel), negative
    gaps = word 0 and server choices = word 1 of stream sid 1."""

    class U01:
        mean, half_interval = 0.5, 0.5

    srcs = {"positives": _gen_src(positive, seed, 0, rep, 0), "services": _gen_src(service, seed, 0, rep, 1),
            "negatives": _gen_src(negative, seed, 1, rep, 0), "choices": _gen_src(("Uniform", U01), seed, 1, rep, 1)}
    return _neg_run(c, buffer, neg_type, srcs, total_served, p_bins, scenario)


# --------------------------------------------------------------------------- size-based disciplines
SIZE_DISCIPLINES = {"FCFS": 0, "SJF": 1, "PSJF": 2, "SRPT": 3, "SPJF": 4, "PSPJF": 5, "SPRPT": 6}
PRED_KINDS = {"perfect": 0, "exp": 1, "lognormal": 2, "given": 3}


@dataclass
class SizeResult:
    w: np.ndarray
    v: np.ndarray
    w_sum: np.ndarray
    v_sum: np.ndarray
    busy: np.ndarray
    slowdown_sum: np.ndarray
    p_time: np.ndarray
    p_over: float
    ttek: float
    counts: dict
    a_used: int
    s_used: int


def _size_run(discipline, buffer, arrivals, sizes, total_served, p_bins, pred_kind="perfect", pred_sigma=0.0,
              preds=None):
    p = np.zeros(int(p_bins))
    out = MqoSizeOut()
    rc = lib().mqo_size_run(SIZE_DISCIPLINES[discipline], C.c_int64(-1 if buffer is None else int(buffer)),
                            C.byref(arrivals), C.byref(sizes), PRED_KINDS[pred_kind], C.c_double(float(pred_sigma)),
                            C.byref(preds) if preds is not None else None, C.c_int64(int(total_served)), _dp(p),
                            C.c_int64(int(p_bins)), C.byref(out))
    if rc != 0:
        raise RuntimeError(f"oracle size-based run failed rc={rc}")
    names = ("busy_n", "arrived", "taked", "served", "dropped", "in_sys", "queued", "preemptions")
    return SizeResult(w=np.array(out.w_run), v=np.array(out.v_run), w_sum=np.array(out.w_sum),
                      v_sum=np.array(out.v_sum), busy=np.array(out.busy_run), slowdown_sum=np.array(out.slowdown_sum),
                      p_time=p, p_over=out.p_over, ttek=out.ttek, counts={n: getattr(out, n) for n in names},
                      a_used=arrivals.pos, s_used=sizes.pos)


def size_replay(discipline, buffer, A, S, total_served, p_bins=128, Y=None) -> SizeResult:
    """SizeBasedQsSim.run() on pre-generated gaps A and sizes S (arrival order; FCFS: services in start order);
    Y: predicted sizes in arrival order (None = perfect predictor)."""
    A = np.ascontiguousarray(A, dtype=np.float64)
    S = np.ascontiguousarray(S, dtype=np.float64)
    if Y is None:
        return _size_run(discipline, buffer, _replay_src(A), _replay_src(S), total_served, p_bins)
    Y = np.ascontiguousarray(Y, dtype=np.float64)
    return _size_run(discipline, buffer, _replay_src(A), _replay_src(S), total_served, p_bins, "given", 0.0,
                     _replay_src(Y))


def size_generate(discipline, buffer, source, server, total_served, seed, rep, p_bins=128, predictor="perfect",
                  sigma=0.5) -> SizeResult:
    """Generator mode: gaps = word 0 and sizes = word 1 of stream 0 (as the FIFO kernel); the predictor's noise =
    word 0 of stream 1 (Exp(1) for "exp", N(0,1) for "lognormal")."""

    class N01:
        mean, std_dev = 0.0, 1.0

    preds = None
    if predictor == "exp":
        preds = _gen_src(("M", 1.0), seed, 1, rep, 0)
    elif predictor == "lognormal":
        preds = _gen_src(("Norm", N01), seed, 1, rep, 0)
    return _size_run(discipline, buffer, _gen_src(source, seed, 0, rep, 0), _gen_src(server, seed, 0, rep, 1),
                     total_served, p_bins, predictor, sigma, preds)
