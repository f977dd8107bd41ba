"""Generate golden fixtures by running the UNMODIFIED reference in the build container.

Run from the repo root (needs /root/reference, which does not exist on the GPU box):

    python tests/golden/make_golden.py

It imports most_queue from /root/reference (with an in-memory stub for the missing
`colorama` package, which the reference only uses for ANSI colours), feeds QsSim /
PriorityQueueSimulator / NetworkSimulator pre-generated variate arrays through objects
whose .generate() pops from an array (the reference only ever calls .generate():
sim/base.py:112,236, sim/utils/servers.py:88), and records inputs + outputs as .npz
under tests/golden/.  The committed fixtures are what the tests read; this script is
the provenance.
"""

from __future__ import annotations

import contextlib
import io
import json
import os
import sys
import types

import numpy as np

REF = os.environ.get("MQ_REFERENCE", "/root/reference")
OUT = os.path.dirname(os.path.abspath(__file__))


def _install_stub_colorama():
    if "colorama" in sys.modules:
        return
    m = types.ModuleType("colorama")

    class _Blank:
        def __getattr__(self, _):
            return ""

    m.Fore = m.Style = m.Back = _Blank()
    m.init = lambda *a, **k: None
    sys.modules["colorama"] = m


def import_reference():
    _install_stub_colorama()
    if REF not in sys.path:
        sys.path.insert(0, REF)
    import most_queue  # noqa: F401


@contextlib.contextmanager
def quiet():
    with contextlib.redirect_stdout(io.StringIO()), contextlib.redirect_stderr(io.StringIO()):
        yield


class ReplayDist:
    """Pops pre-generated variates; stands in for any most_queue.random distribution."""

    def __init__(self, arr, counter):
        self.arr = arr
        self.counter = counter  # shared [index] so that all servers consume one stream

    def generate(self):
        i = self.counter[0]
        self.counter[0] = i + 1
        return float(self.arr[i])


# --------------------------------------------------------------------------- FIFO
def fifo_case(name, c, buffer, A, S, total_served, nprobs=256):
    from most_queue.sim.base import QsSim

    qs = QsSim(c, buffer=buffer)
    qs.set_sources(1.0, "M")
    qs.set_servers(1.0, "M")
    ia, isv = [0], [0]
    qs.source = ReplayDist(A, ia)
    for s in qs.servers:
        s.dist = ReplayDist(S, isv)
    qs.arrival_time = qs.source.generate()  # the draw set_sources makes (base.py:112)

    ws, vs = [], []
    ow, ov = qs.refresh_w_stat, qs.refresh_v_stat

    def rw(x, count=None):
        ws.append(x)
        ow(x, count)

    def rv(x, count=None):
        vs.append(x)
        ov(x, count)

    qs.refresh_w_stat, qs.refresh_v_stat = rw, rv
    with quiet():
        while qs.served < total_served:  # run() minus tqdm and calc_load (base.py:505-506)
            qs.run_one_step()
    p = np.array(qs.p[:nprobs], dtype=np.float64)  # time-in-state sums (base.py:340)
    np.savez_compressed(
        os.path.join(OUT, f"fifo_{name}.npz"),
        c=c, buffer=-1 if buffer is None else buffer, total_served=total_served,
        A=np.asarray(A), S=np.asarray(S),
        w=np.array(qs.w, dtype=np.float64), v=np.array(qs.v, dtype=np.float64),
        busy=np.array(qs.busy, dtype=np.float64), busy_n=qs.busy_moments,
        p_time=p, ttek=float(qs.ttek), arrived=qs.arrived, taked=qs.taked, served=qs.served,
        dropped=qs.dropped, in_sys=qs.in_sys, queued=qs.queue.size(),
        a_used=ia[0], s_used=isv[0],
        w_samples=np.array(ws, dtype=np.float64), v_samples=np.array(vs, dtype=np.float64),
    )
    print(f"fifo_{name}: taked={qs.taked} served={qs.served} dropped={qs.dropped} w1={qs.w[0]:.6g}")


def make_fifo():
    rng = np.random.default_rng(20260101)
    n = 3000
    # M/M/3 rho=0.667 (README quick start), infinite buffer
    fifo_case("mm3_inf", 3, None, rng.exponential(0.5, n + 200), rng.exponential(1.0, n + 200), n)
    # M/M/3/30 at the load of tests/test_qs_sim.py:46-55 (l=1, mu=1/2.1)
    fifo_case("mm3_r30", 3, 30, rng.exponential(1.0, n + 400), rng.exponential(2.1, n + 400), n)
    # tight buffer with many drops
    fifo_case("mg3_r4", 3, 4, rng.exponential(0.4, 2 * n), rng.gamma(2.0, 0.7, 2 * n), n)
    # pure loss system (tests/test_erlang.py:48)
    fifo_case("mm3_r0", 3, 0, rng.exponential(0.5, 3 * n), rng.exponential(1.2, 3 * n), n)
    # cfg2 shape: M/H2/8 rho=0.7 cv=1.5
    p1, mu1, mu2 = 0.447586, 0.0950716, 0.619214
    br = rng.random(n + 300) < p1
    S = rng.exponential(1.0, n + 300) / np.where(br, mu1, mu2)
    fifo_case("mh2_8_inf", 8, None, rng.exponential(1.0, n + 300), S, n)
    # cfg3 shape: Gamma/Pareto/1
    A = rng.gamma(3.18878, 1.0 / 3.18878, n + 300)
    S = 0.395878 * rng.random(n + 300) ** (-1.0 / 2.30171)
    fifo_case("gpa_1_inf", 1, None, A, S, n)
    # deterministic service with exact ties (D/D-like lattice times)
    A = rng.integers(1, 4, 2 * n).astype(np.float64) * 0.5
    S = np.full(2 * n, 2.0)
    fifo_case("lattice_3_r3", 3, 3, A, S, n)
    # heavy load, long queue (exercises queue growth)
    fifo_case("mm2_heavy", 2, None, rng.exponential(0.52, n + 1200), rng.exponential(1.0, n + 1200), n)
    # tiny run
    fifo_case("tiny", 2, None, rng.exponential(1.0, 40), rng.exponential(1.5, 40), 5)


# --------------------------------------------------------------------------- priority
class LowestFirstSet(set):
    """The reference picks a free server with next(iter(self._free_servers)) (priority.py:355):
    CPython set order, which depends on the removal history.  For pre-emptive disciplines the
    server index decides who is pre-empted (priority.py:401-407), so the fixtures pin that choice
    to "lowest free index" -- the rule the oracle and the kernels use."""

    def __iter__(self):
        return iter(sorted(set.__iter__(self)))


def prio_case(name, c, K, prty, A, S, total_served, nprobs=64, buffer=None):
    from most_queue.sim.priority import PriorityQueueSimulator

    qs = PriorityQueueSimulator(c, K, prty, buffer=buffer)
    qs.set_sources([{"type": "M", "params": 1.0} for _ in range(K)])
    qs.set_servers([{"type": "M", "params": 1.0} for _ in range(K)])
    ia = [[0] for _ in range(K)]
    isv = [[0] for _ in range(K)]
    for k in range(K):
        qs.sources[k] = ReplayDist(A[k], ia[k])
        qs.arrival_time[k] = qs.sources[k].generate()  # the draw set_sources makes (priority.py:130)
    for srv in qs.servers:
        srv.dist = [ReplayDist(S[k], isv[k]) for k in range(K)]
    qs._free_servers = LowestFirstSet(range(c))
    with quiet():
        while sum(qs.served) < total_served:  # run() minus tqdm (priority.py:636-637)
            qs.run_one_step()
    out = dict(c=c, K=K, prty=prty, total_served=total_served, ttek=float(qs.ttek))
    if buffer is not None:
        out["buffer"] = buffer
    for k in range(K):
        out[f"A{k}"] = np.asarray(A[k])
        out[f"S{k}"] = np.asarray(S[k])
    np.savez_compressed(
        os.path.join(OUT, f"prio_{name}.npz"), **out,
        w=np.array(qs.w, dtype=np.float64), v=np.array(qs.v, dtype=np.float64),
        p_time=np.array([qs.p[k][:nprobs] for k in range(K)], dtype=np.float64),
        arrived=np.array(qs.arrived), served=np.array(qs.served), taked=np.array(qs.taked),
        dropped=np.array(qs.dropped), in_sys=np.array(qs.in_sys), t_old=np.array(qs.t_old, dtype=np.float64),
        a_used=np.array([x[0] for x in ia]), s_used=np.array([x[0] for x in isv]),
    )
    print(f"prio_{name}: served={qs.served} v1={[round(float(x[0]), 4) for x in qs.v]}")


def make_prio():
    rng = np.random.default_rng(20260202)
    n = 3000
    # cfg4 shape: c=4, two classes, M arrivals l=[0.3,0.5], Gamma service means [2,4] cv 1.2
    def gam(mean, cv, size):
        shape = 1.0 / cv**2
        return rng.gamma(shape, mean / shape, size)

    for prty in ("NP", "PR", "RS", "RW"):
        A = [rng.exponential(1 / 0.3, n), rng.exponential(1 / 0.5, n)]
        S = [gam(2.0, 1.2, 2 * n), gam(4.0, 1.2, 2 * n)]
        prio_case(f"c4k2_{prty}", 4, 2, prty, A, S, n)
    # three classes, heavier load, single and two servers
    for prty in ("NP", "PR"):
        A = [rng.exponential(4.0, n), rng.exponential(3.0, n), rng.exponential(2.5, n)]
        S = [rng.exponential(1.0, 2 * n), gam(1.5, 0.7, 2 * n), rng.exponential(1.2, 2 * n)]
        prio_case(f"c2k3_{prty}", 2, 3, prty, A, S, n)
        A = [rng.exponential(5.0, n), rng.exponential(4.0, n), rng.exponential(3.0, n)]
        S = [rng.exponential(0.8, 2 * n), rng.exponential(1.0, 2 * n), rng.exponential(0.9, 2 * n)]
        prio_case(f"c1k3_{prty}", 1, 3, prty, A, S, n)
    # finite total buffer (priority.py:437-455: it only bounds the pre-emptive disciplines, when no weaker job can be
    # thrown off a server); its own generator so that the fixtures above keep their streams
    rng = np.random.default_rng(20261020)
    for prty, buf in (("PR", 3), ("RS", 5), ("RW", 2), ("PR", 1)):
        A = [rng.exponential(1.6, 2 * n), rng.exponential(1.1, 2 * n), rng.exponential(1.3, 2 * n)]
        S = [rng.exponential(1.0, 3 * n), gam(1.4, 0.8, 3 * n), rng.exponential(1.6, 3 * n)]
        prio_case(f"c2k3_{prty}_buffer{buf}", 2, 3, prty, A, S, n, buffer=buf)


# --------------------------------------------------------------------------- network
NET_R = [
    [1, 0, 0, 0, 0, 0],
    [0, 0.4, 0.6, 0, 0, 0],
    [0, 0, 0.2, 0.4, 0.4, 0],
    [0, 0, 0, 0, 1, 0],
    [0, 0, 0, 0, 1, 0],
    [0, 0, 0, 0, 0, 1],
]  # tests/test_network_no_prty.py:37-46
NET_CH = [3, 2, 3, 4, 3]


def net_case(name, R, channels, A, U, S, job_served):
    import numpy

    from most_queue.sim.networks.network import NetworkSimulator

    m = len(channels)
    qn = NetworkSimulator()
    qn.set_sources(arrival_rate=1.0, R=np.matrix(R))
    qn.set_nodes(serv_params=[{"type": "M", "params": 1.0} for _ in range(m)], n=channels)
    ia, iu = [0], [0]
    isv = [[0] for _ in range(m)]
    qn.source = ReplayDist(A, ia)
    qn.arrival_time = qn.source.generate()  # the draw set_sources makes (network.py:76)
    for i in range(m):
        for srv in qn.qs[i].servers:
            srv.dist = ReplayDist(S[i], isv[i])
        qn.qs[i]._free_servers = LowestFirstSet(range(channels[i]))
    ureplay = ReplayDist(U, iu)
    vs, ws = [], []
    ov, ow = qn.refresh_v_stat, qn.refresh_w_stat

    def rv(x):
        vs.append(x)
        ov(x)

    def rw(x):
        ws.append(x)
        ow(x)

    qn.refresh_v_stat, qn.refresh_w_stat = rv, rw
    orig = numpy.random.rand
    numpy.random.rand = ureplay.generate  # choose_next_node draws its uniform here (network.py:109)
    try:
        with quiet():
            while qn.served < job_served:  # run() minus tqdm (network.py:211-212)
                qn.run_one_step()
    finally:
        numpy.random.rand = orig
    out = dict(R=np.array(R, dtype=np.float64), channels=np.array(channels), job_served=job_served,
               A=np.asarray(A), U=np.asarray(U), ttek=float(qn.ttek), served=qn.served, arrived=qn.arrived,
               in_sys=qn.in_sys, v=np.array(qn.v_network, dtype=np.float64), w=np.array(qn.w_network, dtype=np.float64),
               v_samples=np.array(vs, dtype=np.float64), w_samples=np.array(ws, dtype=np.float64),
               a_used=ia[0], u_used=iu[0], s_used=np.array([x[0] for x in isv]),
               node_v=np.array([q.v for q in qn.qs], dtype=np.float64),
               node_w=np.array([q.w for q in qn.qs], dtype=np.float64),
               node_served=np.array([q.served for q in qn.qs]), node_taked=np.array([q.taked for q in qn.qs]),
               node_arrived=np.array([q.arrived for q in qn.qs]))
    for i in range(m):
        out[f"S{i}"] = np.asarray(S[i])
    np.savez_compressed(os.path.join(OUT, f"net_{name}.npz"), **out)
    print(f"net_{name}: served={qn.served} arrived={qn.arrived} v1={qn.v_network[0]:.5f} u_used={iu[0]}")


def make_net():
    rng = np.random.default_rng(20260303)
    n = 2500
    # cfg5 shape: test_network_no_prty topology, node service H2-like mixtures, rho = 0.7 per node
    S = []
    for i, c in enumerate(NET_CH):
        mean = 0.7 * c / 1.0
        mix = rng.random(8 * n) < 0.3
        S.append(rng.exponential(1.0, 8 * n) * np.where(mix, 2.2 * mean, 0.486 * mean))
    net_case("cfg5_h2", NET_R, NET_CH, rng.exponential(1.0, 2 * n), rng.random(16 * n), S, n)
    # Jackson sub-case (tests/test_jackson_network.py:29-31): exponential service
    S = [rng.exponential(0.7 * c / 1.3, 8 * n) for c in NET_CH]
    net_case("jackson", NET_R, NET_CH, rng.exponential(1.0, 2 * n), rng.random(16 * n), S, n)
    # two-node tandem with feedback, single servers
    R2 = [[1, 0, 0], [0, 0.8, 0.2], [0.3, 0, 0.7]]
    S = [rng.gamma(2.0, 0.25, 8 * n), rng.gamma(0.8, 0.6, 8 * n)]
    net_case("tandem2_fb", R2, [1, 1], rng.exponential(1.0, 2 * n), rng.random(8 * n), S, n)
    # one node == M/M/3 (tests/test_jackson_network.py::test_single_node_reduces_to_mmn)
    net_case("single", [[1, 0], [0, 1]], [3], rng.exponential(0.5, 2 * n), rng.random(4 * n),
             [rng.exponential(1.0, 3 * n)], n)


# --------------------------------------------------------------------------- priority network
def prionet_case(name, R, channels, prty, nodes_prty, A, U, S, job_served):
    import numpy

    from most_queue.sim.networks.priority_network import PriorityNetwork

    m, K = len(channels), len(A)
    qn = PriorityNetwork(K)
    qn.set_sources([1.0] * K, [np.matrix(R[k]) for k in range(K)])
    qn.set_nodes([[{"type": "M", "params": 1.0} for _ in range(K)] for _ in range(m)], channels, prty, nodes_prty)
    ia = [[0] for _ in range(K)]
    iu = [0]
    isv = [[[0] for _ in range(K)] for _ in range(m)]
    for k in range(K):
        qn.sources[k] = ReplayDist(A[k], ia[k])
        qn.arrival_time[k] = qn.sources[k].generate()  # the draw set_sources makes (:83)
    for i in range(m):
        for srv in qn.qs[i].servers:
            srv.dist = [ReplayDist(S[i][j], isv[i][j]) for j in range(K)]  # indexed by NODE class
        qn.qs[i]._free_servers = LowestFirstSet(range(channels[i]))
    ureplay = ReplayDist(U, iu)
    orig = numpy.random.rand
    numpy.random.rand = ureplay.generate  # choose_next_node (:145)
    try:
        with quiet():
            while sum(qn.served) < job_served:  # run() minus tqdm (:257-258)
                qn.run_one_step()
    finally:
        numpy.random.rand = orig
    out = dict(R=np.array(R, dtype=np.float64), channels=np.array(channels), prty=np.array(prty),
               nodes_prty=np.array(nodes_prty), job_served=job_served, U=np.asarray(U), ttek=float(qn.ttek),
               served=np.array(qn.served), arrived=np.array(qn.arrived),
               v=np.array(qn.v_network, dtype=np.float64), w=np.array(qn.w_network, dtype=np.float64),
               a_used=np.array([x[0] for x in ia]), u_used=iu[0],
               s_used=np.array([[isv[i][j][0] for j in range(K)] for i in range(m)]),
               node_v=np.array([q.v for q in qn.qs], dtype=np.float64),
               node_w=np.array([q.w for q in qn.qs], dtype=np.float64),
               node_served=np.array([q.served for q in qn.qs]), node_taked=np.array([q.taked for q in qn.qs]))
    for k in range(K):
        out[f"A{k}"] = np.asarray(A[k])
    for i in range(m):
        for j in range(K):
            out[f"S{i}_{j}"] = np.asarray(S[i][j])
    np.savez_compressed(os.path.join(OUT, f"prionet_{name}.npz"), **out)
    print(f"prionet_{name}: served={qn.served} v1={[round(float(x[0]), 4) for x in qn.v_network]} u_used={iu[0]}")


def make_prionet():
    rng = np.random.default_rng(20260505)
    n = 2000
    # 3 nodes, 2 classes, class-dependent routing, identity class maps
    R0 = [[1, 0, 0, 0], [0, 0, 0.7, 0.3], [0, 0, 0, 1], [0, 0, 0, 1]]
    R1 = [[0.5, 0.5, 0, 0], [0, 0, 0, 1], [0.2, 0, 0.3, 0.5], [0, 0, 0, 1]]
    ch = [2, 1, 2]
    for prty in ("NP", "PR", "RS", "RW"):
        A = [rng.exponential(2.0, 2 * n), rng.exponential(1.6, 2 * n)]
        S = [[rng.gamma(1.5, 0.5 / 1.5 * (1 + i + j), 6 * n) for j in range(2)] for i in range(3)]
        prionet_case(f"m3k2_{prty}", [R0, R1], ch, [prty] * 3, [[0, 1], [0, 1], [0, 1]], A, rng.random(12 * n), S, n)
    # mixed disciplines and a node that inverts the priorities (tests/test_network_im_prty.py style)
    A = [rng.exponential(2.5, 2 * n), rng.exponential(2.0, 2 * n), rng.exponential(3.0, 2 * n)]
    Rk = [[[1, 0, 0], [0, 0, 1], [0, 0, 1]] if k != 1 else [[0.6, 0.4, 0], [0, 0.5, 0.5], [0.1, 0, 0.9]]
          for k in range(3)]
    S = [[rng.exponential(0.5 + 0.2 * j, 8 * n) for j in range(3)] for i in range(2)]
    prionet_case("m2k3_mixed", Rk, [2, 3], ["PR", "NP"], [[0, 1, 2], [2, 1, 0]], A, rng.random(12 * n), S, n)


# --------------------------------------------------------------------------- vacations / N-policy
def vac_case(name, c, buffer, streams, total_served, service_on_warm_up=False, multiple_vacations=False, big_n=0,
             nprobs=128):
    """streams: dict of arrays under arrivals, services and (optional) warm_services, warm, cold, delay."""
    from most_queue.sim.vacations import NPolicyQueueSim, VacationQueueingSystemSimulator

    kw = dict(buffer=buffer, is_service_on_warm_up=service_on_warm_up, is_multiple_vacations=multiple_vacations)
    qs = NPolicyQueueSim(c, big_n, **kw) if big_n else VacationQueueingSystemSimulator(c, **kw)
    qs.set_sources(1.0, "M")
    qs.set_servers(1.0, "M")
    if "warm" in streams or "warm_services" in streams:
        qs.set_warm(1.0, "M")
    if "cold" in streams:
        qs.set_cold(1.0, "M")
    if "delay" in streams:
        qs.set_cold_delay(1.0, "M")
    used = {k: [0] for k in streams}
    qs.source = ReplayDist(streams["arrivals"], used["arrivals"])
    for srv in qs.servers:
        srv.dist = ReplayDist(streams["services"], used["services"])
        if "warm_services" in streams:
            srv.warm_phase.dist = ReplayDist(streams["warm_services"], used["warm_services"])
    if "warm" in streams:
        qs.warm_phase.dist = ReplayDist(streams["warm"], used["warm"])
    if "cold" in streams:
        qs.cold_phase.dist = ReplayDist(streams["cold"], used["cold"])
    if "delay" in streams:
        qs.cold_delay_phase.dist = ReplayDist(streams["delay"], used["delay"])
    qs._free_servers = LowestFirstSet(range(c))
    qs.arrival_time = qs.source.generate()  # the draw set_sources makes (base.py:112)
    with quiet():
        while qs.served < total_served:  # run() minus tqdm and calc_load
            qs.run_one_step()
    out = dict(c=c, buffer=-1 if buffer is None else buffer, total_served=total_served,
               service_on_warm_up=int(service_on_warm_up), multiple_vacations=int(multiple_vacations), big_n=big_n,
               w=np.array(qs.w, dtype=np.float64), v=np.array(qs.v, dtype=np.float64),
               busy=np.array(qs.busy, dtype=np.float64), busy_n=qs.busy_moments,
               p_time=np.array(qs.p[:nprobs], dtype=np.float64), ttek=float(qs.ttek), arrived=qs.arrived,
               taked=qs.taked, served=qs.served, dropped=qs.dropped, in_sys=qs.in_sys, queued=qs.queue.size(),
               warm_prob=float(qs.warm_phase.prob), cold_prob=float(qs.cold_phase.prob),
               delay_prob=float(qs.cold_delay_phase.prob),
               starts=np.array([qs.warm_phase.starts_times, qs.cold_phase.starts_times,
                                qs.cold_delay_phase.starts_times]),
               warm_after_cold=qs.warm_after_cold_starts)
    for k, a in streams.items():
        out["s_" + k] = np.asarray(a, dtype=np.float64)
        out["used_" + k] = used[k][0]
    np.savez_compressed(os.path.join(OUT, f"vac_{name}.npz"), **out)
    print(f"vac_{name}: served={qs.served} taked={qs.taked} dropped={qs.dropped} w1={qs.w[0]:.5g} "
          f"starts={out['starts'].tolist()} used={ {k: v[0] for k, v in used.items()} }")


def make_vac():
    rng = np.random.default_rng(20260606)
    n = 3000
    ex = lambda mean, m=2 * n: rng.exponential(mean, m)
    # M/G/1 with a warm-up period before service resumes (tests/test_mg1_warmup.py shape, jobs wait during it)
    vac_case("c1_warm", 1, None, dict(arrivals=ex(1.0), services=rng.gamma(0.7, 1.0, 2 * n), warm=ex(1.05)), n)
    # the first job of a busy period is served with the warm law (is_service_on_warm_up)
    vac_case("c1_service_on_warm", 1, None,
             dict(arrivals=ex(1.0), services=ex(0.7), warm_services=rng.gamma(2.0, 0.6, 2 * n)), n,
             service_on_warm_up=True)
    # vacations: a cold period starts whenever the system empties (tests/test_mg1_vacations.py shape)
    vac_case("c1_cold", 1, None, dict(arrivals=ex(1.0), services=ex(0.6), cold=ex(1.5)), n)
    vac_case("c1_cold_multiple", 1, None, dict(arrivals=ex(1.0), services=ex(0.5), cold=ex(0.8)), n,
             multiple_vacations=True)
    # warm-up + cold + cold delay (tests/test_mgn_with_h2_delay_cold_warm.py shape), 3 channels
    vac_case("c3_warm_cold_delay", 3, None,
             dict(arrivals=ex(0.5), services=ex(1.0), warm=ex(0.6), cold=ex(0.9), delay=ex(0.7)), n)
    # finite buffer with vacations: drops while the server is away
    vac_case("c2_cold_buffer4", 2, 4, dict(arrivals=ex(0.6), services=ex(0.9), cold=ex(2.5)), n)
    # N-policy (vacations.py:320-364)
    vac_case("c1_npolicy4", 1, None, dict(arrivals=ex(1.0), services=ex(0.6)), n, big_n=4)
    vac_case("c2_npolicy3_warm", 2, None, dict(arrivals=ex(0.6), services=ex(0.8), warm=ex(0.5)), n, big_n=3)


# --------------------------------------------------------------------------- negative arrivals
class ReplayChoice:
    """Stands in for the `random` module inside most_queue.sim.negative: choice(seq) picks element
    floor(u * len(seq)) with u from a replayed uniform stream (the reference calls random.choice, negative.py:345)."""

    def __init__(self, arr, counter):
        self.arr, self.counter = arr, counter

    def choice(self, seq):
        i = self.counter[0]
        self.counter[0] = i + 1
        return seq[min(int(self.arr[i] * len(seq)), len(seq) - 1)]


def neg_case(name, c, buffer, neg_type, streams, total_served, nprobs=128, scenario=None):
    import most_queue.sim.negative as negmod
    from most_queue.sim.negative import DisasterScenario, NegativeServiceType, QsSimNegatives, RcsScenario

    kw = {}
    if scenario is not None:
        if neg_type == "DISASTER":
            kw["disaster_scenario"] = getattr(DisasterScenario, scenario)
        else:
            kw["rcs_scenario"] = getattr(RcsScenario, scenario)
    qs = QsSimNegatives(c, getattr(NegativeServiceType, neg_type), buffer=buffer, **kw)
    qs.set_positive_sources(1.0, "M")
    qs.set_negative_sources(1.0, "M")
    qs.set_servers(1.0, "M")
    used = {k: [0] for k in streams}
    qs.positive_source = ReplayDist(streams["positives"], used["positives"])
    qs.negative_source = ReplayDist(streams["negatives"], used["negatives"])
    for srv in qs.servers:
        srv.dist = ReplayDist(streams["services"], used["services"])
    qs._free_servers = LowestFirstSet(range(c))
    qs.positive_arrival_time = qs.positive_source.generate()  # negative.py:135
    qs.negative_arrival_time = qs.negative_source.generate()  # negative.py:153
    saved = negmod.random
    if "choices" in streams:
        negmod.random = ReplayChoice(streams["choices"], used["choices"])
    try:
        with quiet():
            while qs.served < total_served:
                qs.run_one_step()
                if not isinstance(qs._free_servers, LowestFirstSet):  # DISASTER rebuilds the set (negative.py:295)
                    qs._free_servers = LowestFirstSet(qs._free_servers)
    finally:
        negmod.random = saved
    out = dict(c=c, buffer=-1 if buffer is None else buffer, neg_type=neg_type, total_served=total_served,
               scenario=scenario or ("CLEAR_SYSTEM" if neg_type == "DISASTER" else "REMOVE"),
               w=np.array(qs.w, dtype=np.float64), v=np.array(qs.v, dtype=np.float64),
               v_served=np.array(qs.v_served, dtype=np.float64), v_broken=np.array(qs.v_broken, dtype=np.float64),
               p_time=np.array(qs.p[:nprobs], dtype=np.float64), ttek=float(qs.ttek),
               positive_arrived=qs.positive_arrived, negative_arrived=qs.negative_arrived, taked=qs.taked,
               served=qs.served, broken=qs.broken, total=qs.total, dropped=qs.dropped, in_sys=qs.in_sys,
               queued=qs.queue.size())
    for k, a in streams.items():
        out["s_" + k] = np.asarray(a, dtype=np.float64)
        out["used_" + k] = used[k][0]
    np.savez_compressed(os.path.join(OUT, f"neg_{name}.npz"), **out)
    print(f"neg_{name}: served={qs.served} broken={qs.broken} total={qs.total} dropped={qs.dropped} "
          f"v1={qs.v[0]:.5g} used={ {k: v[0] for k, v in used.items()} }")


def make_neg():
    rng = np.random.default_rng(20260707)
    n = 3000
    ex = lambda mean, m=3 * n: rng.exponential(mean, m)
    # M/G/1 and M/G/n with disasters (tests/test_mg1_disaster.py, test_mgn_disaster.py shapes)
    neg_case("disaster_c1", 1, None, "DISASTER", dict(positives=ex(1.0), negatives=ex(4.0), services=rng.gamma(0.7, 1.0, 3 * n)), n)
    neg_case("disaster_c3", 3, None, "DISASTER", dict(positives=ex(0.4), negatives=ex(3.0), services=ex(1.0)), n)
    neg_case("disaster_c2_buffer3", 2, 3, "DISASTER", dict(positives=ex(0.4), negatives=ex(5.0), services=ex(0.9)), n)
    # remove the customer in service (tests/test_mg1_rcs.py, test_mgn_rcs.py shapes)
    neg_case("rcs_c1", 1, None, "RCS", dict(positives=ex(1.0), negatives=ex(2.5), services=rng.gamma(0.7, 1.0, 3 * n),
                                            choices=rng.random(3 * n)), n)
    neg_case("rcs_c3", 3, None, "RCS", dict(positives=ex(0.4), negatives=ex(1.5), services=ex(1.0),
                                            choices=rng.random(3 * n)), n)
    neg_case("rcs_c2_buffer2", 2, 2, "RCS", dict(positives=ex(0.4), negatives=ex(2.0), services=ex(0.9),
                                                 choices=rng.random(3 * n)), n)
    # interrupted jobs go back to the head of the line and restart (tests/test_mgn_disaster.py, test_mg1_repeat.py)
    for sc in ("REQUEUE_ALL", "RESUME_ALL", "REQUEUE_ALL_NO_RESAMPLING"):
        neg_case(f"disaster_c3_{sc.lower()}", 3, None, "DISASTER",
                 dict(positives=ex(0.5), negatives=ex(4.0), services=rng.gamma(0.8, 1.1, 4 * n)), n, scenario=sc)
    for sc in ("REQUEUE", "RESUME", "REQUEUE_NO_RESAMPLING"):
        neg_case(f"rcs_c2_{sc.lower()}", 2, None, "RCS",
                 dict(positives=ex(0.7), negatives=ex(3.0), services=rng.gamma(0.8, 0.9, 4 * n),
                      choices=rng.random(3 * n)), n, scenario=sc)


# --------------------------------------------------------------------------- size-based disciplines
class ReplayPredictor:
    """Stands in for a SizePredictor (size_based.py:24-30): returns pre-generated predicted sizes in call order."""

    def __init__(self, arr):
        self.arr, self.i = arr, 0

    def predict(self, size, rng):
        y = float(self.arr[self.i])
        self.i += 1
        return y


def size_case(name, discipline, buffer, A, S, total_served, nprobs=128, Y=None):
    from most_queue.sim.size_based import SizeBasedQsSim

    qs = SizeBasedQsSim(1, discipline, buffer=buffer, track_slowdown=True)
    if Y is not None:
        qs.set_predictor(ReplayPredictor(Y))
    qs.set_sources(1.0, "M")
    qs.set_servers(1.0, "M")
    ia, isv = [0], [0]
    qs.source = ReplayDist(A, ia)
    qs._size_dist = ReplayDist(S, isv)       # sizes at arrival (size_based.py:121-124)
    for srv in qs.servers:
        srv.dist = ReplayDist(S, isv)        # FCFS draws the service at start from the server's law
    qs.arrival_time = qs.source.generate()
    with quiet():
        while qs.served < total_served:
            qs.run_one_step()
    sd = np.array(qs.get_slowdown(), dtype=np.float64)
    np.savez_compressed(
        os.path.join(OUT, f"size_{name}.npz"), discipline=discipline, buffer=-1 if buffer is None else buffer,
        total_served=total_served, A=np.asarray(A), S=np.asarray(S), w=np.array(qs.w, dtype=np.float64),
        v=np.array(qs.v, dtype=np.float64), busy=np.array(qs.busy, dtype=np.float64), busy_n=qs.busy_moments,
        p_time=np.array(qs.p[:nprobs], dtype=np.float64), ttek=float(qs.ttek), arrived=qs.arrived, taked=qs.taked,
        served=qs.served, dropped=qs.dropped, in_sys=qs.in_sys, queued=qs.queue.size(), a_used=ia[0], s_used=isv[0],
        slowdown=sd, **({"Y": np.asarray(Y), "y_used": qs._predictor.i} if Y is not None else {}))
    print(f"size_{name}: served={qs.served} taked={qs.taked} dropped={qs.dropped} v1={qs.v[0]:.5g} "
          f"slowdown={sd.mean() if len(sd) else 0:.4g} used={ia[0]},{isv[0]}")


def make_size():
    rng = np.random.default_rng(20260808)
    n = 3000
    for d in ("FCFS", "SJF", "PSJF", "SRPT", "SPJF", "SPRPT"):
        size_case(d.lower(), d, None, rng.exponential(1.0, 2 * n), rng.gamma(0.5, 1.6, 2 * n), n)
    size_case("srpt_buffer5", "SRPT", 5, rng.exponential(1.0, 3 * n), rng.gamma(0.5, 1.9, 3 * n), n)
    size_case("sjf_heavy", "SJF", None, rng.exponential(1.0, 2 * n), 0.3 * (rng.pareto(2.2, 2 * n) + 1.0), n)
    # noisy predictions (sim/utils/predictor.py): Y = X * exp(sigma Z) and Y = X * Exp(1), replayed
    for d in ("SPJF", "PSPJF", "SPRPT"):
        S = rng.gamma(0.5, 1.6, 2 * n)
        size_case(f"{d.lower()}_lognormal", d, None, rng.exponential(1.0, 2 * n), S, n,
                  Y=S * np.exp(0.7 * rng.standard_normal(2 * n)))
    S = rng.gamma(0.5, 1.6, 3 * n)
    size_case("sprpt_expnoise_buffer6", "SPRPT", 6, rng.exponential(1.0, 3 * n), S, n, Y=S * rng.exponential(1.0, 3 * n))


# --------------------------------------------------------------------------- tandem with blocking
def tandem_case(name, capacity, A, S, total_served, warmup_fraction=0.05):
    from most_queue.sim.networks.tandem_blocking import TandemBlockingSim

    m = len(capacity)
    sim = TandemBlockingSim(seed=1)
    sim.set_sources(1.0)
    sim.set_nodes([{"type": "M", "params": 1.0} for _ in range(m)], capacity)
    ia = [0]
    isv = [[0] for _ in range(m)]
    sim.source = ReplayDist(A, ia)
    sim.servers = [ReplayDist(S[i], isv[i]) for i in range(m)]
    res = sim.run(total_served, warmup_fraction=warmup_fraction)
    out = dict(capacity=np.array([-1 if k is None else k for k in capacity]), total_served=total_served,
               warmup_fraction=warmup_fraction, A=np.asarray(A), throughput=float(sim.throughput),
               loss_prob=float(sim.loss_prob), mean_jobs=np.array(res.mean_jobs, dtype=np.float64),
               v0=float(res.v[0]), served=int(res.served), a_used=ia[0], s_used=np.array([x[0] for x in isv]))
    for i in range(m):
        out[f"S{i}"] = np.asarray(S[i])
    np.savez_compressed(os.path.join(OUT, f"tandem_{name}.npz"), **out)
    print(f"tandem_{name}: thr={sim.throughput:.5f} loss={sim.loss_prob:.5f} L={[round(x, 4) for x in res.mean_jobs]}")


def make_tandem():
    rng = np.random.default_rng(20260404)
    n = 3000
    # tests/test_tandem_blocking.py:79-96: three nodes, tight middle buffer
    mu = [1.0, 0.9, 1.1]
    tandem_case("k523", [5, 2, 3], rng.exponential(1 / 0.7, 3 * n), [rng.exponential(1 / x, 3 * n) for x in mu], n)
    # unlimited buffers (test_infinite_buffers_reduce_to_jackson), non-exponential service
    tandem_case("inf2", [None, None], rng.exponential(1.0, 2 * n), [rng.gamma(2.0, 0.3, 2 * n), rng.gamma(0.7, 1.0, 2 * n)], n)
    # heavy blocking: capacity 1 everywhere, overloaded first node, lattice times (exact ties)
    tandem_case("k111_lattice", [1, 1, 1], rng.integers(1, 3, 6 * n) * 0.5,
                [rng.integers(1, 4, 6 * n) * 0.5 for _ in range(3)], n, warmup_fraction=0.0)
    tandem_case("k3_single", [3], rng.exponential(0.8, 4 * n), [rng.exponential(1.0, 4 * n)], n, warmup_fraction=0.2)
    tandem_case("k2241_five", [2, 2, 4, 1, None], rng.exponential(1.1, 4 * n), [rng.exponential(0.9, 4 * n) for _ in range(5)], n)


# --------------------------------------------------------------------------- theory
def _priority_and_network_theory():
    """Calculators SURVEY 8(d) names for BASELINE cfg4 / cfg5: the exact truncated-CTMC M/M/k preemptive-resume solver
    (theory/priority/preemptive/mmk_prty_exact.py:24), the invariant-relations approximation for M/G/n priorities
    (theory/priority/mgn_invar_approx.py:21), the exact Jackson product form (theory/networks/jackson_network.py:25)
    and the decomposition approximation (theory/networks/open_network.py:17)."""
    from most_queue.random.distributions import GammaDistribution, H2Distribution
    from most_queue.theory.networks.jackson_network import JacksonNetworkCalc
    from most_queue.theory.networks.open_network import OpenNetworkCalc
    from most_queue.theory.priority.mgn_invar_approx import MGnInvarApproximation
    from most_queue.theory.priority.preemptive.mmk_prty_exact import MMkPriorityExact

    out = {}
    lam, mu = [0.3, 0.5], [0.5, 0.25]
    with quiet():
        c = MMkPriorityExact(n=4, with_variance=True)
        c.set_sources(lam)
        c.set_servers(mu)
        r = c.run()
    out["cfg4_mm4_pr_exact"] = {"n": 4, "lam": lam, "mu": mu, "v": [list(map(float, x)) for x in r.v],
                                "w": [list(map(float, x)) for x in r.w], "utilization": float(r.utilization),
                                "boundary_mass": float(np.max(c.boundary_mass)) if c.boundary_mass is not None else None}
    g = [GammaDistribution.get_params_by_mean_and_cv(m, 1.2) for m in (2.0, 4.0)]
    b = [list(map(float, GammaDistribution.calc_theory_moments(x, 4))) for x in g]
    for prty in ("NP", "PR"):
        with quiet():
            c = MGnInvarApproximation(n=4, priority=prty)
            c.set_sources(lam)
            c.set_servers(b)
            v = c.get_v()
        out[f"cfg4_mg4_{prty.lower()}_invar"] = {"n": 4, "lam": lam, "b": b, "approximate": True,
                                                 "v": [list(map(float, np.real(x))) for x in v]}
    R = np.matrix([[1, 0, 0, 0, 0, 0], [0, 0.4, 0.6, 0, 0, 0], [0, 0, 0.2, 0.4, 0.4, 0], [0, 0, 0, 0, 1, 0],
                   [0, 0, 0, 0, 1, 0], [0, 0, 0, 0, 0, 1]], dtype=float)
    ch = [3, 2, 3, 4, 3]
    with quiet():
        c = JacksonNetworkCalc()
        c.set_sources(arrival_rate=1.0, R=R)
        lam_i = [float(x) for x in c.solve_intensities()]
        mus = [lam_i[i] / (0.7 * ch[i]) for i in range(5)]
        c.set_nodes(mu=mus, n=ch)
        r = c.run()
    out["cfg5_jackson"] = {"n": ch, "mu": mus, "intensities": lam_i, "loads": list(map(float, r.loads)),
                           "v1": float(r.v[0])}
    hs = [H2Distribution.get_params_by_mean_and_cv(0.7 * c_, 1.2) for c_ in ch]
    bh = [list(map(float, np.real(H2Distribution.calc_theory_moments(h, 4)))) for h in hs]
    with quiet():
        c = OpenNetworkCalc()
        c.set_sources(R=R, arrival_rate=1.0)
        c.set_nodes(b=bh, n=ch)
        r = c.run()
    out["cfg5_open_network_h2"] = {"n": ch, "approximate": True, "h2": [[float(h.p1), float(h.mu1), float(h.mu2)] for h in hs],
                                   "intensities": list(map(float, r.intensities)), "loads": list(map(float, r.loads)),
                                   "v": list(map(float, np.real(r.v)))}
    return out


def make_theory():
    """Analytical values used by the generator-mode (statistical) tests."""
    from most_queue.random.distributions import GammaDistribution, H2Distribution, ParetoDistribution
    from most_queue.random.utils.params import H2Params
    from most_queue.theory.fifo.mgn_takahasi import MGnCalc
    from most_queue.theory.fifo.mmnr import MMnrCalc

    out = {}
    with quiet():
        c = MMnrCalc(n=3, r=100)
        c.set_sources(2.0)
        c.set_servers(1.0)
        r = c.run()
    out["mm3_r100"] = {"l": 2.0, "mu": 1.0, "n": 3, "r": 100, "w": list(map(float, r.w)),
                       "v": list(map(float, r.v)), "p": list(map(float, r.p[:32]))}
    with quiet():
        c = MMnrCalc(n=3, r=30)
        c.set_sources(1.0)
        c.set_servers(1.0 / 2.1)
        r = c.run()
    out["mm3_r30"] = {"l": 1.0, "mu": 1.0 / 2.1, "n": 3, "r": 30, "w": list(map(float, r.w)),
                      "v": list(map(float, r.v)), "p": list(map(float, r.p[:32]))}
    with quiet():
        c = MMnrCalc(n=3, r=0)
        c.set_sources(2.0)
        c.set_servers(1.0)
        r = c.run()
    out["mm3_r0"] = {"l": 2.0, "mu": 1.0, "n": 3, "r": 0, "p": list(map(float, r.p[:4]))}

    h2 = H2Distribution.get_params_by_mean_and_cv(5.6, 1.5)
    with quiet():
        c = MGnCalc(n=8)
        c.set_sources(1.0)
        c.set_servers(H2Params(p1=h2.p1, mu1=h2.mu1, mu2=h2.mu2))
        r = c.run()
    out["mh2_8"] = {"l": 1.0, "n": 8, "h2": {"p1": float(h2.p1), "mu1": float(h2.mu1), "mu2": float(h2.mu2)},
                    "w": list(map(float, np.real(r.w))), "v": list(map(float, np.real(r.v))),
                    "p": list(map(float, np.real(r.p[:32])))}

    # fitter pins (host-side parameter helpers re-implemented in most_queue_b200.random)
    fits = {}
    for mean, cv in [(5.6, 1.5), (1.0, 1.2), (2.0, 1.05), (0.7, 3.0)]:
        h = H2Distribution.get_params_by_mean_and_cv(mean, cv)
        fits[f"h2_{mean}_{cv}"] = [float(h.p1), float(h.mu1), float(h.mu2)]
    for mean, cv in [(1.0, 0.56), (2.0, 1.2), (4.0, 1.2), (0.3, 2.0)]:
        g = GammaDistribution.get_params_by_mean_and_cv(mean, cv)
        fits[f"gamma_{mean}_{cv}"] = [float(g.mu), float(g.alpha)]
    for mean, cv in [(0.7, 1.2), (1.0, 0.8)]:
        pa = ParetoDistribution.get_params_by_mean_and_cv(mean, cv)
        fits[f"pareto_{mean}_{cv}"] = [float(pa.alpha), float(pa.K)]
    out["fits"] = fits
    # moments of the remaining samplers' laws (most_queue/random/distributions.py calc_theory_moments)
    from most_queue.random.distributions import (CoxDistribution, DeterministicDistribution, ErlangDistribution,
                                                 ExpDistribution, NormalDistribution, UniformDistribution)
    from most_queue.random.utils.params import Cox2Params, GaussianParams, UniformParams

    u = UniformDistribution.get_params_by_mean_and_cv(2.0, 0.4)
    n = NormalDistribution.get_params_by_mean_and_cv(5.0, 0.2)
    e = ErlangDistribution.get_params_by_mean_and_cv(2.0, 0.5)
    out["laws"] = {
        "uniform": {"params": [float(u.mean), float(u.half_interval)],
                    "moments": list(map(float, UniformDistribution.calc_theory_moments(u, 3))),
                    "refit": [float(x) for x in (lambda q: (q.mean, q.half_interval))(
                        UniformDistribution.get_params(UniformDistribution.calc_theory_moments(u, 3)))]},
        "normal": {"params": [float(n.mean), float(n.std_dev)],
                   "moments": list(map(float, NormalDistribution.calc_theory_moments(n, 2)))[:2]},
        "cox": {"params": [0.7, 2.0, 0.5],
                "moments": list(map(float, np.real(CoxDistribution.calc_theory_moments(
                    Cox2Params(p1=0.7, mu1=2.0, mu2=0.5), 3))))},
        "det": {"params": 1.7, "moments": list(map(float, DeterministicDistribution.calc_theory_moments(1.7, 3)))},
        "erlang": {"params": [int(e.r), float(e.mu)],
                   "moments": list(map(float, ErlangDistribution.calc_theory_moments(e, 3)))},
        "exp": {"params": 0.8, "moments": list(map(float, ExpDistribution.calc_theory_moments(0.8, 3)))},
    }
    out.update(_priority_and_network_theory())
    with open(os.path.join(OUT, "theory.json"), "w") as f:
        json.dump(out, f, indent=1)
    print("theory.json written")


if __name__ == "__main__":
    import_reference()
    which = sys.argv[1:] or ["fifo", "prio", "net", "prionet", "tandem", "vac", "neg", "size", "theory"]
    if "fifo" in which:
        make_fifo()
    if "prio" in which:
        make_prio()
    if "net" in which and "make_net" in globals():
        make_net()
    if "prionet" in which:
        make_prionet()
    if "tandem" in which:
        make_tandem()
    if "vac" in which:
        make_vac()
    if "neg" in which:
        make_neg()
    if "size" in which:
        make_size()
    if "theory" in which:
        make_theory()
