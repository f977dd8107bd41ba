"""Statistical fixtures: the UNMODIFIED reference simulators run with their own numpy generators.

    python tests/golden/make_stat.py [cfg1 cfg2 cfg4 cfg5 ...] [--seeds 64] [--procs 8]

For every BASELINE.json configuration that has a generator-mode kernel this runs the reference's own classes —
`QsSim(seed=s).run(n)` (most_queue/sim/base.py:29,488), `PriorityQueueSimulator.run` (sim/priority.py:622),
`NetworkSimulator.run` (sim/networks/network.py:196) — at SEEDS independent seeds through their public API and stock
code path, and stores the per-seed raw moments `w[0..3]`, `v[0..3]` and the state probabilities `p[:16]` in
`tests/golden/stat_<cfg>.npz`.  The `-m gpu` tests (tests/test_gpu_vs_reference_stat.py) then require the GPU mean at the
same `total_served` to lie inside the 99 % confidence interval of the reference's across-seed mean, for all four moments.

Seeding.  `QsSim` takes `seed=`.  `PriorityQueueSimulator` / `NetworkSimulator` have no seed argument: they call
`np.random.default_rng()` (sim/base_core.py:23, sim/networks/network.py:76) and the global `np.random.rand`
(sim/networks/network.py:109, random/distributions.py:416-417).  To make this script reproducible the harness — not the
reference — replaces `np.random.default_rng` by a function that returns PCG64 generators seeded from (run seed, call index)
whenever it is called without a seed, and calls `np.random.seed(run seed)`; the reference's code is untouched.
"""

from __future__ import annotations

import multiprocessing as mp
import os
import sys
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
from make_golden import import_reference, quiet  # noqa: E402

OUT = HERE
NP_BINS = 16

# cfg5 topology (tests/test_network_no_prty.py:33-46 of the reference)
NET_R = [[1, 0, 0, 0, 0, 0], [0, 0.4, 0.6, 0, 0, 0], [0, 0, 0.2, 0.4, 0.4, 0], [0, 0, 0, 0, 1, 0], [0, 0, 0, 0, 1, 0],
         [0, 0, 0, 0, 0, 1]]
NET_C = [3, 2, 3, 4, 3]


def _seeded_numpy(seed: int):
    """Harness-level seeding of numpy for the simulators that take no seed (see the module docstring)."""
    real = np.random.default_rng
    if getattr(real, "_mq_wrapped", None) is not None:
        real = real._mq_wrapped
    calls = [0]

    def rng(s=None):
        if s is None:
            calls[0] += 1
            return real(np.random.SeedSequence([20261020, seed, calls[0]]))
        return real(s)

    rng._mq_wrapped = real
    np.random.default_rng = rng
    np.random.seed(seed)


def _fifo(args):
    c, buffer, src, srv, n, seed = args
    import_reference()
    from most_queue.sim.base import QsSim

    with quiet():
        qs = QsSim(c, buffer=buffer, verbose=False, seed=seed)
        qs.set_sources(*src)
        qs.set_servers(*srv)
        t0 = time.perf_counter()
        r = qs.run(n)
        dt = time.perf_counter() - t0
    return dict(w=np.array(r.w, float), v=np.array(r.v, float), p=np.array(r.p[:NP_BINS], float),
                busy=np.array(qs.busy, float), served=qs.served, dropped=qs.dropped, arrived=qs.arrived, sec=dt)


def _prio(args):
    c, k, prty, srcs, srvs, n, seed = args
    import_reference()
    from most_queue.sim.priority import PriorityQueueSimulator

    _seeded_numpy(seed)
    with quiet():
        qs = PriorityQueueSimulator(c, k, prty)
        qs.set_sources(srcs)
        qs.set_servers(srvs)
        t0 = time.perf_counter()
        r = qs.run(n)
        dt = time.perf_counter() - t0
    return dict(w=np.array(r.w, float), v=np.array(r.v, float), p=np.array([pk[:NP_BINS] for pk in r.p], float),
                served=np.array(qs.served), sec=dt)


def _net(args):
    lam, serv, n, seed = args
    import_reference()
    from most_queue.sim.networks.network import NetworkSimulator

    _seeded_numpy(seed)
    with quiet():
        qn = NetworkSimulator()
        qn.set_sources(arrival_rate=lam, R=np.matrix(NET_R, dtype=float))
        qn.set_nodes(serv_params=serv, n=NET_C)
        t0 = time.perf_counter()
        r = qn.run(n)
        dt = time.perf_counter() - t0
    node_v = np.array([q.v for q in qn.qs], float)
    node_w = np.array([q.w for q in qn.qs], float)
    node_served = np.array([q.served for q in qn.qs])
    return dict(v=np.array(r.v, float), w=np.array(qn.w_network, float), node_v=node_v, node_w=node_w,
                node_served=node_served, served=r.served, arrived=r.arrived, sec=dt)


def _stack(rows):
    keys = rows[0].keys()
    return {k: np.array([r[k] for r in rows]) for k in keys}


def _run(pool, fn, argsets, name, meta):
    t0 = time.perf_counter()
    rows = pool.map(fn, argsets, chunksize=1)
    st = _stack(rows)
    np.savez_compressed(os.path.join(OUT, f"stat_{name}.npz"), seeds=np.array([a[-1] for a in argsets]), **st, **meta)
    w = st["w"].mean(0) if "w" in st else None
    print(f"stat_{name}: {len(rows)} seeds in {time.perf_counter() - t0:.0f} s; v1 mean={np.ravel(st['v'].mean(0))[:2]} "
          f"w1 mean={None if w is None else np.ravel(w)[:2]}", flush=True)


def main():
    argv = [a for a in sys.argv[1:] if not a.startswith("--")]
    opt = {a.split("=")[0]: a.split("=")[1] for a in sys.argv[1:] if a.startswith("--") and "=" in a}
    seeds = int(opt.get("--seeds", 64))
    procs = int(opt.get("--procs", os.cpu_count() or 1))
    which = argv or ["cfg1", "cfg1r30", "cfg2", "cfg4", "cfg4m", "cfg5", "cfg5m"]
    import_reference()
    from most_queue.random.distributions import GammaDistribution, H2Distribution
    from most_queue.random.utils.params import H2Params

    S = list(range(1000, 1000 + seeds))
    with mp.get_context("fork").Pool(procs) as pool:
        if "cfg1" in which:  # README quick start: M/M/3 with buffer 100, lambda 2, mu 1, 1e5 jobs
            _run(pool, _fifo, [(3, 100, (2.0, "M"), (1.0, "M"), 100_000, s) for s in S], "cfg1_mm3_r100",
                 dict(c=3, buffer=100, lam=2.0, mu=1.0, total_served=100_000))
        if "cfg1r30" in which:  # tests/test_qs_sim.py:46-55: M/M/3/30, lambda 1, mu 1/2.1
            _run(pool, _fifo, [(3, 30, (1.0, "M"), (1.0 / 2.1, "M"), 100_000, s) for s in S], "cfg1_mm3_r30",
                 dict(c=3, buffer=30, lam=1.0, mu=1.0 / 2.1, total_served=100_000))
        if "cfg2" in which:  # M/H2/8, rho 0.7, cv 1.5, 1e5 jobs
            h = H2Distribution.get_params_by_mean_and_cv(5.6, 1.5)
            h2 = H2Params(p1=float(h.p1), mu1=float(h.mu1), mu2=float(h.mu2))
            _run(pool, _fifo, [(8, None, (1.0, "M"), (h2, "H"), 100_000, s) for s in S], "cfg2_mh2_8",
                 dict(c=8, buffer=-1, lam=1.0, h2=np.array([h2.p1, h2.mu1, h2.mu2]), total_served=100_000))
        if "cfg4" in which:  # M/Gamma/4, two classes, NP and PR, 1e5 jobs
            g = [GammaDistribution.get_params_by_mean_and_cv(m, 1.2) for m in (2.0, 4.0)]
            srcs = [{"type": "M", "params": 0.3}, {"type": "M", "params": 0.5}]
            srvs = [{"type": "Gamma", "params": x} for x in g]
            for prty in ("NP", "PR"):
                _run(pool, _prio, [(4, 2, prty, srcs, srvs, 100_000, s) for s in S], f"cfg4_mg4_{prty}",
                     dict(c=4, k=2, lam=np.array([0.3, 0.5]), gamma=np.array([[x.mu, x.alpha] for x in g]),
                          total_served=100_000))
        if "cfg4m" in which:  # the exact sub-case: M service mu = [0.5, 0.25]
            srcs = [{"type": "M", "params": 0.3}, {"type": "M", "params": 0.5}]
            srvs = [{"type": "M", "params": 0.5}, {"type": "M", "params": 0.25}]
            for prty in ("NP", "PR"):
                _run(pool, _prio, [(4, 2, prty, srcs, srvs, 100_000, s) for s in S], f"cfg4_mm4_{prty}",
                     dict(c=4, k=2, lam=np.array([0.3, 0.5]), mu=np.array([0.5, 0.25]), total_served=100_000))
        if "cfg5" in which:  # 5-node network of H2 nodes, 1e4 network jobs
            hs = [H2Distribution.get_params_by_mean_and_cv(0.7 * c, 1.2) for c in NET_C]
            serv = [{"type": "H", "params": H2Params(p1=float(x.p1), mu1=float(x.mu1), mu2=float(x.mu2))} for x in hs]
            _run(pool, _net, [(1.0, serv, 10_000, s) for s in S], "cfg5_net_h2",
                 dict(h2=np.array([[x.p1, x.mu1, x.mu2] for x in hs]), total_served=10_000))
        if "cfg5m" in which:  # Jackson sub-case (tests/test_jackson_network.py:29-31), mu_i = lambda_i / (0.7 c_i)
            from most_queue.theory.networks.jackson_network import JacksonNetworkCalc

            calc = JacksonNetworkCalc()
            calc.set_sources(arrival_rate=1.0, R=np.matrix(NET_R, dtype=float))
            lam_i = calc.solve_intensities()
            mu = [float(lam_i[i] / (0.7 * NET_C[i])) for i in range(5)]
            serv = [{"type": "M", "params": m} for m in mu]
            _run(pool, _net, [(1.0, serv, 10_000, s) for s in S], "cfg5_net_jackson",
                 dict(mu=np.array(mu), total_served=10_000))


if __name__ == "__main__":
    main()
