"""K6 (tandem line with blocking after service) through the C ABI: replay tier bit-exact against the
reference's golden fixtures and the oracle; generator tier against the oracle and exact M/M/1/K theory."""

import glob
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

CASES = sorted(glob.glob(os.path.join(os.path.dirname(__file__), "golden", "tandem_*.npz")))


@pytest.mark.parametrize("path", CASES, ids=[os.path.basename(p)[7:-4] for p in CASES])
def test_replay_matches_reference_bitwise(path):
    from most_queue_b200.sim.networks.tandem_blocking import replay

    g = np.load(path)
    cap = [None if k < 0 else int(k) for k in g["capacity"]]
    m = len(cap)
    rep = lambda a: np.repeat(np.asarray(a, dtype=np.float64)[:, None], 2, axis=1)
    out = replay(cap, rep(g["A"]), [rep(g[f"S{i}"]) for i in range(m)], int(g["total_served"]),
                 float(g["warmup_fraction"]))
    for r in range(2):
        assert out["throughput"][r] == float(g["throughput"])
        assert out["loss_prob"][r] == float(g["loss_prob"])
        assert out["v0"][r] == float(g["v0"])
        assert np.array_equal(out["mean_jobs"][r], g["mean_jobs"])
        assert out["served"][r] == int(g["served"]) and out["a_used"][r] == int(g["a_used"])
        assert np.array_equal(out["s_used"][r], g["s_used"])


def test_replay_matches_oracle_many_replications():
    from most_queue_b200.sim.networks.tandem_blocking import replay
    from oracle import oracle

    rng = np.random.default_rng(8)
    reps, n = 64, 500
    cap = [4, 2, None, 3]
    A = rng.exponential(1.0, (4 * n, reps))
    S = [rng.gamma(1.3, 0.7 / 1.3, (4 * n, reps)) for _ in cap]
    out = replay(cap, A, S, n, 0.1)
    for r in range(reps):
        o = oracle.tandem_replay(cap, A[:, r].copy(), [s[:, r].copy() for s in S], n, 0.1)
        assert out["throughput"][r] == o.throughput and out["loss_prob"][r] == o.loss_prob
        assert np.array_equal(out["mean_jobs"][r], np.array(o.mean_jobs[:4])) and out["v0"][r] == o.v0
        assert out["a_used"][r] == o.a_used and list(out["s_used"][r]) == o.s_used


def test_generator_matches_oracle_and_mm1k_theory():
    from most_queue_b200 import _lib
    from most_queue_b200.sim.networks.tandem_blocking import TandemBlockingSim
    from oracle import oracle

    # per-replication agreement with the oracle driven by the same Philox words
    cap, mu, lam = [5, 2, 3], [1.0, 0.9, 1.1], 0.7
    ctx = _lib.default_context()
    agg, rec = ctx.sim_tandem(cap, _lib.make_dist("M", lam), [_lib.make_dist("M", x) for x in mu], 2000, 0.05, 32, 4, 21,
                              per_replication=True)
    same = 0
    for r in range(32):
        o = oracle.tandem_generate(cap, ("M", lam), [("M", x) for x in mu], 2000, _lib.seed_to_key(21), 4 + r, 0.05)
        if int(rec[_lib.TG_ARRIVED, r]) == o.arrived and int(rec[_lib.TG_LOST, r]) == o.lost:
            same += 1
            np.testing.assert_allclose(rec[_lib.TG_ELAPSED, r], o.elapsed, rtol=1e-5)
            np.testing.assert_allclose(rec[_lib.TG_AREASUM, r], sum(o.area[:3]), rtol=1e-4)
    assert same >= 24

    # one node = M/M/1/K: loss probability and mean number in system in closed form
    K, rho = 4, 0.8
    sim = TandemBlockingSim(seed=3, replications=2048)
    sim.set_sources(rho)
    sim.set_nodes([{"type": "M", "params": 1.0}], [K])
    res = sim.run(20000)
    pk = np.array([rho**k for k in range(K + 1)])
    pk /= pk.sum()
    assert abs(sim.loss_prob - pk[K]) < sim.ci["loss_prob"] + 2e-3
    assert abs(res.mean_jobs[0] - (np.arange(K + 1) * pk).sum()) < sim.ci["mean_jobs"][0] + 5e-3
    assert abs(sim.throughput - rho * (1 - pk[K])) < sim.ci["throughput"] + 2e-3
    assert res.served == 21000
