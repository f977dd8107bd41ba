"""K5 (open-network kernel) through the C ABI: replay tier bit-exact against the reference's golden
fixtures and the oracle; generator tier against the oracle driven by the same Philox words."""

import glob
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

CASES = sorted(glob.glob(os.path.join(os.path.dirname(__file__), "golden", "net_*.npz")))
NET_R = [[1, 0, 0, 0, 0, 0], [0, 0.4, 0.6, 0, 0, 0], [0, 0, 0.2, 0.4, 0.4, 0], [0, 0, 0, 0, 1, 0],
         [0, 0, 0, 0, 1, 0], [0, 0, 0, 0, 0, 1]]
NET_CH = [3, 2, 3, 4, 3]


def _ulps(a, b):
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    return np.abs(a.view(np.int64) - b.view(np.int64))


@pytest.mark.parametrize("path", CASES, ids=[os.path.basename(p)[4:-4] for p in CASES])
def test_replay_matches_reference(path):
    from most_queue_b200.sim.networks.network import replay
    from oracle import oracle

    g = np.load(path)
    m = len(g["channels"])
    rep = lambda a: np.repeat(np.asarray(a)[:, None], 2, axis=1)
    out = replay(g["R"], g["channels"], rep(g["A"]), rep(g["U"]), [rep(g[f"S{i}"]) for i in range(m)],
                 int(g["job_served"]))
    o = oracle.net_replay(g["R"], g["channels"], g["A"], g["U"], [g[f"S{i}"] for i in range(m)], int(g["job_served"]))
    for r in range(2):
        assert out["ttek"][r] == float(g["ttek"])
        assert (out["served"][r], out["arrived"][r], out["in_sys"][r]) == (int(g["served"]), int(g["arrived"]), int(g["in_sys"]))
        assert (out["a_used"][r], out["u_used"][r]) == (int(g["a_used"]), int(g["u_used"]))
        assert np.array_equal(out["s_used"][r], g["s_used"])
        # node-level running moments (repeated multiplication): bit-exact with the reference
        assert np.array_equal(out["node_v"][r], g["node_v"]) and np.array_equal(out["node_w"][r], g["node_w"])
        assert np.array_equal(out["node_served"][r], g["node_served"])
        assert np.array_equal(out["node_taked"][r], g["node_taked"])
        # network-level power sums: bit-exact with the oracle (same samples, same order)
        assert np.array_equal(out["v_sum"][r], o.v_sum) and np.array_equal(out["w_sum"][r], o.w_sum)
        # network-level running moments use math.pow in the reference: first two bit-exact,
        # the cube within 2 ulp per update -> a few ulp on the running mean
        assert np.array_equal(out["v"][r][:2], g["v"][:2]) and np.array_equal(out["w"][r][:2], g["w"][:2])
        assert _ulps(out["v"][r][2], g["v"][2]) <= 64 and _ulps(out["w"][r][2], g["w"][2]) <= 64


def test_replay_matches_oracle_many_replications():
    from most_queue_b200.sim.networks.network import replay
    from oracle import oracle

    rng = np.random.default_rng(99)
    reps, n = 48, 400
    A = rng.exponential(1.0, (2 * n, reps))
    U = rng.random((14 * n, reps))
    S = [rng.gamma(0.8, (0.75 * c) / 0.8, (7 * n, reps)) for c in NET_CH]
    out = replay(NET_R, NET_CH, A, U, S, n)
    for r in range(reps):
        o = oracle.net_replay(NET_R, NET_CH, A[:, r].copy(), U[:, r].copy(), [s[:, r].copy() for s in S], n)
        assert o.ttek == out["ttek"][r] and o.arrived == out["arrived"][r] and o.u_used == out["u_used"][r]
        assert np.array_equal(o.v_sum, out["v_sum"][r]) and np.array_equal(o.w_sum, out["w_sum"][r])
        assert np.array_equal(o.node_v, out["node_v"][r]) and np.array_equal(o.node_w, out["node_w"][r])
        assert np.array_equal(o.node_v_sum, out["node_v_sum"][r])
        assert np.array_equal(o.s_used, out["s_used"][r])


def test_generator_matches_oracle():
    """cfg5 shape: H2 service per node with mean 0.7 c_i / lambda, cv 1.2 (tests/test_network_no_prty.py)."""
    from most_queue_b200 import _lib
    from most_queue_b200.random.distributions import H2Distribution
    from oracle import oracle

    srv = [("H", H2Distribution.get_params_by_mean_and_cv(0.7 * c, 1.2)) for c in NET_CH]
    ctx = _lib.default_context()
    seed, reps, n = 3, 32, 1500
    agg, rec = ctx.sim_net(NET_R, NET_CH, _lib.make_dist("M", 1.0), [_lib.make_dist(*s) for s in srv], n, reps, 2,
                           seed, per_replication=True)
    same = 0
    for r in range(reps):
        o = oracle.net_generate(NET_R, NET_CH, ("M", 1.0), srv, n, _lib.seed_to_key(seed), 2 + r)
        assert int(rec[_lib.NG_SERVED, r]) == n
        if int(rec[_lib.NG_ARRIVED, r]) == o.arrived and int(rec[_lib.NG_UUSED, r]) == o.u_used:
            same += 1
            np.testing.assert_allclose(rec[_lib.NG_VSUM, r], o.v_sum[0], rtol=2e-4)
            np.testing.assert_allclose(rec[_lib.NG_TTEK, r], o.ttek, rtol=1e-5)
    assert same >= reps * 3 // 4, same  # the rest: an fp32-draw near-tie resolved the other way (another sample path)
    np.testing.assert_allclose(agg[5] / reps, (rec[_lib.NG_VSUM] / rec[_lib.NG_SERVED]).mean(), rtol=1e-12)


def test_network_simulator_api_jackson_mean():
    """Jackson sub-case (tests/test_jackson_network.py:60-71): exponential service, utilisation 0.7 per
    node; mean sojourn = sum of visit ratios times M/M/c node sojourn times."""
    from most_queue_b200.sim.networks.network import NetworkSimulator

    Rm = np.array(NET_R, dtype=float)
    m = 5
    P = Rm[1:, :m]
    lam = np.linalg.solve(np.eye(m) - P.T, Rm[0, :m])
    mu = [lam[i] / (0.7 * NET_CH[i]) for i in range(m)]

    def mmc_sojourn(l, mu_, c):
        from math import factorial
        a, rho = l / mu_, l / (c * mu_)
        p0 = 1.0 / (sum(a**k / factorial(k) for k in range(c)) + a**c / (factorial(c) * (1 - rho)))
        lq = p0 * a**c * rho / (factorial(c) * (1 - rho) ** 2)
        return lq / l + 1.0 / mu_

    v_theory = sum(lam[i] * mmc_sojourn(lam[i], mu[i], NET_CH[i]) for i in range(m)) / 1.0
    sim = NetworkSimulator(seed=42, replications=512)
    sim.set_sources(arrival_rate=1.0, R=np.matrix(NET_R))
    sim.set_nodes(serv_params=[{"type": "M", "params": x} for x in mu], n=NET_CH)
    res = sim.run(20000)
    assert res.served == pytest.approx(20000)
    assert abs(res.v[0] - v_theory) < res.ci["v"][0] + 5e-3 * v_theory, (res.v[0], v_theory)
    assert len(sim.qs) == 5 and sim.qs[0].v[0] > 0
