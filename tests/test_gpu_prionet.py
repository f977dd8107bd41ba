"""K7 (priority-network kernel) through the C ABI: replay tier bit-exact against the reference's golden
fixtures and the oracle; generator tier against the oracle driven by the same Philox words; the host mirror
PriorityNetwork against the single-node priority kernel and an M/M/1 priority closed form."""

import glob
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

CASES = sorted(glob.glob(os.path.join(os.path.dirname(__file__), "golden", "prionet_*.npz")))
R0 = [[1, 0, 0, 0], [0, 0, 0.7, 0.3], [0, 0, 0, 1], [0, 0, 0, 1]]
R1 = [[0.5, 0.5, 0, 0], [0, 0, 0, 1], [0.2, 0, 0.3, 0.5], [0, 0, 0, 1]]
CH = [2, 1, 2]


def _ulps(a, b):
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    return np.abs(a.view(np.int64) - b.view(np.int64))


def _load(g):
    m, K = len(g["channels"]), g["R"].shape[0]
    return m, K, [g[f"A{k}"] for k in range(K)], [[g[f"S{i}_{j}"] for j in range(K)] for i in range(m)]


@pytest.mark.parametrize("path", CASES, ids=[os.path.basename(p)[8:-4] for p in CASES])
def test_replay_matches_reference(path):
    from most_queue_b200.sim.networks.priority_network import replay
    from oracle import oracle

    g = np.load(path)
    m, K, A, S = _load(g)
    prty = [str(x) for x in g["prty"]]
    rep = lambda a: np.repeat(np.asarray(a)[:, None], 2, axis=1)
    out = replay(g["R"], g["channels"], prty, g["nodes_prty"], [rep(a) for a in A], rep(g["U"]),
                 [[rep(x) for x in row] for row in S], int(g["job_served"]))
    o = oracle.prionet_replay(g["R"], g["channels"], prty, g["nodes_prty"], A, g["U"], S, int(g["job_served"]))
    for r in range(2):
        assert out["ttek"][r] == float(g["ttek"])
        assert np.array_equal(out["served"][r], g["served"]) and np.array_equal(out["arrived"][r], g["arrived"])
        assert np.array_equal(out["a_used"][r], g["a_used"]) and out["u_used"][r] == int(g["u_used"])
        assert np.array_equal(out["s_used"][r], g["s_used"])
        # node-level running moments (repeated multiplication): bit-exact with the reference
        assert np.array_equal(out["node_v"][r], g["node_v"]) and np.array_equal(out["node_w"][r], g["node_w"])
        assert np.array_equal(out["node_served"][r], g["node_served"])
        assert np.array_equal(out["node_taked"][r], g["node_taked"])
        # network-level power sums: bit-exact with the oracle
        assert np.array_equal(out["v_sum"][r], o.v_sum) and np.array_equal(out["w_sum"][r], o.w_sum)
        # network-level running moments use math.pow in the reference: the first two bit-exact, the cube
        # within a few ulp on the running mean
        assert np.array_equal(out["v"][r][:, :2], g["v"][:, :2]) and np.array_equal(out["w"][r][:, :2], g["w"][:, :2])
        assert np.all(_ulps(out["v"][r][:, 2], g["v"][:, 2]) <= 64)
        assert np.all(_ulps(out["w"][r][:, 2], g["w"][:, 2]) <= 64)


@pytest.mark.parametrize("prty", [["NP", "PR", "RS"], ["RW", "RW", "PR"]])
def test_replay_matches_oracle_many_replications(prty):
    from most_queue_b200.sim.networks.priority_network import replay
    from oracle import oracle

    rng = np.random.default_rng(17)
    reps, n = 40, 300
    A = [rng.exponential(1.6, (2 * n, reps)), rng.exponential(1.4, (2 * n, reps))]
    U = rng.random((12 * n, reps))
    S = [[rng.gamma(0.7, 1.0, (8 * n, reps)) * (0.6 + 0.3 * j) for j in range(2)] for _ in range(3)]
    npr = [[0, 1], [1, 0], [0, 1]]
    out = replay([R0, R1], CH, prty, npr, A, U, S, n)
    for r in range(reps):
        o = oracle.prionet_replay([R0, R1], CH, prty, npr, [a[:, r].copy() for a in A], U[:, r].copy(),
                                  [[x[:, r].copy() for x in row] for row in S], n)
        assert o.ttek == out["ttek"][r] and o.u_used == out["u_used"][r]
        assert np.array_equal(o.arrived, out["arrived"][r]) and np.array_equal(o.served, out["served"][r])
        assert np.array_equal(o.v_sum, out["v_sum"][r]) and np.array_equal(o.w_sum, out["w_sum"][r])
        assert np.array_equal(o.node_v, out["node_v"][r]) and np.array_equal(o.node_w, out["node_w"][r])
        assert np.array_equal(o.s_used, out["s_used"][r]) and np.array_equal(o.node_taked, out["node_taked"][r])


def test_short_arrays_and_overflow_are_reported():
    from most_queue_b200 import _lib
    from most_queue_b200.sim.networks.priority_network import replay

    rng = np.random.default_rng(5)
    n = 200
    A = [rng.exponential(1.6, (2 * n, 3)), rng.exponential(1.4, (2 * n, 3))]
    U = rng.random((12 * n, 3))
    S = [[rng.exponential(0.6, (8 * n, 3)) for _ in range(2)] for _ in range(3)]
    with pytest.raises(_lib.MqError) as e:
        replay([R0, R1], CH, ["PR"] * 3, [[0, 1]] * 3, A, U[:50], S, n)
    assert e.value.code == _lib.MQ_E_INPUT_SHORT
    slow = [[x * 40.0 for x in row] for row in S]
    with pytest.raises(_lib.MqError) as e:
        replay([R0, R1], CH, ["PR"] * 3, [[0, 1]] * 3, A, U, slow, n, queue_cap=4)
    assert e.value.code == _lib.MQ_E_OVERFLOW
    with pytest.raises(ValueError):
        replay([R0, R1], CH, ["PR", "XX", "NP"], [[0, 1]] * 3, A, U, S, n)


def test_generator_matches_oracle():
    from most_queue_b200 import _lib
    from most_queue_b200.random.distributions import H2Distribution
    from oracle import oracle

    K, m = 2, 3
    npr = [[0, 1], [1, 0], [0, 1]]
    prty = ["PR", "NP", "RS"]
    srv = [[("H", H2Distribution.get_params_by_mean_and_cv(0.5 + 0.2 * j + 0.1 * i, 1.3)) for j in range(K)]
           for i in range(m)]
    src = [("M", 0.6), ("M", 0.7)]
    ctx = _lib.default_context()
    seed, reps, n = 11, 32, 1200
    agg, rec = ctx.sim_prionet([R0, R1], CH, prty, npr, [_lib.make_dist(*s) for s in src],
                               [[_lib.make_dist(*x) for x in row] for row in srv], n, reps, 4, seed,
                               per_replication=True)
    L = _lib
    same = 0
    for r in range(reps):
        o = oracle.prionet_generate([R0, R1], CH, prty, npr, src, srv, n, L.seed_to_key(seed), 4 + r)
        cls = lambda k, f: rec[L.PN_G_ROWS + k * L.PNC_ROWS + f, r]
        assert sum(int(cls(k, L.PNC_SERVED)) for k in range(K)) == n
        if all(int(cls(k, L.PNC_ARRIVED)) == o.arrived[k] for k in range(K)) and int(rec[L.PN_G_UUSED, r]) == o.u_used:
            same += 1
            np.testing.assert_allclose([cls(k, L.PNC_VSUM) for k in range(K)], o.v_sum[:, 0], rtol=3e-4)
            np.testing.assert_allclose(rec[L.PN_G_TTEK, r], o.ttek, rtol=1e-5)
    assert same >= reps // 2
    v0 = (rec[L.PN_G_ROWS + L.PNC_VSUM] / rec[L.PN_G_ROWS + L.PNC_SERVED]).mean()
    np.testing.assert_allclose(agg[4] / reps, v0, rtol=1e-12)


def test_priority_network_api_single_node_mm1_pr():
    """One PR node, two classes, exponential service: the network reduces to M/M/1 with preemptive-resume
    priorities.  v1 = (1/mu1) / (1 - rho1); v2 = [1/mu2 (1 - rho1 - rho2) + R] / ((1 - rho1)(1 - rho1 - rho2))
    with R = l1/mu1^2 + l2/mu2^2 (Kleinrock vol. 2, preemptive resume)."""
    from most_queue_b200.sim.networks.priority_network import PriorityNetwork

    l1, l2, mu1, mu2 = 0.3, 0.25, 1.0, 0.8
    r1, r2 = l1 / mu1, l2 / mu2
    Rsum = l1 / mu1**2 + l2 / mu2**2
    v1 = (1 / mu1) / (1 - r1)
    v2 = ((1 / mu2) * (1 - r1 - r2) + Rsum) / ((1 - r1) * (1 - r1 - r2))
    Rm = [np.matrix([[1.0, 0.0], [0.0, 1.0]])] * 2
    sim = PriorityNetwork(2, seed=7, replications=512)
    sim.set_sources([l1, l2], Rm)
    sim.set_nodes([[{"type": "M", "params": mu1}, {"type": "M", "params": mu2}]], [1], ["PR"], [[0, 1]])
    res = sim.run(20000)
    assert sum(res.served) == pytest.approx(20000)
    assert abs(res.v[0][0] - v1) < res.ci["v"][0][0] + 5e-3 * v1, (res.v[0][0], v1)
    assert abs(res.v[1][0] - v2) < res.ci["v"][1][0] + 1e-2 * v2, (res.v[1][0], v2)
    with pytest.raises(ValueError):
        PriorityNetwork(2).run(10)
