"""Every template instantiation and every general (fallback) kernel stays under the bit-exact oracle check: the small
configurations run the register / shared-memory resident kernels (k4v2, k5v2, k6v2, k7v2), larger ones the general
kernels (k4_prio, k5_net, k6_tandem, k7_prionet), and MQ_K*_V1=1 forces the general kernel on a small configuration."""

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _check_prio(c, K, prty, buffer, seed):
    from most_queue_b200.sim.priority import replay
    from oracle import oracle

    rng = np.random.default_rng(seed)
    R, n = 24, 250
    A = [rng.exponential(K * 1.1 / c * (1 + 0.2 * k), (3 * n, R)) for k in range(K)]
    S = [rng.gamma(0.7, 1.3 * (1 + 0.3 * k), (4 * n, R)) for k in range(K)]
    out = replay(c, K, prty, A, S, n, buffer=buffer, p_bins=32)
    for r in range(R):
        o = oracle.prio_replay(c, K, prty, [a[:, r].copy() for a in A], [s[:, r].copy() for s in S], n, buffer=buffer,
                               p_bins=32)
        assert np.array_equal(out["v"][r], o.v) and np.array_equal(out["w"][r], o.w)
        assert np.array_equal(out["p_time"][r], o.p_time) and out["ttek"][r] == o.ttek
        assert np.array_equal(out["dropped"][r], o.dropped)


@pytest.mark.parametrize("c,K,prty,buffer", [
    (8, 2, "PR", None),   # k4v2<8,2>
    (5, 1, "NP", None),   # k4v2<8,2>, one class
    (7, 4, "RS", 4),      # k4v2<8,4>, finite buffer
    (2, 4, "RW", 3),      # k4v2<4,4>, finite buffer
    (10, 2, "PR", None),  # general kernel: c > 8
    (3, 5, "NP", None),   # general kernel: K > 4
    (12, 6, "RW", 5),     # general kernel
])
def test_priority_kernels_all_variants(c, K, prty, buffer):
    _check_prio(c, K, prty, buffer, 1000 + 10 * c + K)


def test_priority_general_kernel_on_a_small_configuration(monkeypatch):
    monkeypatch.setenv("MQ_K4_V1", "1")
    _check_prio(4, 2, "PR", None, 77)
    _check_prio(3, 3, "RS", 4, 78)


def _ring_network(m, ch, rng):
    """Feed-forward ring with feedback: node i -> i+1 w.p. 0.6, back to 0 w.p. 0.1, out w.p. 0.3."""
    Rm = np.zeros((m + 1, m + 1))
    Rm[0, 0] = 1.0
    for i in range(m):
        nxt = (i + 1) % m
        Rm[i + 1, nxt] += 0.6
        Rm[i + 1, 0] += 0.1 if nxt != 0 else 0.0
        Rm[i + 1, m] = 1.0 - Rm[i + 1, :m].sum()
    return Rm


def _check_net(m, ch, seed):
    from most_queue_b200.sim.networks.network import replay
    from oracle import oracle

    rng = np.random.default_rng(seed)
    Rm = _ring_network(m, ch, rng)
    R, n = 16, 200
    A = rng.exponential(1.0, (3 * n, R))
    U = rng.random((16 * n, R))
    S = [rng.gamma(0.8, 0.25 * c, (8 * n, R)) for c in ch]
    out = replay(Rm, ch, A, U, S, n)
    for r in range(R):
        o = oracle.net_replay(Rm, ch, A[:, r].copy(), U[:, r].copy(), [s[:, r].copy() for s in S], n)
        assert o.ttek == out["ttek"][r] and o.arrived == out["arrived"][r] and o.u_used == out["u_used"][r]
        assert np.array_equal(o.v_sum, out["v_sum"][r]) and np.array_equal(o.w_sum, out["w_sum"][r])
        assert np.array_equal(o.node_v, out["node_v"][r]) and np.array_equal(o.node_w, out["node_w"][r])
        assert np.array_equal(o.s_used, out["s_used"][r]) and np.array_equal(o.node_taked, out["node_taked"][r])


@pytest.mark.parametrize("m,ch", [
    (2, [1, 2]),                    # k5v2<8>
    (4, [2, 2, 2, 2]),              # k5v2<8>, exactly 8 servers
    (6, [3, 2, 3, 2, 3, 3]),        # k5v2<16>, 16 servers
    (3, [6, 6, 6]),                 # general kernel: 18 servers
    (10, [1] * 10),                 # general kernel: 10 nodes
])
def test_network_kernels_all_variants(m, ch):
    _check_net(m, ch, 2000 + m)


def test_network_general_kernel_on_a_small_configuration(monkeypatch):
    monkeypatch.setenv("MQ_K5_V1", "1")
    _check_net(3, [2, 1, 2], 88)


def _check_prionet(m, K, ch, prty, seed):
    from most_queue_b200.sim.networks.priority_network import replay
    from oracle import oracle

    rng = np.random.default_rng(seed)
    Rk = []
    for k in range(K):
        Rm = np.zeros((m + 1, m + 1))
        Rm[0, k % m] = 1.0
        for i in range(m):
            Rm[i + 1, (i + 1 + k) % m] = 0.5
            Rm[i + 1, m] = 0.5
        Rk.append(Rm)
    npr = [[(k + i) % K for k in range(K)] for i in range(m)]
    R, n = 12, 200
    A = [rng.exponential(1.2 * K, (3 * n, R)) for _ in range(K)]
    U = rng.random((12 * n, R))
    S = [[rng.gamma(0.8, 0.4 * ch[i], (8 * n, R)) for _ in range(K)] for i in range(m)]
    out = replay(Rk, ch, prty, npr, A, U, S, n)
    for r in range(R):
        o = oracle.prionet_replay(Rk, ch, prty, npr, [a[:, r].copy() for a in A], U[:, r].copy(),
                                  [[x[:, r].copy() for x in row] for row in S], n)
        assert o.ttek == out["ttek"][r] and o.u_used == out["u_used"][r]
        assert np.array_equal(o.served, out["served"][r]) and np.array_equal(o.arrived, out["arrived"][r])
        assert np.array_equal(o.v_sum, out["v_sum"][r]) and np.array_equal(o.w_sum, out["w_sum"][r])
        assert np.array_equal(o.node_v, out["node_v"][r]) and np.array_equal(o.node_w, out["node_w"][r])
        assert np.array_equal(o.s_used, out["s_used"][r]) and np.array_equal(o.node_taked, out["node_taked"][r])


@pytest.mark.parametrize("m,K,ch,prty", [
    (2, 2, [2, 2], ["PR", "NP"]),                          # k7v2<8>
    (4, 3, [3, 3, 3, 3], ["RS", "RW", "PR", "NP"]),        # k7v2<16>, 12 servers
    (5, 4, [3, 3, 3, 3, 4], ["PR", "RS", "NP", "RW", "PR"]),  # k7v2<16>, 16 servers, 4 classes
    (4, 2, [5, 5, 5, 5], ["PR", "NP", "RS", "RW"]),        # general kernel: 20 servers
])
def test_priority_network_kernels_all_variants(m, K, ch, prty):
    _check_prionet(m, K, ch, prty, 3000 + m * 10 + K)


def test_priority_network_general_kernel_on_a_small_configuration(monkeypatch):
    monkeypatch.setenv("MQ_K7_V1", "1")
    _check_prionet(3, 2, [2, 1, 2], ["PR", "RS", "NP"], 99)


def _check_tandem(cap, seed, reps=24, n=300):
    """Replay tier of the tandem line, bit-exact against the oracle, on a line that blocks often."""
    from most_queue_b200.sim.networks.tandem_blocking import replay
    from oracle import oracle

    rng = np.random.default_rng(seed)
    m = len(cap)
    A = rng.exponential(0.6, (6 * n, reps))
    S = [rng.gamma(1.2, (0.5 + 0.1 * (i % 3)) / 1.2, (6 * n, reps)) for i in range(m)]
    out = replay(cap, A, S, n, 0.1)
    for r in range(reps):
        o = oracle.tandem_replay(cap, A[:, r].copy(), [s[:, r].copy() for s in S], n, 0.1)
        assert out["throughput"][r] == o.throughput and out["loss_prob"][r] == o.loss_prob
        assert np.array_equal(out["mean_jobs"][r], np.array(o.mean_jobs[:m])) and out["v0"][r] == o.v0
        assert out["a_used"][r] == o.a_used and list(out["s_used"][r]) == o.s_used


@pytest.mark.parametrize("cap", [
    [3, 1],                                                     # k6v2<4>
    [2, 1, 1, 2, 1, 3],                                         # k6v2<8>: long unblocking cascades
    [4, 1, 2, 1, 1, 3, None, 1, 2, 1, 1, 2],                    # k6v2<16>
    [3, 1, 1, 2, 1, 1, 2, 1, 1, 2, 1, 1, 2, 1, 1, 2],           # k6v2<16>, 16 nodes
], ids=["m2", "m6", "m12", "m16"])
def test_tandem_kernels_all_variants(cap):
    _check_tandem(cap, 700 + len(cap))


def test_tandem_general_kernel_on_a_small_configuration(monkeypatch):
    monkeypatch.setenv("MQ_K6_V1", "1")
    _check_tandem([2, 1, 1, 2], 77)


def test_tandem_generator_tiers_agree():
    """Generator tier: the predicated kernel and the general kernel draw the same streams and walk the same events."""
    import os

    from most_queue_b200 import _lib

    ctx = _lib.default_context()
    cap, mu = [3, 1, 2, 1, 2], [1.0, 0.9, 1.1, 0.8, 1.2]
    laws = [_lib.make_dist("Gamma", {"mu": 1.5 * x, "alpha": 1.5}) if i % 2 else _lib.make_dist("M", x)
            for i, x in enumerate(mu)]
    args = (cap, _lib.make_dist("M", 0.9), laws, 1500, 0.1, 96, 0, 5)
    _, rec2 = ctx.sim_tandem(*args, per_replication=True)
    os.environ["MQ_K6_V1"] = "1"
    try:
        _, rec1 = ctx.sim_tandem(*args, per_replication=True)
    finally:
        del os.environ["MQ_K6_V1"]
    assert np.array_equal(rec1, rec2)


def _both_kernels(env, call):
    """The same generator-tier run through the predicated kernel and (env=1) through the general kernel."""
    import os

    new = call()
    os.environ[env] = "1"
    try:
        old = call()
    finally:
        del os.environ[env]
    return new, old


def test_priority_generator_tiers_agree():
    """Draw-ahead rings, RED-accumulated sums and the trip body of k4v2 against k4_prio on the same Philox streams:
    every per-replication row is bit-equal (4 disciplines, Gamma and H2 service, finite buffer)."""
    from most_queue_b200 import _lib as L
    from most_queue_b200.random.distributions import GammaDistribution, H2Distribution

    ctx = L.default_context()
    g = [GammaDistribution.get_params_by_mean_and_cv(1.0 + 0.5 * k, 1.2) for k in range(3)]
    h = H2Distribution.get_params_by_mean_and_cv(1.4, 1.5)
    for prty, buffer in (("NP", None), ("PR", None), ("RS", 12), ("RW", None)):
        src = [L.make_dist("M", 0.35), L.make_dist("M", 0.3), L.make_dist("E", {"r": 2, "mu": 0.5})]
        srv = [L.make_dist("Gamma", g[0]), L.make_dist("H", h), L.make_dist("Gamma", g[2])]
        new, old = _both_kernels("MQ_K4_V1", lambda: ctx.sim_prio(3, 3, prty, buffer, src, srv, 4000, 2048, 5, 17,
                                                                   p_bins=16, per_replication=True))
        # k4v2's generator tier reports sums / count in the running-mean rows: equal to rounding only
        rows = L.prio_rows(16)
        soft = {L.PG_ROWS + k * rows + r for k in range(3) for r in range(L.PF_WRUN, L.PF_VRUN + 4)}
        bad = set(np.nonzero((new[1] != old[1]).any(axis=1))[0].tolist())
        assert bad <= soft, (prty, sorted(bad - soft))
        np.testing.assert_allclose(new[1], old[1], rtol=1e-11, atol=0)
        np.testing.assert_allclose(new[0], old[0], rtol=1e-11, atol=0)


def test_network_generator_tiers_agree():
    from most_queue_b200 import _lib as L
    from most_queue_b200.random.distributions import GammaDistribution, H2Distribution

    ctx = L.default_context()
    Rm = [[1, 0, 0, 0, 0, 0], [0, 0.4, 0.6, 0, 0, 0], [0, 0, 0.2, 0.4, 0.4, 0], [0, 0, 0, 0, 1, 0],
          [0, 0, 0, 0, 1, 0], [0, 0, 0, 0, 0, 1]]
    ch = [3, 2, 3, 4, 3]
    srv = [L.make_dist("H", H2Distribution.get_params_by_mean_and_cv(0.7 * c, 1.2)) if i % 2 == 0 else
           L.make_dist("Gamma", GammaDistribution.get_params_by_mean_and_cv(0.7 * c, 0.8)) for i, c in enumerate(ch)]
    new, old = _both_kernels("MQ_K5_V1", lambda: ctx.sim_net(Rm, ch, L.make_dist("M", 1.0), srv, 1500, 2048, 3, 23,
                                                              per_replication=True))
    # the predicated kernel does not track the reference's running means in the generator tier (the aggregate is formed
    # from the power sums): every other per-replication row is bit-equal
    soft = set(range(L.NG_VRUN, L.NG_VRUN + 3)) | set(range(L.NG_WRUN, L.NG_WRUN + 3))
    for i in range(5):
        b = L.NG_ROWS + i * L.NN_ROWS
        soft |= set(range(b + L.NN_VRUN, b + L.NN_VRUN + 4)) | set(range(b + L.NN_WRUN, b + L.NN_WRUN + 4))
    bad = set(np.nonzero((new[1] != old[1]).any(axis=1))[0].tolist())
    assert bad <= soft, sorted(bad - soft)
    np.testing.assert_allclose(new[0], old[0], rtol=1e-11, atol=0)


def test_priority_network_generator_tiers_agree():
    from most_queue_b200 import _lib as L
    from most_queue_b200.random.distributions import H2Distribution

    ctx = L.default_context()
    R0 = [[1, 0, 0, 0], [0, 0, 0.7, 0.3], [0, 0, 0, 1], [0, 0, 0, 1]]
    R1 = [[0.5, 0.5, 0, 0], [0, 0, 0, 1], [0.2, 0, 0.3, 0.5], [0, 0, 0, 1]]
    srv = [[L.make_dist("H", H2Distribution.get_params_by_mean_and_cv(0.5 + 0.2 * j + 0.1 * i, 1.3)) for j in range(2)]
           for i in range(3)]
    new, old = _both_kernels("MQ_K7_V1", lambda: ctx.sim_prionet(
        [R0, R1], [2, 1, 2], ["PR", "NP", "RS"], [[0, 1], [1, 0], [0, 1]],
        [L.make_dist("M", 0.6), L.make_dist("M", 0.7)], srv, 1500, 2048, 2, 11, per_replication=True))
    m, K = 3, 2
    soft = set()
    for k in range(K):
        b = L.PN_G_ROWS + k * L.PNC_ROWS
        soft |= set(range(b + L.PNC_VRUN, b + L.PNC_VRUN + 3)) | set(range(b + L.PNC_WRUN, b + L.PNC_WRUN + 3))
    for nk in range(m * K):
        b = L.PN_G_ROWS + K * L.PNC_ROWS + nk * L.PNN_ROWS
        soft |= set(range(b + L.PNN_VRUN, b + L.PNN_VRUN + 4)) | set(range(b + L.PNN_WRUN, b + L.PNN_WRUN + 4))
    bad = set(np.nonzero((new[1] != old[1]).any(axis=1))[0].tolist())
    assert bad <= soft, sorted(bad - soft)
    np.testing.assert_allclose(new[0], old[0], rtol=1e-11, atol=0)
