"""K8 (vacation / N-policy queue kernel) through the C ABI: replay tier bit-exact against the reference's golden
fixtures and the oracle; generator tier against the oracle driven by the same Philox words and against closed
forms (M/G/1 with multiple vacations, N-policy)."""

import glob
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

CASES = sorted(glob.glob(os.path.join(os.path.dirname(__file__), "golden", "vac_*.npz")))
STREAMS = ("arrivals", "services", "warm_services", "warm", "cold", "delay")


def _streams(g):
    return {k: g["s_" + k] for k in STREAMS if "s_" + k in g.files}


@pytest.mark.parametrize("path", CASES, ids=[os.path.basename(p)[4:-4] for p in CASES])
def test_replay_matches_reference(path):
    from most_queue_b200.sim.vacations import replay

    g = np.load(path)
    buffer = None if int(g["buffer"]) < 0 else int(g["buffer"])
    rep = lambda a: np.repeat(np.asarray(a)[:, None], 3, axis=1)
    out = replay(int(g["c"]), buffer, {k: rep(v) for k, v in _streams(g).items()}, int(g["total_served"]),
                 is_service_on_warm_up=bool(g["service_on_warm_up"]),
                 is_multiple_vacations=bool(g["multiple_vacations"]), big_n=int(g["big_n"]), p_bins=len(g["p_time"]))
    for r in range(3):
        assert out["ttek"][r] == float(g["ttek"])
        for k in ("arrived", "taked", "served", "dropped", "in_sys", "queued", "busy_n", "warm_after_cold"):
            assert out[k][r] == int(g[k]), k
        assert np.array_equal(out["w"][r], g["w"]) and np.array_equal(out["v"][r], g["v"])
        assert np.array_equal(out["busy"][r], g["busy"])
        assert np.array_equal(out["p_time"][r], g["p_time"])
        assert out["warm_prob"][r] == float(g["warm_prob"]) and out["cold_prob"][r] == float(g["cold_prob"])
        assert out["delay_prob"][r] == float(g["delay_prob"])
        assert out["starts"][r].tolist() == g["starts"].tolist()
        for k in _streams(g):
            assert out["used"][k][r] == int(g["used_" + k]), k


@pytest.mark.parametrize("cfg", [
    dict(c=1, buffer=None, use=("warm",)),
    dict(c=2, buffer=None, use=("warm_services", "warm", "cold"), sow=True),
    dict(c=3, buffer=6, use=("warm", "cold", "delay")),
    dict(c=1, buffer=None, use=("cold",), multi=True),
    dict(c=2, buffer=None, use=("warm",), big_n=5),
    dict(c=6, buffer=None, use=("warm", "cold", "delay")),   # 8 server slots
    dict(c=12, buffer=20, use=("warm", "cold")),             # 32 server slots
], ids=["warm", "service_on_warm+cold", "buffer6_all", "multiple", "npolicy5_warm", "c6_all", "c12_buffer20"])
def test_replay_matches_oracle_many_replications(cfg):
    from most_queue_b200.sim.vacations import replay
    from oracle import oracle

    rng = np.random.default_rng(3)
    reps, n = 40, 400
    c = cfg["c"]
    X = {"arrivals": rng.exponential(1.0 / c, (3 * n, reps)), "services": rng.gamma(0.8, 0.9, (3 * n, reps))}
    for k in cfg["use"]:
        X[k] = rng.exponential(0.8, (3 * n, reps))
    kw = dict(is_service_on_warm_up=cfg.get("sow", False), is_multiple_vacations=cfg.get("multi", False),
              big_n=cfg.get("big_n", 0))
    out = replay(c, cfg["buffer"], X, n, p_bins=64, **kw)
    for r in range(reps):
        o = oracle.vac_replay(c, cfg["buffer"], {k: v[:, r].copy() for k, v in X.items()}, n,
                              service_on_warm_up=kw["is_service_on_warm_up"],
                              multiple_vacations=kw["is_multiple_vacations"], big_n=kw["big_n"], p_bins=64)
        assert o.ttek == out["ttek"][r] and o.arrived == out["arrived"][r] and o.dropped == out["dropped"][r]
        assert np.array_equal(o.w, out["w"][r]) and np.array_equal(o.v, out["v"][r])
        assert np.array_equal(o.w_sum, out["w_sum"][r]) and np.array_equal(o.v_sum, out["v_sum"][r])
        assert np.array_equal(o.p_time, out["p_time"][r]) and o.p_over == out["p_over"][r]
        assert (o.warm_prob, o.cold_prob, o.delay_prob) == (out["warm_prob"][r], out["cold_prob"][r],
                                                            out["delay_prob"][r])
        assert list(o.starts) == out["starts"][r].tolist()
        for k in X:
            assert o.used[k] == out["used"][k][r], k


def test_short_arrays_overflow_and_bad_configuration():
    from most_queue_b200 import _lib
    from most_queue_b200.sim.vacations import NPolicyQueueSim, VacationQueueingSystemSimulator, replay

    rng = np.random.default_rng(5)
    X = {"arrivals": rng.exponential(1.0, (900, 2)), "services": rng.exponential(0.6, (900, 2)),
         "cold": rng.exponential(1.0, (20, 2))}
    with pytest.raises(_lib.MqError) as e:
        replay(1, None, X, 300)
    assert e.value.code == _lib.MQ_E_INPUT_SHORT
    X["cold"] = rng.exponential(60.0, (900, 2))
    with pytest.raises(_lib.MqError) as e:
        replay(1, None, X, 300, queue_cap=8)
    assert e.value.code == _lib.MQ_E_OVERFLOW
    q = VacationQueueingSystemSimulator(1)
    with pytest.raises(ValueError):
        q.set_warm(1.0, "M")  # servers first (vacations.py:61)
    with pytest.raises(ValueError):
        q.set_cold_delay(1.0, "M")  # cold first (vacations.py:99)
    with pytest.raises(ValueError):
        NPolicyQueueSim(1, 0)


def test_generator_matches_oracle():
    from most_queue_b200 import _lib
    from most_queue_b200.random.distributions import GammaDistribution
    from oracle import oracle

    L = _lib
    gam = GammaDistribution.get_params_by_mean_and_cv(0.7, 0.8)
    laws = {"arrivals": ("M", 1.0), "services": ("Gamma", gam), "warm": ("M", 1.2), "cold": ("M", 0.9),
            "delay": ("M", 1.5)}
    ctx = L.default_context()
    seed, reps, n = 21, 32, 1500
    agg, rec = ctx.sim_vacation(2, None, {k: L.make_dist(*v) for k, v in laws.items()}, n, reps, 3, seed, p_bins=64,
                                per_replication=True)
    same = 0
    for r in range(reps):
        o = oracle.vac_generate(2, None, laws, n, L.seed_to_key(seed), 3 + r, p_bins=64)
        assert int(rec[L.V_SERVED, r]) == n
        if int(rec[L.V_ARRIVED, r]) == o.arrived and rec[L.V_STARTS:L.V_STARTS + 3, r].tolist() == list(o.starts):
            same += 1
            np.testing.assert_allclose(rec[L.V_VSUM, r], o.v_sum[0], rtol=3e-4)
            np.testing.assert_allclose(rec[L.V_TTEK, r], o.ttek, rtol=1e-5)
            np.testing.assert_allclose(rec[L.V_COLDPROB, r], o.cold_prob, rtol=1e-4)
    assert same >= reps // 2
    np.testing.assert_allclose(agg[16] / reps, (rec[L.V_VSUM] / rec[L.V_SERVED]).mean(), rtol=1e-12)


def test_mg1_multiple_vacations_closed_form():
    """M/G/1 with multiple exponential vacations: E[W] = lambda b2 / (2 (1 - rho)) + E[V^2] / (2 E[V])
    (decomposition, e.g. Takagi vol. 1); fraction of time on vacation = 1 - rho."""
    from most_queue_b200.sim.vacations import VacationQueueingSystemSimulator

    lam, mu, nu = 0.6, 1.0, 2.0
    rho = lam / mu
    w_theory = lam * (2 / mu**2) / (2 * (1 - rho)) + (2 / nu**2) / (2 / nu)
    sim = VacationQueueingSystemSimulator(1, is_multiple_vacations=True, seed=4, replications=512)
    sim.set_sources(lam, "M")
    sim.set_servers(mu, "M")
    sim.set_cold(nu, "M")
    res = sim.run(20000)
    assert abs(res.w[0] - w_theory) < res.ci["w"][0] + 5e-3 * w_theory, (res.w[0], w_theory)
    assert abs(res.cold_prob - (1 - rho)) < 5e-3, res.cold_prob
    assert res.warmup_prob == 0.0 and abs(sum(res.p) - 1.0) < 1e-6


def test_npolicy_closed_form():
    """M/M/1 under N-policy: E[W] = rho / (mu (1 - rho)) + (N - 1) / (2 lambda)."""
    from most_queue_b200.sim.vacations import NPolicyQueueSim

    lam, mu, N = 0.5, 1.0, 4
    rho = lam / mu
    w_theory = rho / (mu * (1 - rho)) + (N - 1) / (2 * lam)
    sim = NPolicyQueueSim(1, N, seed=8, replications=512)
    sim.set_sources(lam, "M")
    sim.set_servers(mu, "M")
    res = sim.run(20000)
    assert abs(res.w[0] - w_theory) < res.ci["w"][0] + 5e-3 * w_theory, (res.w[0], w_theory)
