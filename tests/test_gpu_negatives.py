"""K9 (negative-arrivals kernel) through the C ABI: replay tier bit-exact against the reference's golden fixtures and
the oracle; generator tier against the oracle driven by the same Philox words and against the M/M/1 disaster closed
form."""

import glob
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

CASES = sorted(glob.glob(os.path.join(os.path.dirname(__file__), "golden", "neg_*.npz")))
STREAMS = ("positives", "negatives", "services", "choices")
COUNTS = ("positive_arrived", "negative_arrived", "taked", "served", "broken", "total", "dropped", "in_sys", "queued")


def _streams(g):
    return {k: g["s_" + k] for k in STREAMS if "s_" + k in g.files}


@pytest.mark.parametrize("path", CASES, ids=[os.path.basename(p)[4:-4] for p in CASES])
def test_replay_matches_reference(path):
    from most_queue_b200.sim.negative import replay

    g = np.load(path)
    buffer = None if int(g["buffer"]) < 0 else int(g["buffer"])
    rep = lambda a: np.repeat(np.asarray(a)[:, None], 3, axis=1)
    out = replay(int(g["c"]), buffer, str(g["neg_type"]), {k: rep(v) for k, v in _streams(g).items()},
                 int(g["total_served"]), p_bins=len(g["p_time"]), scenario=str(g["scenario"]))
    for r in range(3):
        assert out["ttek"][r] == float(g["ttek"])
        for k in COUNTS:
            assert out[k][r] == int(g[k]), k
        assert np.array_equal(out["w"][r], g["w"]) and np.array_equal(out["v"][r], g["v"])
        assert np.array_equal(out["v_served"][r], g["v_served"]) and np.array_equal(out["v_broken"][r], g["v_broken"])
        assert np.array_equal(out["p_time"][r], g["p_time"])
        for k in _streams(g):
            assert out["used"][k][r] == int(g["used_" + k]), k


@pytest.mark.parametrize("neg_type,c,buffer,scenario", [
    ("DISASTER", 1, None, "CLEAR_SYSTEM"), ("DISASTER", 4, 5, "CLEAR_SYSTEM"), ("RCS", 1, None, "REMOVE"),
    ("RCS", 3, None, "REMOVE"), ("RCS", 2, 3, "REMOVE"), ("DISASTER", 3, None, "REQUEUE_ALL"),
    ("DISASTER", 4, 6, "RESUME_ALL"), ("DISASTER", 2, None, "REQUEUE_ALL_NO_RESAMPLING"), ("RCS", 3, None, "REQUEUE"),
    ("RCS", 2, 4, "RESUME"), ("RCS", 4, None, "REQUEUE_NO_RESAMPLING"),
    ("DISASTER", 6, None, "REQUEUE_ALL"), ("RCS", 7, 9, "REMOVE"),          # 8 server slots
    ("DISASTER", 12, None, "RESUME_ALL"), ("RCS", 11, None, "REQUEUE")])    # 32 server slots
def test_replay_matches_oracle_many_replications(neg_type, c, buffer, scenario):
    from most_queue_b200.sim.negative import replay
    from oracle import oracle

    rng = np.random.default_rng(7)
    reps, n = 40, 400
    X = {"positives": rng.exponential(0.9 / c, (4 * n, reps)), "negatives": rng.exponential(3.0, (2 * n, reps)),
         "services": rng.gamma(0.8, 1.1, (4 * n, reps)), "choices": rng.random((2 * n, reps))}
    out = replay(c, buffer, neg_type, X, n, p_bins=64, scenario=scenario)
    for r in range(reps):
        o = oracle.neg_replay(c, buffer, neg_type, {k: v[:, r].copy() for k, v in X.items()}, n, p_bins=64,
                              scenario=scenario)
        assert o.ttek == out["ttek"][r]
        for k in COUNTS:
            assert o.counts[k] == out[k][r], k
        assert np.array_equal(o.w, out["w"][r]) and np.array_equal(o.v, out["v"][r])
        assert np.array_equal(o.v_served, out["v_served"][r]) and np.array_equal(o.v_broken, out["v_broken"][r])
        assert np.array_equal(o.vs_sum, out["vs_sum"][r]) and np.array_equal(o.vb_sum, out["vb_sum"][r])
        assert np.array_equal(o.p_time, out["p_time"][r]) and o.p_over == out["p_over"][r]
        for k in X:
            assert o.used[k] == out["used"][k][r], k


def test_errors_and_unsupported_scenarios():
    from most_queue_b200 import _lib
    from most_queue_b200.sim.negative import (DisasterScenario, NegativeServiceType, QsSimNegatives, RcsScenario,
                                              replay)

    rng = np.random.default_rng(1)
    X = {"positives": rng.exponential(1.0, (900, 2)), "negatives": rng.exponential(3.0, (30, 2)),
         "services": rng.exponential(0.6, (900, 2))}
    with pytest.raises(_lib.MqError) as e:
        replay(1, None, "DISASTER", X, 400)
    assert e.value.code == _lib.MQ_E_INPUT_SHORT
    with pytest.raises(NotImplementedError):
        QsSimNegatives(1, NegativeServiceType.RCH)
    with pytest.raises(NotImplementedError):
        QsSimNegatives(1, NegativeServiceType.RCE)
    assert QsSimNegatives(1, NegativeServiceType.DISASTER, disaster_scenario=DisasterScenario.REQUEUE_ALL)._scenario == 2
    assert QsSimNegatives(2, NegativeServiceType.RCS, rcs_scenario=RcsScenario.RESUME)._scenario == 3
    with pytest.raises(RuntimeError):
        QsSimNegatives(1).run(10)


def test_generator_matches_oracle():
    from most_queue_b200 import _lib
    from most_queue_b200.random.distributions import H2Distribution
    from oracle import oracle

    L = _lib
    h2 = H2Distribution.get_params_by_mean_and_cv(0.8, 1.4)
    pos, neg, srv = ("M", 1.0), ("M", 0.3), ("H", h2)
    ctx = L.default_context()
    seed, reps, n = 31, 32, 1500
    for name in ("DISASTER", "RCS"):
        agg, rec = ctx.sim_negatives(2, None, L.NEG_TYPES[name], L.make_dist(*pos), L.make_dist(*neg),
                                     L.make_dist(*srv), n, reps, 5, seed, p_bins=64, per_replication=True)
        same = 0
        for r in range(reps):
            o = oracle.neg_generate(2, None, name, pos, neg, srv, n, L.seed_to_key(seed), 5 + r, p_bins=64)
            assert int(rec[L.N_SERVED, r]) == n
            if int(rec[L.N_TOTAL, r]) == o.counts["total"] and int(rec[L.N_POS_ARRIVED, r]) == o.counts["positive_arrived"]:
                same += 1
                np.testing.assert_allclose(rec[L.N_VSUM, r], o.v_sum[0], rtol=3e-4)
                np.testing.assert_allclose(rec[L.N_TTEK, r], o.ttek, rtol=1e-5)
        assert same >= reps // 2, (name, same)
        np.testing.assert_allclose(agg[16] / reps, (rec[L.N_VSUM] / rec[L.N_TOTAL]).mean(), rtol=1e-12)


def test_mm1_disaster_closed_form():
    """M/M/1 with disasters at rate d (system cleared): the number in system is geometric with ratio r, the smaller
    root of mu r^2 - (lam + mu + d) r + lam = 0, so E[N] = r / (1 - r) and, every job leaving either way,
    E[sojourn] = E[N] / lam."""
    from most_queue_b200.sim.negative import NegativeServiceType, QsSimNegatives

    lam, mu, d = 0.8, 1.0, 0.2
    s = lam + mu + d
    r = (s - np.sqrt(s * s - 4 * lam * mu)) / (2 * mu)
    v_theory = r / (1 - r) / lam
    sim = QsSimNegatives(1, NegativeServiceType.DISASTER, seed=12, replications=512)
    sim.set_positive_sources(lam, "M")
    sim.set_negative_sources(d, "M")
    sim.set_servers(mu, "M")
    res = sim.run(20000)
    assert abs(res.v[0] - v_theory) < res.ci["v"][0] + 5e-3 * v_theory, (res.v[0], v_theory)
    assert abs(res.p[0] - (1 - r)) < 5e-3
    assert 0.0 < res.q < 1.0 and res.v_broken[0] > 0 and res.v_served[0] > 0


def test_mm1_resume_scenario_is_work_conserving():
    """RCS / RESUME on one exponential server interrupts and resumes the job in service without losing work, so the
    sojourn time is that of the plain M/M/1 queue: 1 / (mu - lam)."""
    from most_queue_b200.sim.negative import NegativeServiceType, QsSimNegatives, RcsScenario

    lam, mu = 0.7, 1.0
    sim = QsSimNegatives(1, NegativeServiceType.RCS, rcs_scenario=RcsScenario.RESUME, seed=3, replications=512)
    sim.set_positive_sources(lam, "M")
    sim.set_negative_sources(0.5, "M")
    sim.set_servers(mu, "M")
    res = sim.run(20000)
    v_theory = 1.0 / (mu - lam)
    assert abs(res.v[0] - v_theory) < res.ci["v"][0] + 5e-3 * v_theory, (res.v[0], v_theory)
    assert res.q == pytest.approx(1.0) and sim.broken == 0
