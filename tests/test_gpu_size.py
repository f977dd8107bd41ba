"""K10 (size-based single-server kernel) through the C ABI: replay tier bit-exact against the reference's golden
fixtures and the oracle; generator tier against the oracle and the M/M/1 SRPT-vs-FCFS ordering / M/G/1 SJF formula."""

import glob
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

CASES = sorted(glob.glob(os.path.join(os.path.dirname(__file__), "golden", "size_*.npz")))
COUNTS = ("busy_n", "arrived", "taked", "served", "dropped", "in_sys", "queued")


@pytest.mark.parametrize("path", CASES, ids=[os.path.basename(p)[5:-4] for p in CASES])
def test_replay_matches_reference(path):
    from most_queue_b200.sim.size_based import replay

    g = np.load(path)
    buffer = None if int(g["buffer"]) < 0 else int(g["buffer"])
    rep = lambda a: np.repeat(np.asarray(a)[:, None], 3, axis=1)
    Y = rep(g["Y"]) if "Y" in g.files else None
    out = replay(str(g["discipline"]), buffer, rep(g["A"]), rep(g["S"]), int(g["total_served"]), p_bins=len(g["p_time"]),
                 Y=Y)
    sd = g["slowdown"]
    s1 = s2 = 0.0
    for x in sd:
        s1 += float(x)
        s2 += float(x) * float(x)
    for r in range(3):
        assert out["ttek"][r] == float(g["ttek"])
        for k in COUNTS:
            assert out[k][r] == int(g[k]), k
        assert (out["a_used"][r], out["s_used"][r]) == (int(g["a_used"]), int(g["s_used"]))
        assert np.array_equal(out["w"][r], g["w"]) and np.array_equal(out["v"][r], g["v"])
        assert np.array_equal(out["busy"][r], g["busy"]) and np.array_equal(out["p_time"][r], g["p_time"])
        assert out["slowdown_sum"][r].tolist() == [s1, s2]
        if Y is not None:
            assert out["y_used"][r] == int(g["y_used"])


@pytest.mark.parametrize("discipline,buffer,noisy", [
    ("FCFS", None, False), ("SJF", None, False), ("PSJF", None, False), ("SRPT", None, False), ("SRPT", 4, False),
    ("SJF", 3, False), ("PSPJF", None, False), ("SPJF", None, True), ("PSPJF", 5, True), ("SPRPT", None, True)])
def test_replay_matches_oracle_many_replications(discipline, buffer, noisy):
    from most_queue_b200.sim.size_based import replay
    from oracle import oracle

    rng = np.random.default_rng(11)
    reps, n = 40, 500
    A = rng.exponential(1.0, (3 * n, reps))
    S = rng.gamma(0.6, 1.4, (3 * n, reps))
    Y = S * np.exp(0.8 * rng.standard_normal(S.shape)) if noisy else None
    out = replay(discipline, buffer, A, S, n, p_bins=64, Y=Y)
    for r in range(reps):
        o = oracle.size_replay(discipline, buffer, A[:, r].copy(), S[:, r].copy(), n, p_bins=64,
                               Y=None if Y is None else Y[:, r].copy())
        assert o.ttek == out["ttek"][r]
        for k in COUNTS + ("preemptions",):
            assert o.counts[k] == out[k][r], k
        assert np.array_equal(o.w, out["w"][r]) and np.array_equal(o.v, out["v"][r])
        assert np.array_equal(o.w_sum, out["w_sum"][r]) and np.array_equal(o.v_sum, out["v_sum"][r])
        assert np.array_equal(o.p_time, out["p_time"][r]) and np.array_equal(o.slowdown_sum, out["slowdown_sum"][r])
        assert (o.a_used, o.s_used) == (out["a_used"][r], out["s_used"][r])


def test_errors_and_api_validation():
    from most_queue_b200 import _lib
    from most_queue_b200.sim.size_based import PerfectSimPredictor, SizeBasedQsSim, replay

    rng = np.random.default_rng(2)
    with pytest.raises(_lib.MqError) as e:
        replay("SRPT", None, rng.exponential(1.0, (100, 2)), rng.exponential(0.5, (100, 2)), 400)
    assert e.value.code == _lib.MQ_E_INPUT_SHORT
    with pytest.raises(ValueError):
        SizeBasedQsSim(2, "SRPT")
    with pytest.raises(ValueError):
        SizeBasedQsSim(1, "LIFO")
    from most_queue_b200.sim.size_based import ExpNoiseSimPredictor, LognormalNoiseSimPredictor

    q = SizeBasedQsSim(1, "SPJF")
    q.set_predictor(None)
    q.set_predictor(PerfectSimPredictor())
    q.set_predictor(ExpNoiseSimPredictor())
    q.set_predictor(LognormalNoiseSimPredictor(0.3))
    assert (q._pred_kind, q._pred_sigma) == ("lognormal", 0.3)
    with pytest.raises(ValueError):
        LognormalNoiseSimPredictor(0.0)
    with pytest.raises(NotImplementedError):
        q.set_predictor(object())


def test_generator_matches_oracle():
    from most_queue_b200 import _lib
    from most_queue_b200.random.distributions import H2Distribution
    from oracle import oracle

    L = _lib
    h2 = H2Distribution.get_params_by_mean_and_cv(0.75, 1.6)
    ctx = L.default_context()
    seed, reps, n = 41, 32, 1500
    for name, pred in (("SJF", "perfect"), ("SRPT", "perfect"), ("SPJF", "exp"), ("SPRPT", "lognormal")):
        agg, rec = ctx.sim_size(L.SIZE_DISCIPLINES[name], None, L.make_dist("M", 1.0), L.make_dist("H", h2), n, reps, 2,
                                seed, p_bins=64, per_replication=True, predictor=pred, sigma=0.6)
        same = 0
        for r in range(reps):
            o = oracle.size_generate(name, None, ("M", 1.0), ("H", h2), n, L.seed_to_key(seed), 2 + r, p_bins=64,
                                     predictor=pred, sigma=0.6)
            assert int(rec[L.Z_SERVED, r]) == n
            if int(rec[L.Z_TAKED, r]) == o.counts["taked"] and int(rec[L.Z_ARRIVED, r]) == o.counts["arrived"]:
                same += 1
                np.testing.assert_allclose(rec[L.Z_VSUM, r], o.v_sum[0], rtol=3e-4)
                np.testing.assert_allclose(rec[L.Z_TTEK, r], o.ttek, rtol=1e-5)
        assert same >= reps // 2, (name, same)


def test_mm1_disciplines_closed_forms():
    """M/M/1, rho = 0.7: FCFS E[T] = 1 / (mu - lam); non-pre-emptive SJF E[W] = integral of the Pollaczek-Khinchine
    residual over (1 - rho(x))^2 (Conway, Maxwell & Miller); SRPT must beat both."""
    from scipy import integrate

    from most_queue_b200.sim.size_based import SizeBasedQsSim

    lam, mu = 0.7, 1.0
    res = {}
    for d in ("FCFS", "SJF", "SRPT"):
        sim = SizeBasedQsSim(1, d, seed=5, replications=512, track_slowdown=True)
        sim.set_sources(lam, "M")
        sim.set_servers(mu, "M")
        res[d] = (sim.run(20000), sim)
    t_fcfs = 1.0 / (mu - lam)
    r, _ = res["FCFS"]
    assert abs(r.v[0] - t_fcfs) < r.ci["v"][0] + 5e-3 * t_fcfs
    # SJF: E[W] = int_0^inf  (lam E[S^2] / 2) / (1 - rho(x))^2 f(x) dx,  rho(x) = lam int_0^x t f(t) dt
    rho_x = lambda x: lam * (1.0 - np.exp(-mu * x) * (1.0 + mu * x)) / mu
    w_sjf, _ = integrate.quad(lambda x: (lam * 2.0 / mu**2 / 2.0) / (1.0 - rho_x(x)) ** 2 * mu * np.exp(-mu * x), 0, 60)
    r, _ = res["SJF"]
    assert abs(r.w[0] - w_sjf) < r.ci["w"][0] + 1e-2 * w_sjf, (r.w[0], w_sjf)
    assert res["SRPT"][0].v[0] < res["SJF"][0].v[0] < res["FCFS"][0].v[0]
    assert res["SRPT"][1].preemptions > 0 and res["SJF"][1].preemptions == 0
    assert res["SRPT"][1].get_slowdown() < res["SJF"][1].get_slowdown()


def test_noisy_predictions_cost_performance_but_beat_fcfs():
    """M/M/1, rho = 0.7: SPRPT with log-normal prediction noise sits between SRPT (perfect information) and FCFS."""
    from most_queue_b200.sim.size_based import LognormalNoiseSimPredictor, SizeBasedQsSim

    def run(d, pred=None):
        sim = SizeBasedQsSim(1, d, seed=6, replications=256)
        sim.set_sources(0.7, "M")
        sim.set_servers(1.0, "M")
        if pred is not None:
            sim.set_predictor(pred)
        return sim.run(20000).v[0]

    v_srpt, v_noisy, v_fcfs = run("SRPT"), run("SPRPT", LognormalNoiseSimPredictor(1.0)), run("FCFS")
    assert v_srpt < v_noisy < v_fcfs, (v_srpt, v_noisy, v_fcfs)
    assert abs(run("SPRPT") - v_srpt) < 1e-9  # the perfect predictor makes SPRPT the same schedule as SRPT
