"""BASELINE cfg2 at full size (1e6 replications x 1e5 jobs = 1e11 simulated jobs): properties that do
not need the oracle -- conservation of jobs, Little's law as an exact sample-path identity, the
utilisation law, determinism and agreement with Takahashi-Takami within the CI plus start-up bias."""

import json
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

H2 = dict(p1=0.447586, mu1=0.0950716, mu2=0.619214)


@pytest.fixture(scope="module")
def full_run():
    from most_queue_b200 import _lib
    from most_queue_b200.random.utils.params import H2Params

    ctx = _lib.default_context()
    src, srv = _lib.make_dist("M", 1.0), _lib.make_dist("H", H2Params(**H2))
    R, N, PB = 1_000_000, 100_000, 128
    agg, rec = ctx.sim_fifo(8, None, src, srv, N, R, 0, 20260101, p_bins=PB, per_replication=True)
    return _lib, agg, rec, R, N, PB, ctx, src, srv


def test_conservation_and_stop_rule(full_run):
    L, agg, rec, R, N, PB, *_ = full_run
    assert agg[L.A_REPS] == R and agg[L.A_SERVED] == R * N and agg[L.A_DROPPED] == 0
    assert np.all(rec[L.F_SERVED] == N)
    assert np.all(rec[L.F_TAKED] >= N) and np.all(rec[L.F_TAKED] <= N + 8)   # <= c jobs in service at the stop
    assert np.all(rec[L.F_ARRIVED] >= rec[L.F_TAKED])
    assert np.all(rec[L.F_FLAGS] == 0)
    # the time-in-state sums tile the whole run
    tot = rec[L.F_P:L.F_P + PB].sum(axis=0) + rec[L.F_POVER]
    np.testing.assert_allclose(tot, rec[L.F_TTEK], rtol=2e-5)


def test_littles_law_sample_path_identity(full_run):
    """integral of N(t) dt = sum of sojourns of served jobs + time accrued by the jobs still inside:
    per replication the two sides differ by at most (jobs inside at the stop) x (their age)."""
    L, agg, rec, R, N, PB, *_ = full_run
    j = np.arange(PB)[:, None]
    area = (j * rec[L.F_P:L.F_P + PB]).sum(axis=0)        # states >= PB are ~never reached at rho = 0.7
    sum_v = rec[L.F_V]
    inside = rec[L.F_ARRIVED] - rec[L.F_SERVED]
    assert np.all(area >= sum_v * (1 - 1e-4))
    assert np.all(area - sum_v <= (inside + 1) * 400.0)    # ages are bounded by a few long services
    rel = (area - sum_v) / sum_v
    assert rel.mean() < 2e-3


def test_utilisation_law_and_theory(full_run):
    L, agg, rec, R, N, PB, *_ = full_run
    p = agg[L.A_P:L.A_P + PB] / R
    busy = np.minimum(np.arange(PB), 8) @ p              # mean number of busy servers
    b1 = H2["p1"] / H2["mu1"] + (1 - H2["p1"]) / H2["mu2"]
    assert abs(busy - 1.0 * b1) < 2e-3 * b1               # lambda * E[S], start-up effects ~1e-4
    with open(os.path.join(os.path.dirname(__file__), "golden", "theory.json")) as f:
        th = json.load(f)["mh2_8"]
    w1, v1 = agg[L.A_W] / R, agg[L.A_V] / R
    ci_w = 2.5758 * np.sqrt(max(agg[L.A_W2] / R - w1 * w1, 0) / R)
    # QsSim starts empty and keeps every job: bias ~ -41/N on the mean wait
    assert abs(w1 - th["w"][0]) < ci_w + 60.0 / N
    assert abs(v1 - th["v"][0]) < 3e-4 * th["v"][0] + 60.0 / N
    assert np.allclose(p[:8], th["p"][:8], atol=2e-4)


def test_deterministic_rerun(full_run):
    L, agg, rec, R, N, PB, ctx, src, srv = full_run
    agg2, _ = ctx.sim_fifo(8, None, src, srv, N, R, 0, 20260101, p_bins=PB)
    assert np.array_equal(agg, agg2)
