"""K1 (replay kernel, fp64) against the golden fixtures of the reference and against the CPU
oracle -- bit for bit, through the C ABI (mq_replay_fifo)."""

import glob
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

CASES = sorted(glob.glob(os.path.join(os.path.dirname(__file__), "golden", "fifo_*.npz")))


@pytest.mark.parametrize("path", CASES, ids=[os.path.basename(p)[5:-4] for p in CASES])
def test_replay_matches_reference_bitwise(path):
    from most_queue_b200.sim.base import replay

    g = np.load(path)
    buf = int(g["buffer"])
    nb = len(g["p_time"])
    # replicate the same input over 3 lanes: every lane must give the reference's answer
    A = np.repeat(g["A"][:, None], 3, axis=1)
    S = np.repeat(g["S"][:, None], 3, axis=1)
    out = replay(int(g["c"]), None if buf < 0 else buf, A, S, int(g["total_served"]), p_bins=nb)
    for r in range(3):
        for k in ("arrived", "taked", "served", "dropped", "in_sys", "queued", "a_used", "s_used", "busy_n"):
            assert int(out[k][r]) == int(g[k]), k
        assert out["ttek"][r] == float(g["ttek"])
        assert np.array_equal(out["w"][r], g["w"])      # the reference's running-mean moments
        assert np.array_equal(out["v"][r], g["v"])
        assert np.array_equal(out["busy"][r], g["busy"])
        assert np.array_equal(out["p_time"][r], g["p_time"])


@pytest.mark.parametrize("c,buffer", [(1, None), (2, 5), (3, None), (4, 0), (8, None), (8, 12), (12, None), (20, 3)])
def test_replay_matches_oracle_many_replications(c, buffer):
    """Distinct inputs per lane (ragged consumption), compared with the CPU oracle bitwise,
    including the plain power sums accumulated in the reference's sample order."""
    from most_queue_b200.sim.base import replay
    from oracle import oracle

    rng = np.random.default_rng(100 + c)
    R, n = 70, 400
    rows = 3 * n
    rho = rng.uniform(0.5, 0.95, R)
    A = rng.exponential(1.0, (rows, R))
    S = rng.gamma(0.7, 1.0 / 0.7, (rows, R)) * (rho * c)[None, :]
    out = replay(c, buffer, A, S, n, p_bins=40)
    for r in range(R):
        o = oracle.fifo_replay(c, buffer, A[:, r].copy(), S[:, r].copy(), n, p_bins=40)
        assert (o.arrived, o.taked, o.served, o.dropped) == (
            int(out["arrived"][r]), int(out["taked"][r]), int(out["served"][r]), int(out["dropped"][r]))
        assert o.ttek == out["ttek"][r]
        assert np.array_equal(o.w, out["w"][r]) and np.array_equal(o.v, out["v"][r])
        assert np.array_equal(o.w_sum, out["w_sum"][r]) and np.array_equal(o.v_sum, out["v_sum"][r])
        assert np.array_equal(o.p_time, out["p_time"][r]) and o.p_over == out["p_over"][r]
        assert np.array_equal(o.busy, out["busy"][r]) and o.busy_n == int(out["busy_n"][r])


def test_replay_input_too_short_raises():
    from most_queue_b200 import _lib
    from most_queue_b200.sim.base import replay

    A = np.ones((10, 2))
    S = np.ones((10, 2))
    with pytest.raises(_lib.MqError):
        replay(1, None, A, S, 50)
