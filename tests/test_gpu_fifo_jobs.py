"""K1b (csrc/k1b_fifo_jobs.cu), the job-indexed HBM-roofline replay tier, through the C ABI: per-job waits and sojourn
times and the arrival-order power sums equal the oracle's job-indexed mode BIT FOR BIT (that mode is pinned to the
unmodified reference by tests/test_oracle_fifo_jobs.py), and the reference's own fixtures directly."""

import glob
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

FIX = [p for p in sorted(glob.glob(os.path.join(os.path.dirname(__file__), "golden", "fifo_*.npz")))
       if int(np.load(p)["buffer"]) < 0]


@pytest.mark.parametrize("path", FIX, ids=[os.path.basename(p)[5:-4] for p in FIX])
def test_per_job_samples_equal_the_reference_fixtures(path):
    from most_queue_b200.sim.base import replay_jobs

    g = np.load(path)
    c, n = int(g["c"]), int(g["taked"])
    A = np.repeat(g["A"][:n, None], 3, axis=1)
    S = np.repeat(g["S"][:n, None], 3, axis=1)
    out = replay_jobs(c, A, S, n, want_samples=True)
    for r in range(3):
        assert np.array_equal(out["W"][:, r], g["w_samples"][:n])     # wait samples in start (= arrival) order
        left = np.cumsum(g["A"][:n]) + out["V"][:, r] <= float(g["ttek"]) + 1e-9
        assert np.array_equal(np.sort(out["V"][left, r]), np.sort(g["v_samples"]))


@pytest.mark.parametrize("c", [1, 2, 3, 5, 8, 12, 32])
def test_many_replications_bitwise_vs_oracle(c):
    from most_queue_b200.sim.base import replay_jobs
    from oracle import oracle

    rng = np.random.default_rng(100 + c)
    n, R = 1203, 70                       # not a multiple of the prefetch depth
    A = rng.exponential(1.0, (n, R))
    S = rng.gamma(0.7, 0.9 * c / 0.7, (n, R)) * rng.uniform(0.7, 1.1, (1, R))   # loads from 0.63 to 0.99
    out = replay_jobs(c, A, S, want_samples=True)
    for r in range(R):
        o = oracle.fifo_jobs(c, A[:, r].copy(), S[:, r].copy())
        assert np.array_equal(out["W"][:, r], o["w"]) and np.array_equal(out["V"][:, r], o["v"])
        assert np.array_equal(out["w_sum"][r], o["w_sum"]) and np.array_equal(out["v_sum"][r], o["v_sum"])
        assert out["t_last"][r] == o["t_last"] and out["d_max"][r] == o["d_max"] and out["waited"][r] == o["waited"]
    # the sums without sample output take the other instantiation: same bits
    out2 = replay_jobs(c, A, S)
    assert np.array_equal(out2["w_sum"], out["w_sum"]) and np.array_equal(out2["v_sum"], out["v_sum"])
    agg = out2["agg"]
    np.testing.assert_allclose(agg[0], (out["w_sum"][:, 0] / n).sum(), rtol=1e-13)


def test_integer_min_max_variant_and_event_loop_kernel_agree():
    """MQ_K1B_INT=1 (64-bit integer compares on the bit patterns) gives the same bits; and the samples are the ones K1,
    the event-loop replay kernel, accumulates: its running mean of w over the jobs it started equals the mean of the
    first `taked` per-job waits to rounding."""
    from most_queue_b200.sim.base import replay, replay_jobs

    rng = np.random.default_rng(3)
    n, R = 2000, 96
    A = rng.exponential(1.0, (n + 100, R))
    S = rng.exponential(2.6, (n + 100, R))
    a = replay_jobs(3, A, S, n)
    os.environ["MQ_K1B_INT"] = "1"
    try:
        b = replay_jobs(3, A, S, n)
    finally:
        os.environ.pop("MQ_K1B_INT")
    assert np.array_equal(a["w_sum"], b["w_sum"]) and np.array_equal(a["v_sum"], b["v_sum"])
    ev = replay(3, None, A, S, n - 50, p_bins=16)
    full = replay_jobs(3, A, S, n, want_samples=True)
    for r in range(0, R, 7):
        k = int(ev["taked"][r])
        np.testing.assert_allclose(ev["w"][r][0], full["W"][:k, r].mean(), rtol=1e-12)


def test_errors():
    from most_queue_b200 import _lib

    ctx = _lib.default_context()
    with pytest.raises(ValueError):
        ctx.replay_fifo_jobs(0, np.ones((4, 2)), np.ones((4, 2)))
    with pytest.raises(ValueError):
        ctx.replay_fifo_jobs(2, np.ones((4, 2)), np.ones((4, 3)))
