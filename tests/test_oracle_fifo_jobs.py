"""Pins the job-indexed oracle mode (oracle/mq_oracle.c mqo_fifo_jobs, the checker of csrc/k1b_fifo_jobs.cu) to the
REFERENCE: on the infinite-buffer FIFO fixtures (unmodified QsSim with replayed variates, tests/golden/make_golden.py)
the per-job waits equal the reference's wait samples in order, bit for bit, and the per-job sojourn times of the jobs
that had left by the reference's stop equal its sojourn samples as a multiset, bit for bit."""

import glob
import os

import numpy as np
import pytest

from oracle import oracle

CASES = [p for p in sorted(glob.glob(os.path.join(os.path.dirname(__file__), "golden", "fifo_*.npz")))
         if int(np.load(p)["buffer"]) < 0]


def test_fixtures_present():
    assert len(CASES) >= 5


@pytest.mark.parametrize("path", CASES, ids=[os.path.basename(p)[5:-4] for p in CASES])
def test_job_indexed_samples_equal_the_reference(path):
    g = np.load(path)
    c, A, S = int(g["c"]), g["A"], g["S"]
    n = int(g["taked"])                          # jobs that had started service when the reference stopped
    r = oracle.fifo_jobs(c, A, S, n)
    # FIFO: service starts in arrival order, so the reference's k-th wait sample belongs to job k
    assert np.array_equal(r["w"], g["w_samples"][:n])
    # the reference sampled v for the jobs that had LEFT by its stop (clock ttek): the same jobs here
    t_arr = np.cumsum(A[:n])                     # sequential fp64 sum == t_n of the recursion
    left = (t_arr + 0.0) + r["v"] <= float(g["ttek"]) + 1e-9
    assert int(left.sum()) == int(g["served"])
    assert np.array_equal(np.sort(r["v"][left]), np.sort(g["v_samples"]))
    # the sums are the kernel's association of exactly these samples
    assert r["w_sum"][0] == pytest.approx(r["w"].sum(), rel=1e-12)
    assert r["waited"] == int(np.count_nonzero(r["w"] > 0))


def test_job_indexed_mode_against_the_event_loop_oracle_random():
    """Multi-channel, heavy load, ties on a lattice: per-job samples against the event-loop oracle's sample lists."""
    rng = np.random.default_rng(5)
    for c, lattice in ((1, False), (2, True), (5, False), (8, True), (12, False)):
        n = 4000
        A = rng.exponential(1.0, n + 50) if not lattice else rng.integers(1, 4, n + 50) * 0.5
        S = rng.exponential(0.92 * c, n + 50) if not lattice else rng.integers(1, 6, n + 50) * 0.5 * c / 1.6
        ev = oracle.fifo_replay(c, None, A, S, n - 3 * c, p_bins=16, samples=True)
        r = oracle.fifo_jobs(c, A, S, ev.taked)
        assert np.array_equal(r["w"], ev.w_samples[:ev.taked])
        # every sojourn sample of the event loop is the sojourn of some job of the recursion (multiset inclusion)
        vs = np.sort(r["v"])
        idx = np.searchsorted(vs, np.sort(ev.v_samples))
        assert np.all(vs[np.minimum(idx, len(vs) - 1)] == np.sort(ev.v_samples))
