"""The CPU oracle must reproduce the reference QsSim bit-for-bit on the golden fixtures
(tests/golden/fifo_*.npz, produced by tests/golden/make_golden.py from the unmodified
reference with replayed variate arrays)."""

import glob
import os

import numpy as np
import pytest

from oracle import oracle

CASES = sorted(glob.glob(os.path.join(os.path.dirname(__file__), "golden", "fifo_*.npz")))


def test_fixtures_present():
    assert len(CASES) >= 9


@pytest.mark.parametrize("path", CASES, ids=[os.path.basename(p)[5:-4] for p in CASES])
def test_oracle_matches_reference_bitwise(path):
    g = np.load(path)
    buf = int(g["buffer"])
    r = oracle.fifo_replay(int(g["c"]), None if buf < 0 else buf, g["A"], g["S"],
                           int(g["total_served"]), p_bins=len(g["p_time"]), samples=True)
    for k in ("arrived", "taked", "served", "dropped", "in_sys", "queued", "a_used", "s_used", "busy_n"):
        assert getattr(r, k) == int(g[k]), k
    assert r.ttek == float(g["ttek"])
    # per-job samples in the reference's own order
    assert np.array_equal(r.w_samples, g["w_samples"])
    assert np.array_equal(r.v_samples, g["v_samples"])
    # the reference's running-mean moments, bit for bit
    assert np.array_equal(r.w, g["w"])
    assert np.array_equal(r.v, g["v"])
    assert np.array_equal(r.busy, g["busy"])
    # time-in-state sums
    assert np.array_equal(r.p_time, g["p_time"])
    assert r.p_over == 0.0
    # plain power sums agree with the running means to rounding
    np.testing.assert_allclose(r.w_sum / max(r.taked, 1), g["w"], rtol=1e-11, atol=1e-300)
    np.testing.assert_allclose(r.v_sum / r.served, g["v"], rtol=1e-11)
