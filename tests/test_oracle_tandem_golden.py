"""The CPU oracle must reproduce the reference TandemBlockingSim bit-for-bit on the golden fixtures
(tests/golden/tandem_*.npz: unmodified reference with replayed external gaps and per-node service draws)."""

import glob
import os

import numpy as np
import pytest

from oracle import oracle

CASES = sorted(glob.glob(os.path.join(os.path.dirname(__file__), "golden", "tandem_*.npz")))


def test_fixtures_present():
    assert len(CASES) >= 5


@pytest.mark.parametrize("path", CASES, ids=[os.path.basename(p)[7:-4] for p in CASES])
def test_oracle_matches_reference_bitwise(path):
    g = np.load(path)
    cap = [None if k < 0 else int(k) for k in g["capacity"]]
    m = len(cap)
    o = oracle.tandem_replay(cap, g["A"], [g[f"S{i}"] for i in range(m)], int(g["total_served"]),
                             float(g["warmup_fraction"]))
    assert o.throughput == float(g["throughput"])
    assert o.loss_prob == float(g["loss_prob"])
    assert o.v0 == float(g["v0"])
    assert np.array_equal(np.array(o.mean_jobs[:m]), g["mean_jobs"])
    assert o.served == int(g["served"]) and o.a_used == int(g["a_used"])
    assert np.array_equal(np.array(o.s_used), g["s_used"])
