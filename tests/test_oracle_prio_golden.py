"""The CPU oracle must reproduce the reference PriorityQueueSimulator bit-for-bit on the golden
fixtures (tests/golden/prio_*.npz: unmodified reference, replayed per-class variate arrays, free
server choice pinned to the lowest index -- see make_golden.py LowestFirstSet)."""

import glob
import os

import numpy as np
import pytest

from oracle import oracle

CASES = sorted(glob.glob(os.path.join(os.path.dirname(__file__), "golden", "prio_*.npz")))


def test_fixtures_present():
    assert len(CASES) >= 12


@pytest.mark.parametrize("path", CASES, ids=[os.path.basename(p)[5:-4] for p in CASES])
def test_oracle_matches_reference_bitwise(path):
    g = np.load(path)
    K = int(g["K"])
    buffer = int(g["buffer"]) if "buffer" in g.files else None  # finite total buffer (priority.py:437-455)
    r = oracle.prio_replay(int(g["c"]), K, str(g["prty"]), [g[f"A{k}"] for k in range(K)],
                           [g[f"S{k}"] for k in range(K)], int(g["total_served"]), buffer=buffer,
                           p_bins=g["p_time"].shape[1])
    if buffer is not None:
        assert int(np.sum(g["dropped"])) > 0  # the fixture really exercises the drop branch
    assert r.ttek == float(g["ttek"])
    for name in ("arrived", "served", "taked", "dropped", "in_sys", "a_used", "s_used"):
        assert np.array_equal(getattr(r, name), g[name]), name
    assert np.array_equal(r.t_old, g["t_old"])
    assert np.array_equal(r.w, g["w"])
    assert np.array_equal(r.v, g["v"])
    assert np.array_equal(r.p_time, g["p_time"])
