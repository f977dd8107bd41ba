"""The CPU oracle must reproduce the reference NetworkSimulator bit-for-bit on the golden fixtures
(tests/golden/net_*.npz: unmodified reference with replayed external gaps, routing uniforms and
per-node service draws)."""

import glob
import os

import numpy as np
import pytest

from oracle import oracle

CASES = sorted(glob.glob(os.path.join(os.path.dirname(__file__), "golden", "net_*.npz")))


def test_fixtures_present():
    assert len(CASES) >= 4


@pytest.mark.parametrize("path", CASES, ids=[os.path.basename(p)[4:-4] for p in CASES])
def test_oracle_matches_reference_bitwise(path):
    g = np.load(path)
    m = len(g["channels"])
    r = oracle.net_replay(g["R"], g["channels"], g["A"], g["U"], [g[f"S{i}"] for i in range(m)],
                          int(g["job_served"]), samples=True)
    assert (r.served, r.arrived, r.in_sys) == (int(g["served"]), int(g["arrived"]), int(g["in_sys"]))
    assert r.ttek == float(g["ttek"])
    assert (r.a_used, r.u_used) == (int(g["a_used"]), int(g["u_used"]))
    assert np.array_equal(r.s_used, g["s_used"])
    assert np.array_equal(r.v_samples, g["v_samples"]) and np.array_equal(r.w_samples, g["w_samples"])
    assert np.array_equal(r.v, g["v"]) and np.array_equal(r.w, g["w"])  # math.pow running moments
    assert np.array_equal(r.node_v, g["node_v"]) and np.array_equal(r.node_w, g["node_w"])
    assert np.array_equal(r.node_served, g["node_served"])
    assert np.array_equal(r.node_taked, g["node_taked"])
    assert np.array_equal(r.node_arrived, g["node_arrived"])
