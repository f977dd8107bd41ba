"""The CPU oracle must reproduce the reference PriorityNetwork bit-for-bit on the golden fixtures
(tests/golden/prionet_*.npz: unmodified reference, replayed per-class external gaps, routing uniforms and
per-(node, node-class) service draws, free-server choice pinned to the lowest index)."""

import glob
import os

import numpy as np
import pytest

from oracle import oracle

CASES = sorted(glob.glob(os.path.join(os.path.dirname(__file__), "golden", "prionet_*.npz")))


def load_case(g):
    m, K = len(g["channels"]), g["R"].shape[0]
    A = [g[f"A{k}"] for k in range(K)]
    S = [[g[f"S{i}_{j}"] for j in range(K)] for i in range(m)]
    return m, K, A, S


def test_fixtures_present():
    assert len(CASES) >= 5


@pytest.mark.parametrize("path", CASES, ids=[os.path.basename(p)[8:-4] for p in CASES])
def test_oracle_matches_reference_bitwise(path):
    g = np.load(path)
    m, K, A, S = load_case(g)
    r = oracle.prionet_replay(g["R"], g["channels"], [str(x) for x in g["prty"]], g["nodes_prty"], A, g["U"], S,
                              int(g["job_served"]))
    assert r.ttek == float(g["ttek"])
    assert np.array_equal(r.served, g["served"]) and np.array_equal(r.arrived, g["arrived"])
    assert np.array_equal(r.a_used, g["a_used"]) and r.u_used == int(g["u_used"])
    assert np.array_equal(r.s_used, g["s_used"])
    assert np.array_equal(r.v, g["v"]) and np.array_equal(r.w, g["w"])
    assert np.array_equal(r.node_v, g["node_v"]) and np.array_equal(r.node_w, g["node_w"])
    assert np.array_equal(r.node_served, g["node_served"]) and np.array_equal(r.node_taked, g["node_taked"])
