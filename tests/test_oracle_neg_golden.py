"""The CPU oracle must reproduce the reference's QsSimNegatives (DISASTER / CLEAR_SYSTEM and RCS / REMOVE) bit-for-bit
on the golden fixtures (tests/golden/neg_*.npz: unmodified reference with replayed positive gaps, negative gaps,
service draws and -- for RCS -- server choices)."""

import glob
import os

import numpy as np
import pytest

from oracle import oracle

CASES = sorted(glob.glob(os.path.join(os.path.dirname(__file__), "golden", "neg_*.npz")))
COUNTS = ("positive_arrived", "negative_arrived", "taked", "served", "broken", "total", "dropped", "in_sys", "queued")


def streams_of(g):
    return {k: g["s_" + k] for k in oracle.NEG_STREAMS if "s_" + k in g.files}


def test_fixtures_present():
    assert len(CASES) >= 12


@pytest.mark.parametrize("path", CASES, ids=[os.path.basename(p)[4:-4] for p in CASES])
def test_oracle_matches_reference_bitwise(path):
    g = np.load(path)
    buffer = None if int(g["buffer"]) < 0 else int(g["buffer"])
    r = oracle.neg_replay(int(g["c"]), buffer, str(g["neg_type"]), streams_of(g), int(g["total_served"]),
                          p_bins=len(g["p_time"]), scenario=str(g["scenario"]))
    assert r.ttek == float(g["ttek"])
    for k in COUNTS:
        assert r.counts[k] == int(g[k]), k
    assert np.array_equal(r.w, g["w"]) and np.array_equal(r.v, g["v"])
    assert np.array_equal(r.v_served, g["v_served"]) and np.array_equal(r.v_broken, g["v_broken"])
    assert np.array_equal(r.p_time, g["p_time"])
    for k in streams_of(g):
        assert r.used[k] == int(g["used_" + k]), k
