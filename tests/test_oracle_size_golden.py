"""The CPU oracle must reproduce the reference's SizeBasedQsSim (FCFS, SJF, PSJF, SRPT and, with the default perfect
predictor, SPJF / PSPJF / SPRPT) bit-for-bit on the golden fixtures (tests/golden/size_*.npz: unmodified reference with
replayed gaps and job sizes)."""

import glob
import os

import numpy as np
import pytest

from oracle import oracle

CASES = sorted(glob.glob(os.path.join(os.path.dirname(__file__), "golden", "size_*.npz")))
COUNTS = ("busy_n", "arrived", "taked", "served", "dropped", "in_sys", "queued")


def test_fixtures_present():
    assert len(CASES) >= 12


@pytest.mark.parametrize("path", CASES, ids=[os.path.basename(p)[5:-4] for p in CASES])
def test_oracle_matches_reference_bitwise(path):
    g = np.load(path)
    buffer = None if int(g["buffer"]) < 0 else int(g["buffer"])
    Y = g["Y"] if "Y" in g.files else None
    r = oracle.size_replay(str(g["discipline"]), buffer, g["A"], g["S"], int(g["total_served"]), p_bins=len(g["p_time"]),
                           Y=Y)
    assert r.ttek == float(g["ttek"])
    for k in COUNTS:
        assert r.counts[k] == int(g[k]), k
    assert (r.a_used, r.s_used) == (int(g["a_used"]), int(g["s_used"]))
    assert np.array_equal(r.w, g["w"]) and np.array_equal(r.v, g["v"]) and np.array_equal(r.busy, g["busy"])
    assert np.array_equal(r.p_time, g["p_time"])
    sd = g["slowdown"]
    if len(sd):  # the reference keeps the samples; sums in completion order
        s1 = s2 = 0.0
        for x in sd:
            s1 += float(x)
            s2 += float(x) * float(x)
        assert r.slowdown_sum[0] == s1 and r.slowdown_sum[1] == s2
