"""The CPU oracle must reproduce the reference's VacationQueueingSystemSimulator / NPolicyQueueSim bit-for-bit on
the golden fixtures (tests/golden/vac_*.npz: unmodified reference with replayed draws for the arrival, service,
warm-service, warm-up, cold and cold-delay streams)."""

import glob
import os

import numpy as np
import pytest

from oracle import oracle

CASES = sorted(glob.glob(os.path.join(os.path.dirname(__file__), "golden", "vac_*.npz")))


def streams_of(g):
    return {k: g["s_" + k] for k in oracle.VAC_STREAMS if "s_" + k in g.files}


def test_fixtures_present():
    assert len(CASES) >= 8


@pytest.mark.parametrize("path", CASES, ids=[os.path.basename(p)[4:-4] for p in CASES])
def test_oracle_matches_reference_bitwise(path):
    g = np.load(path)
    buffer = None if int(g["buffer"]) < 0 else int(g["buffer"])
    r = oracle.vac_replay(int(g["c"]), buffer, streams_of(g), int(g["total_served"]),
                          service_on_warm_up=bool(g["service_on_warm_up"]),
                          multiple_vacations=bool(g["multiple_vacations"]), big_n=int(g["big_n"]),
                          p_bins=len(g["p_time"]))
    assert r.ttek == float(g["ttek"])
    assert (r.arrived, r.taked, r.served, r.dropped, r.in_sys, r.queued) == tuple(
        int(g[k]) for k in ("arrived", "taked", "served", "dropped", "in_sys", "queued"))
    assert np.array_equal(r.w, g["w"]) and np.array_equal(r.v, g["v"])
    assert r.busy_n == int(g["busy_n"]) and np.array_equal(r.busy, g["busy"])
    assert np.array_equal(r.p_time, g["p_time"])
    assert (r.warm_prob, r.cold_prob, r.delay_prob) == (float(g["warm_prob"]), float(g["cold_prob"]),
                                                        float(g["delay_prob"]))
    assert list(r.starts) == g["starts"].tolist() and r.warm_after_cold == int(g["warm_after_cold"])
    for k in streams_of(g):
        assert r.used[k] == int(g["used_" + k]), k
