"""Times the replay form of the Lindley scan (device-resident fp64 A, S): one pass vs the three-launch path.

    python tools/bench_scan.py [log2 jobs ...]
"""
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from most_queue_b200 import _lib  # noqa: E402


def run(nr, two_pass, reps=5):
    ctx = _lib.default_context()
    g = torch.Generator(device="cuda").manual_seed(1)
    A = torch.empty(nr + 1, device="cuda", dtype=torch.float64).exponential_(1.0, generator=g)
    S = torch.empty(nr, device="cuda", dtype=torch.float64).exponential_(1.0 / 0.7, generator=g)
    torch.cuda.synchronize()
    if two_pass:
        os.environ["MQ_SCAN_TWO_PASS"] = "1"
    else:
        os.environ.pop("MQ_SCAN_TWO_PASS", None)
    ms = []
    for i in range(reps + 1):
        st, _ = ctx.scan_apply(nr, 0.0, A=A.data_ptr(), S=S.data_ptr(), device_ptrs=True)
        if i:
            ms.append(ctx.last_timing()["sim_ms"])
    os.environ.pop("MQ_SCAN_TWO_PASS", None)
    t = float(np.median(ms)) * 1e-3
    return {"jobs": nr, "two_pass": two_pass, "launches": ctx.last_timing()["launches"], "ms": t * 1e3, "ms_all": ms, "jobs_per_s": nr / t, "GBps_algorithmic": 16.0 * nr / t / 1e9,
            "w1": float(st[_lib.S_W] / st[_lib.S_TAKED])}


if __name__ == "__main__":
    logs = [int(x) for x in sys.argv[1:]] or [26, 28]
    leads = [int(x) for x in os.environ.get("LEADS", "16").split(",")]
    for lg in logs:
        for lead in leads:
            os.environ["MQ_SCAN_LEAD"] = str(lead)
            r = run(1 << lg, False)
            r["lead"] = lead
            r.pop("ms_all")
            print(json.dumps(r), flush=True)
        if not os.environ.get("NO_TWO_PASS"):
            r = run(1 << lg, True)
            r.pop("ms_all")
            print(json.dumps(r), flush=True)
