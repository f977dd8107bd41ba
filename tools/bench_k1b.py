"""Times K1b (job-indexed replay) on HBM-resident arrays: both min/max variants, several channel counts."""
import json, os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from most_queue_b200 import _lib

ctx = _lib.default_context()
reps, rows = 65536, 4096
g = torch.Generator(device="cuda").manual_seed(1)
A = torch.empty((rows, reps), device="cuda", dtype=torch.float64).exponential_(1.0, generator=g)
S = torch.empty((rows, reps), device="cuda", dtype=torch.float64).exponential_(1.0 / 5.6, generator=g)
torch.cuda.synchronize()
for c in (8, 1, 3, 16):
    for var in ("fp64cmp", "int"):
        if var == "int":
            os.environ["MQ_K1B_INT"] = "1"
        else:
            os.environ.pop("MQ_K1B_INT", None)
        ms = []
        for i in range(4):
            agg = ctx.replay_fifo_jobs_device(c, A.data_ptr(), S.data_ptr(), rows, reps)
            if i:
                ms.append(ctx.last_timing()["sim_ms"])
        t = float(np.median(ms)) * 1e-3
        print(json.dumps({"c": c, "variant": var, "ms": t * 1e3, "jobs_per_s": rows * reps / t,
                          "GBps": 16.0 * rows * reps / t / 1e9, "w1": float(agg[0] / reps)}), flush=True)
os.environ.pop("MQ_K1B_INT", None)
