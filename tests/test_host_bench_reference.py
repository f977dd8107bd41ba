"""The reference arm of bench.py (`--impl reference`) runs on host cores only, so it is checked here: it must print one
JSON line in the bench contract, say which CPU implementation it timed (the unmodified reference class from
baseline/_ref when build() has installed it, else the C port) and estimate the same quantity the GPU arm reports."""

import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_prints_a_contract_line():
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1",
                          "--warmup", "0", "--ref-jobs", "3000"], capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    line = json.loads(out.stdout.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["unit"] == "jobs/s" and line["higher_is_better"] is True
    assert line["metric"].startswith("simulated jobs/sec") and line["n_gpus"] == 1 and line["steps"] == 1
    cb = line["cpu_baseline"]
    assert cb["kind"] in ("reference", "port") and cb["cores"] >= 1 and cb["value"] == line["value"] > 0
    if os.path.isdir(os.path.join(ROOT, "baseline", "_ref", "most_queue")):
        assert cb["kind"] == "reference"  # the unmodified most_queue.sim.base.QsSim, not the port
    e2e = line["e2e"]
    assert e2e["value"] == line["value"] and e2e["h2d_bytes_per_step"] == 0 and e2e["d2h_bytes_per_step"] == 0
    # the workload is the GPU arm's: M/H2/8 at rho = 0.7 -- the mean wait of a short run is already in the right region
    assert 0.2 < cb["w1"] < 3.0 and 5.0 < cb["v1"] < 9.0
