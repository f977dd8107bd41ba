import os, sys
import numpy as np
sys.path.insert(0, "/root/repo")
from most_queue_b200.sim.lindley import replay_chain
N = 40_000_000
rng = np.random.default_rng(3)
A = rng.exponential(1.0, N + 1); S = rng.exponential(0.9, N)
os.environ["MQ_SCAN_TWO_PASS"] = "1"
o2, W2 = replay_chain(A, S, N)
os.environ.pop("MQ_SCAN_TWO_PASS")
for rep in range(80):
    o1, W1 = replay_chain(A, S, N)
    bad = np.flatnonzero(W1 != W2)
    if bad.size == 0:
        print(rep, "ok"); continue
    tiles = np.unique(bad // 4096)
    first = bad[0]
    runs = np.split(bad, np.flatnonzero(np.diff(bad) > 1) + 1)
    print(rep, "mismatches", bad.size, "tiles", tiles[:8], "n_tiles", tiles.size, "first idx", first, "tile", first // 4096, "pos in tile", first % 4096,
          "seg", (first % 4096) // 16, "pos in seg", first % 16, "run lengths", [len(r) for r in runs[:6]],
          "cta", (first // 4096) % 148, "iter", (first // 4096) // 148, "chunkpos", (first // 4096) % 10,
          "W1", W1[first], "W2", W2[first], "pairs equal", o1["pairs_apply"] == o2["pairs_apply"], "sums", o1["sum_a"] == o2["sum_a"], o1["sum_s"] == o2["sum_s"])
