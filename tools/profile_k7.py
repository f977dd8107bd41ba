"""Times K7 v2 (network of priority nodes, 3 nodes x 2 classes, the shape of tests/soak/measure_kernels.py)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from most_queue_b200 import _lib as L
from most_queue_b200.random.distributions import H2Distribution

ctx = L.default_context()
R0 = [[1, 0, 0, 0], [0, 0, 0.7, 0.3], [0, 0, 0, 1], [0, 0, 0, 1]]
R1 = [[0.5, 0.5, 0, 0], [0, 0, 0, 1], [0.2, 0, 0.3, 0.5], [0, 0, 0, 1]]
pn = dict(ch=[2, 1, 2], prty=["PR", "NP", "RS"], npr=[[0, 1], [1, 0], [0, 1]])
psv = [[("H", H2Distribution.get_params_by_mean_and_cv(0.5 + 0.2 * j + 0.1 * i, 1.3)) for j in range(2)] for i in range(3)]
R, N = 100000, 5000
args = ([R0, R1], pn["ch"], pn["prty"], pn["npr"], [L.make_dist("M", 0.6), L.make_dist("M", 0.7)],
        [[L.make_dist(*x) for x in row] for row in psv])
ctx.sim_prionet(*args, 500, 4096, 0, 11)
for _ in range(2):
    agg, _ = ctx.sim_prionet(*args, N, R, 0, 11)
t = ctx.last_timing()
print("k7", t["sim_ms"], "ms", R * N / (t["sim_ms"] * 1e-3), "network jobs/s")
