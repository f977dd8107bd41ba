"""Small driver for ncu captures of K4 / K5 (one launch each at a reduced size)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from most_queue_b200 import _lib
from most_queue_b200.random.distributions import GammaDistribution, H2Distribution

L = _lib
ctx = L.default_context(0)
g0 = GammaDistribution.get_params_by_mean_and_cv(2.0, 1.2)
g1 = GammaDistribution.get_params_by_mean_and_cv(4.0, 1.2)
psrc = [L.make_dist("M", 0.3), L.make_dist("M", 0.5)]
psrv = [L.make_dist("Gamma", g0), L.make_dist("Gamma", g1)]
which = sys.argv[1] if len(sys.argv) > 1 else "k4"
if which == "k4":
    for _ in range(2):
        ctx.sim_prio(4, 2, "PR", None, psrc, psrv, 5000, 100000, 0, 7, p_bins=32)
    print("k4", ctx.last_timing())
else:
    NET_R = [[1, 0, 0, 0, 0, 0], [0, 0.4, 0.6, 0, 0, 0], [0, 0, 0.2, 0.4, 0.4, 0], [0, 0, 0, 0, 1, 0],
             [0, 0, 0, 0, 1, 0], [0, 0, 0, 0, 0, 1]]
    NET_CH = [3, 2, 3, 4, 3]
    nsrv = [L.make_dist("H", H2Distribution.get_params_by_mean_and_cv(0.7 * c, 1.2)) for c in NET_CH]
    for _ in range(2):
        ctx.sim_net(NET_R, NET_CH, L.make_dist("M", 1.0), nsrv, 1000, 100000, 0, 9)
    print("k5", ctx.last_timing())
