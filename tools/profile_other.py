"""Small driver for ncu captures of K1 / K3 / K6 / K7 (a few launches each at a reduced size).

    python tools/profile_other.py k1|k3|k3r|k6|k7
"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from most_queue_b200 import _lib
from most_queue_b200.random.distributions import GammaDistribution, H2Distribution, ParetoDistribution

L = _lib
ctx = L.default_context(0)
which = sys.argv[1] if len(sys.argv) > 1 else "k1"
if which == "k1":
    import torch

    R, N = 65536, 1024
    h2 = H2Distribution.get_params_by_mean_and_cv(5.6, 1.5)
    g = torch.Generator(device="cuda").manual_seed(1)
    A = torch.empty((N + 256, R), device="cuda", dtype=torch.float64).exponential_(1.0, generator=g)
    u = torch.rand((N + 256, R), device="cuda", dtype=torch.float64, generator=g)
    S = torch.empty((N + 256, R), device="cuda", dtype=torch.float64).exponential_(1.0, generator=g)
    S = torch.where(u < h2.p1, S / h2.mu1, S / h2.mu2).contiguous()
    torch.cuda.synchronize()
    rec = torch.empty((L.replay_rows(128), R), device="cuda", dtype=torch.float64)
    for _ in range(2):
        ctx.replay_fifo_device(8, None, A.data_ptr(), S.data_ptr(), N + 256, N + 256, R, N, rec.data_ptr(), p_bins=128)
    t = ctx.last_timing()
    print("k1", t, "jobs/s", R * N / (t["sim_ms"] * 1e-3))
elif which in ("k3", "k3r"):
    gam = GammaDistribution.get_params_by_mean_and_cv(1.0, 0.56)
    par = ParetoDistribution.get_params_by_mean_and_cv(0.7, 1.2)
    if which == "k3":
        Nj = 1 << 28
        for _ in range(2):
            ctx.scan_apply(Nj, 0.0, src=L.make_dist("Gamma", gam), srv=L.make_dist("Pa", par), seed=3)
    else:
        import torch

        Nj = 1 << 26
        g = torch.Generator(device="cuda").manual_seed(1)
        A = torch.empty(Nj + 1, device="cuda", dtype=torch.float64).exponential_(1.0, generator=g)
        S = torch.empty(Nj, device="cuda", dtype=torch.float64).exponential_(1.0 / 0.7, generator=g)
        torch.cuda.synchronize()
        for _ in range(2):
            ctx.scan_apply(Nj, 0.0, A=A.data_ptr(), S=S.data_ptr(), device_ptrs=True)
    t = ctx.last_timing()
    print(which, t, "jobs/s", Nj / (t["sim_ms"] * 1e-3))
elif which == "k6":
    caps = [5, 3, 3, 2]
    srv = [L.make_dist("M", 1.2), L.make_dist("M", 1.0), L.make_dist("M", 1.1), L.make_dist("M", 1.3)]
    R, N = 200000, 5000
    for _ in range(2):
        ctx.sim_tandem(caps, L.make_dist("M", 0.9), srv, N, 0.05, R, 0, 5)
    t = ctx.last_timing()
    print("k6", t, "jobs/s", R * N / (t["sim_ms"] * 1e-3))
elif which == "k7":
    R0 = [[1, 0, 0, 0], [0, 0, 0.7, 0.3], [0, 0, 0, 1], [0, 0, 0, 1]]
    R1 = [[0.5, 0.5, 0, 0], [0, 0, 0, 1], [0.2, 0, 0.3, 0.5], [0, 0, 0, 1]]
    srv = [[L.make_dist("H", H2Distribution.get_params_by_mean_and_cv(0.5 + 0.2 * j + 0.1 * i, 1.3)) for j in range(2)]
           for i in range(3)]
    R, N = 100000, 2000
    for _ in range(2):
        ctx.sim_prionet([R0, R1], [2, 1, 2], ["PR", "NP", "RS"], [[0, 1], [1, 0], [0, 1]],
                        [L.make_dist("M", 0.6), L.make_dist("M", 0.7)], srv, N, R, 0, 11)
    t = ctx.last_timing()
    print("k7", t, "jobs/s", R * N / (t["sim_ms"] * 1e-3))
if which in ("k8", "k9"):
    import time as _t
    R, N = 200000, 5000
    if which == "k8":
        d = {"arrivals": L.make_dist("M", 1.0), "services": L.make_dist("M", 1.6), "warm": L.make_dist("M", 2.0),
             "cold": L.make_dist("M", 1.5)}
        for _ in range(2):
            ctx.sim_vacation(1, None, d, N, R, 0, 3, p_bins=32)
    else:
        for _ in range(2):
            ctx.sim_negatives(2, None, L.NEG_TYPES["RCS"], L.make_dist("M", 1.0), L.make_dist("M", 0.3),
                              L.make_dist("M", 0.8), N, R, 0, 3, p_bins=32)
    t = ctx.last_timing()
    print(which, t, "jobs/s", R * N / (t["sim_ms"] * 1e-3))
if which == "k10":
    from most_queue_b200.random.distributions import H2Distribution as _H2
    R, N = 200000, 5000
    h2 = _H2.get_params_by_mean_and_cv(0.75, 1.6)
    for _ in range(2):
        ctx.sim_size(L.SIZE_DISCIPLINES["SRPT"], None, L.make_dist("M", 1.0), L.make_dist("H", h2), N, R, 0, 3, p_bins=32)
    t = ctx.last_timing()
    print(which, t, "jobs/s", R * N / (t["sim_ms"] * 1e-3))
if which == "k8g":
    from most_queue_b200.random.distributions import GammaDistribution as _G
    R, N = 200000, 5000
    g = _G.get_params_by_mean_and_cv(0.625, 0.8)
    d = {"arrivals": L.make_dist("M", 1.0), "services": L.make_dist("Gamma", g), "warm": L.make_dist("Gamma", g),
         "cold": L.make_dist("M", 1.5)}
    for _ in range(2):
        ctx.sim_vacation(1, None, d, N, R, 0, 3, p_bins=32)
    t = ctx.last_timing()
    print(which, t, "jobs/s", R * N / (t["sim_ms"] * 1e-3))
