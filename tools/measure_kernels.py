"""Throughput of every kernel of the hot path on the BASELINE.json configurations (one GPU).

    python tools/measure_kernels.py [--quick] > gpurun_out/kernels.json

Device time comes from the library's CUDA events (mq_last_timing).  Prints one JSON object.
"""

import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

NET_R = [[1, 0, 0, 0, 0, 0], [0, 0.4, 0.6, 0, 0, 0], [0, 0, 0.2, 0.4, 0.4, 0], [0, 0, 0, 0, 1, 0],
         [0, 0, 0, 0, 1, 0], [0, 0, 0, 0, 0, 1]]
NET_CH = [3, 2, 3, 4, 3]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--quick", action="store_true")
    args = ap.parse_args()
    import torch

    from most_queue_b200 import _lib
    from most_queue_b200.random.distributions import GammaDistribution, H2Distribution, ParetoDistribution
    from most_queue_b200.random.utils.params import H2Params

    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        peaks = {"hbm_gbs": 6650.0}
    hbm = float(peaks.get("hbm_gbs", 6650.0))
    ctx = _lib.default_context(0)
    L = _lib
    out = {"gpu": torch.cuda.get_device_name(0), "hbm_peak_gbs": hbm}
    scale = 10 if args.quick else 1

    # ---- cfg1: M/M/3/100 (README quick start), 1e4 replications x 1e5 jobs, K2 finite buffer ----
    R, N = 10000, 100000 // scale
    src, srv = L.make_dist("M", 2.0), L.make_dist("M", 1.0)
    ctx.sim_fifo(3, 100, src, srv, N, R, 0, 1)
    agg, _ = ctx.sim_fifo(3, 100, src, srv, N, R, 0, 2)
    t = ctx.last_timing()
    out["cfg1_mm3_r100"] = {"kernel": "k2_fifo_gen<3,M,M,finite>", "replications": R, "jobs": N,
                            "jobs_per_s": agg[L.A_SERVED] / (t["sim_ms"] * 1e-3), "sim_ms": t["sim_ms"],
                            "w1": agg[L.A_W] / R, "w1_theory": 0.4444444}

    # ---- cfg3: Gamma/Pareto/1, one chain: generator form and replay form (K3) ----
    gam = GammaDistribution.get_params_by_mean_and_cv(1.0, 0.56)
    par = ParetoDistribution.get_params_by_mean_and_cv(0.7, 1.2)
    Nj = 10_000_000_000 // scale
    dsrc, dsrv = L.make_dist("Gamma", gam), L.make_dist("Pa", par)
    ctx.scan_apply(1 << 24, 0.0, src=dsrc, srv=dsrv, seed=3)
    t0 = time.perf_counter()
    stats, _ = ctx.scan_apply(Nj, 0.0, src=dsrc, srv=dsrv, seed=3)
    wall = time.perf_counter() - t0
    t = ctx.last_timing()
    out["cfg3_gamma_pareto_1_generator"] = {
        "kernel": "k3_summary+k3_top+k3_walk (generator)", "jobs": Nj, "jobs_per_s": Nj / (t["sim_ms"] * 1e-3),
        "sim_ms": t["sim_ms"], "wall_s": wall, "w1": stats[L.S_W] / stats[L.S_TAKED],
        "v1": stats[L.S_V] / stats[L.S_SERVED], "utilization_measured": stats[L.S_SUMS] / (stats[L.S_SUMA] + stats[L.S_TAILV])}
    t0 = time.perf_counter()
    stats, _ = ctx.scan_apply(Nj // 10, 0.0, src=dsrc, srv=dsrv, seed=3, want_p=True)
    t = ctx.last_timing()
    T = stats[L.S_SUMA] + stats[L.S_TAILV]
    out["cfg3_gamma_pareto_1_generator_with_p"] = {
        "kernel": "k3 walk with the time-in-state level walk (MQ_SCAN_WANT_P)", "jobs": Nj // 10,
        "jobs_per_s": (Nj // 10) / (t["sim_ms"] * 1e-3), "sim_ms": t["sim_ms"],
        "p0": 1.0 - stats[L.S_LEVELS] / T, "p1": (stats[L.S_LEVELS] - stats[L.S_LEVELS + 1]) / T}
    Nr = (1 << 28) // scale
    g = torch.Generator(device="cuda").manual_seed(1)
    A = torch.empty(Nr + 1, device="cuda", dtype=torch.float64).exponential_(1.0, generator=g)
    S = torch.empty(Nr, device="cuda", dtype=torch.float64).exponential_(1.0 / 0.7, generator=g)
    torch.cuda.synchronize()
    ms = []
    for i in range(4):
        stats, _ = ctx.scan_apply(Nr, 0.0, A=A.data_ptr(), S=S.data_ptr(), device_ptrs=True)
        if i:
            ms.append(ctx.last_timing()["sim_ms"])
    tm = float(np.mean(ms)) * 1e-3
    out["cfg3_replay_2p28"] = {
        "kernel": "k3_summary+k3_top+k3_walk (replay, fp64 A,S resident in HBM)", "jobs": Nr,
        "jobs_per_s": Nr / tm, "sim_ms": tm * 1e3, "w1": stats[L.S_W] / stats[L.S_TAKED],
        "w1_theory_mm1": 0.7 * 0.7 / 0.3,
        "roofline": {"bound": "hbm", "achieved": 16.0 * Nr / tm / 1e9, "peak": hbm, "unit": "GB/s",
                     "frac": 16.0 * Nr / tm / 1e9 / hbm,
                     "note": "algorithmic bytes 16 B/job; this version reads A and S twice (32 B/job of traffic)"}}
    del A, S

    # ---- cfg4: M/Gamma/4, two classes, NP and PR (K4) ----
    R, N = 100000 // scale, 100000 // scale
    g0 = GammaDistribution.get_params_by_mean_and_cv(2.0, 1.2)
    g1 = GammaDistribution.get_params_by_mean_and_cv(4.0, 1.2)
    psrc = [L.make_dist("M", 0.3), L.make_dist("M", 0.5)]
    psrv = [L.make_dist("Gamma", g0), L.make_dist("Gamma", g1)]
    for prty in ("NP", "PR"):
        ctx.sim_prio(4, 2, prty, None, psrc, psrv, 1000, 4096, 0, 1, p_bins=32)
        agg, _ = ctx.sim_prio(4, 2, prty, None, psrc, psrv, N, R, 0, 7, p_bins=32)
        t = ctx.last_timing()
        b0, b1 = 4, 4 + L.prio_agg(32)
        out[f"cfg4_prio_{prty}"] = {"kernel": "k4_prio", "replications": R, "jobs": N,
                                    "jobs_per_s": R * N / (t["sim_ms"] * 1e-3), "sim_ms": t["sim_ms"],
                                    "v1_class0": agg[b0 + 8] / R, "v1_class1": agg[b1 + 8] / R,
                                    "w1_class0": agg[b0] / R, "w1_class1": agg[b1] / R}

    # ---- cfg5: 5-node open network of H2 nodes (K5) ----
    R, N = 100000 // scale, 10000
    nsrv = [L.make_dist("H", H2Distribution.get_params_by_mean_and_cv(0.7 * c, 1.2)) for c in NET_CH]
    ctx.sim_net(NET_R, NET_CH, L.make_dist("M", 1.0), nsrv, 500, 4096, 0, 1)
    agg, _ = ctx.sim_net(NET_R, NET_CH, L.make_dist("M", 1.0), nsrv, N, R, 0, 9)
    t = ctx.last_timing()
    visits = sum(agg[20 + 16 * i + 8] for i in range(5)) / agg[3]
    out["cfg5_network"] = {"kernel": "k5_net", "replications": R, "network_jobs": N,
                           "network_jobs_per_s": agg[3] / (t["sim_ms"] * 1e-3),
                           "node_services_per_s": agg[3] * visits / (t["sim_ms"] * 1e-3), "visits_per_job": visits,
                           "sim_ms": t["sim_ms"], "v1": agg[5] / R}
    print(json.dumps(out, indent=1, default=float))


if __name__ == "__main__":
    main()
