"""Throughput of every kernel of the hot path on the BASELINE.json configurations (one GPU).

    python tests/soak/measure_kernels.py [--quick] > gpurun_out/kernels.json

Device time comes from the library's CUDA events (mq_last_timing).  Prints one JSON object.
"""

import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
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
    from concurrent.futures import ThreadPoolExecutor

    from oracle import oracle

    ncpu = os.cpu_count() or 1

    def cpu_rate(fn, jobs_per_call, calls):
        """jobs/s of the CPU oracle port (oracle/, fp64, the same event loop) on all host threads: `calls`
        independent replications through ctypes (the GIL is released inside the C call)."""
        t0 = time.perf_counter()
        with ThreadPoolExecutor(ncpu) as ex:
            list(ex.map(fn, range(calls)))
        return jobs_per_call * calls / (time.perf_counter() - t0)
    out = {"gpu": torch.cuda.get_device_name(0), "hbm_peak_gbs": hbm}
    scale = 10 if args.quick else 1
    args_quick = args.quick

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

    gsrv = [("Gamma", g0), ("Gamma", g1)]
    out["cfg4_prio_PR"]["cpu_oracle_jobs_per_s"] = cpu_rate(
        lambda r: oracle.prio_generate(4, 2, "PR", [("M", 0.3), ("M", 0.5)], gsrv, 20000, 7, r), 20000, 4 * ncpu)
    out["cfg4_prio_PR"]["cpu_threads"] = ncpu

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
    h2n = [("H", H2Distribution.get_params_by_mean_and_cv(0.7 * c, 1.2)) for c in NET_CH]
    out["cfg5_network"]["cpu_oracle_network_jobs_per_s"] = cpu_rate(
        lambda r: oracle.net_generate(NET_R, NET_CH, ("M", 1.0), h2n, 5000, 9, r), 5000, 4 * ncpu)
    out["cfg5_network"]["cpu_threads"] = ncpu

    # ---- tandem line with blocking (K6) and network of priority nodes (K7) ----
    caps = [5, 3, 3, 2]
    tsrv = [("M", 1.2), ("M", 1.0), ("M", 1.1), ("M", 1.3)]
    R, N = 200000 // scale, 10000
    ctx.sim_tandem(caps, L.make_dist("M", 0.9), [L.make_dist(*x) for x in tsrv], 500, 0.05, 4096, 0, 5)
    agg, _ = ctx.sim_tandem(caps, L.make_dist("M", 0.9), [L.make_dist(*x) for x in tsrv], N, 0.05, R, 0, 5)
    t = ctx.last_timing()
    out["tandem_blocking_4_nodes"] = {
        "kernel": "k6_tandem", "replications": R, "jobs": N, "jobs_per_s": R * N / (t["sim_ms"] * 1e-3),
        "sim_ms": t["sim_ms"], "cpu_threads": ncpu,
        "cpu_oracle_jobs_per_s": cpu_rate(lambda r: oracle.tandem_generate(caps, ("M", 0.9), tsrv, 20000, 5, r),
                                          20000, 4 * ncpu)}
    R0 = [[1, 0, 0, 0], [0, 0, 0.7, 0.3], [0, 0, 0, 1], [0, 0, 0, 1]]
    R1 = [[0.5, 0.5, 0, 0], [0, 0, 0, 1], [0.2, 0, 0.3, 0.5], [0, 0, 0, 1]]
    pn = dict(ch=[2, 1, 2], prty=["PR", "NP", "RS"], npr=[[0, 1], [1, 0], [0, 1]])
    psv = [[("H", H2Distribution.get_params_by_mean_and_cv(0.5 + 0.2 * j + 0.1 * i, 1.3)) for j in range(2)]
           for i in range(3)]
    R, N = 100000 // scale, 5000
    args = ([R0, R1], pn["ch"], pn["prty"], pn["npr"], [L.make_dist("M", 0.6), L.make_dist("M", 0.7)],
            [[L.make_dist(*x) for x in row] for row in psv])
    ctx.sim_prionet(*args, 500, 4096, 0, 11)
    agg, _ = ctx.sim_prionet(*args, N, R, 0, 11)
    t = ctx.last_timing()
    out["priority_network_3_nodes_2_classes"] = {
        "kernel": "k7_prionet", "replications": R, "network_jobs": N,
        "network_jobs_per_s": R * N / (t["sim_ms"] * 1e-3), "sim_ms": t["sim_ms"], "cpu_threads": ncpu,
        "cpu_oracle_network_jobs_per_s": cpu_rate(
            lambda r: oracle.prionet_generate([R0, R1], pn["ch"], pn["prty"], pn["npr"], [("M", 0.6), ("M", 0.7)],
                                              psv, 5000, 11, r), 5000, 4 * ncpu)}

    # ---- cfg2r: replay tier (K1), 65 536 replications x 2 048 jobs of M/H2/8, inputs resident in HBM ----
    R, N = 65536, 2048 // (2 if args_quick else 1)
    h2 = H2Distribution.get_params_by_mean_and_cv(5.6, 1.5)
    g = torch.Generator(device="cuda").manual_seed(1)
    A = torch.empty((N + 256, R), device="cuda", dtype=torch.float64).exponential_(1.0, generator=g)
    u = torch.rand((N + 256, R), device="cuda", dtype=torch.float64, generator=g)
    S = torch.empty((N + 256, R), device="cuda", dtype=torch.float64).exponential_(1.0, generator=g)
    S = torch.where(u < h2.p1, S / h2.mu1, S / h2.mu2).contiguous()
    del u
    rec = torch.empty((L.replay_rows(64), R), device="cuda", dtype=torch.float64)
    torch.cuda.synchronize()
    ms = []
    for i in range(4):
        ctx.replay_fifo_device(8, None, A.data_ptr(), S.data_ptr(), N + 256, N + 256, R, N, rec.data_ptr(), p_bins=64)
        if i:
            ms.append(ctx.last_timing()["sim_ms"])
    tm = float(np.mean(ms)) * 1e-3
    a1, s1 = A[:, 0].cpu().numpy().copy(), S[:, 0].cpu().numpy().copy()
    out["cfg2r_replay"] = {
        "kernel": "k1_fifo_replay<8,infinite>", "replications": R, "jobs": N, "jobs_per_s": R * N / tm,
        "sim_ms": tm * 1e3, "cpu_threads": ncpu,
        "cpu_oracle_jobs_per_s": cpu_rate(lambda r: oracle.fifo_replay(8, None, a1, s1, N), N, 64 * ncpu),
        "roofline": {"bound": "hbm", "achieved": 16.0 * R * N / tm / 1e9, "peak": hbm, "unit": "GB/s",
                     "frac": 16.0 * R * N / tm / 1e9 / hbm,
                     "note": "algorithmic bytes 16 B/job; the kernel is latency / issue bound (fp64, bit-exact)"}}
    print(json.dumps(out, indent=1, default=float))


if __name__ == "__main__":
    main()
