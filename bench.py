"""Benchmark of the hot path: simulated jobs/sec of M/H2/8 FIFO (BASELINE.json configs[1]).

    python bench.py --gpus N --steps K --warmup W            # our arm (CUDA kernels via the C ABI)
    python bench.py --impl reference --gpus N --steps K ...   # CPU arm: the reference algorithm on
                                                              # the host cores (oracle port, all threads)

One "step" = one pass of QsSim.run(): `replications` independent M/H2/8 replications of
`total_served` jobs each, per GPU (weak scaling: every rank simulates the full configuration with
its own global replication ids).  Rank 0 prints ONE JSON line.

    --config cfg2|cfg3|cfg5   cfg2 (default, the headline): M/H2/8; cfg3: one Gamma/Pareto/1 chain of 1e10 jobs cut into
                              one contiguous range per GPU (summary -> all-gather of (a, b) -> carried walk); cfg5: the
                              5-node H2 network, replications sharded
    --scaling weak|strong     weak (default): every GPU gets the full configuration; strong: the configuration's total
                              (cfg2: 1e6 replications, cfg5: 1e5 replications, cfg3: always the one 1e10-job chain) is
                              split over the GPUs
The reference arm (--impl reference) times the UNMODIFIED reference class (baseline/_ref) on all host cores.
"""

from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

# cfg2 of BASELINE.json: M/H2/8, rho = 0.7, CV = 1.5 (H2 fitted to mean 5.6, cv 1.5)
H2 = {"p1": 0.447586, "mu1": 0.0950716, "mu2": 0.619214}
C_SERVERS = 8
ARRIVAL_RATE = 1.0
I_ALG = 160.0  # planning lane-instructions per job for cfg2: the issue roofline's work unit
METRIC = "simulated jobs/sec (M/H2/8 FIFO)"
# planning budgets of SURVEY 8(d), lane-instructions per job, and what ncu measured (profiles/, issue-lane-slots per job
# = warp instructions x 32 / jobs)
I_ALG_OTHER = {"cfg3": 120.0, "cfg4": 220.0, "cfg5": 520.0}
I_MEASURED = {"cfg2": 227.0, "cfg4": 53.0 * 32.0, "cfg5": None, "cfg3": None}
NET_R = [[1, 0, 0, 0, 0, 0], [0, 0.4, 0.6, 0, 0, 0], [0, 0, 0.2, 0.4, 0.4, 0], [0, 0, 0, 0, 1, 0], [0, 0, 0, 0, 1, 0],
         [0, 0, 0, 0, 0, 1]]
NET_CH = [3, 2, 3, 4, 3]


def issue_roofline(jobs_per_s_per_gpu, i_alg, kernel, peaks, peak_src, sms=148, extra=None):
    """Issue roofline of a generate-and-simulate kernel: achieved = jobs/s/GPU x planning lane-instructions per job."""
    sm_max = float(peaks.get("sm_max_mhz", 1965.0))
    peak = sms * 128 * sm_max * 1e6 / 1e9
    achieved = jobs_per_s_per_gpu * i_alg / 1e9
    r = {"bound": "issue", "achieved": achieved, "peak": peak, "unit": "Glane-instr/s", "frac": achieved / peak,
         "traffic": None, "kernel": kernel,
         "note": f"achieved = jobs/s/GPU x {i_alg:.0f} planning lane-instructions per job (SURVEY 8d); peak = {sms} SMs x "
                 f"128 lanes x {sm_max:.0f} MHz ({peak_src} clock)"}
    if extra:
        r.update(extra)
    return r


def _peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return json.load(f), "measured"
    except Exception:
        return {"hbm_gbs": 6650.0, "sm_max_mhz": 1965.0}, "fallback"


class ClockSampler:
    """Samples nvidia-smi clocks / throttle reasons while the timed region runs."""

    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index, self.rows, self._stop, self._t = index, [], threading.Event(), None

    def _run(self):
        while not self._stop.is_set():
            try:
                out = subprocess.run(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                      "-i", str(self.index)], capture_output=True, text=True, timeout=5).stdout
                self.rows.append([x.strip() for x in out.strip().split(",")])
            except Exception:
                pass
            self._stop.wait(0.2)

    def __enter__(self):
        self._t = threading.Thread(target=self._run, daemon=True)
        self._t.start()
        return self

    def __exit__(self, *a):
        self._stop.set()
        self._t.join(timeout=6)

    def summary(self):
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[0]))
                mx.append(float(r[1]))
                for n, v in zip(names, r[4:8]):
                    if v.lower().startswith("active"):
                        reasons.add(n)
            except Exception:
                continue
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        return {"sm_mhz": float(np.median(sm)), "sm_max_mhz": float(max(mx)), "reasons": sorted(reasons),
                "samples": len(sm)}


def cpu_port(seconds_target: float, jobs: int, threads: int | None = None):
    """Times the oracle port (oracle/mq_oracle.c: the builder's C restatement of QsSim.run, fp64, libm) on a bounded
    sample of the same workload: `reps` replications of `jobs` jobs spread over all host threads.  A second, clearly
    labelled CPU figure (`cpu_port`): ~190x faster per core than the reference class, so the conservative one."""
    from oracle import oracle

    threads = threads or os.cpu_count() or 1

    class Prm:
        p1, mu1, mu2 = H2["p1"], H2["mu1"], H2["mu2"]

    src, srv = ("M", ARRIVAL_RATE), ("H", Prm)
    # calibrate with one replication per thread, then size the sample for ~seconds_target
    t0 = time.perf_counter()
    oracle.fifo_generate(C_SERVERS, None, src, srv, jobs, 12345, 0, threads, p_bins=32, threads=threads)
    dt = max(time.perf_counter() - t0, 1e-3)
    rounds = max(1, int(seconds_target / dt))
    reps = threads * rounds
    t0 = time.perf_counter()
    res = oracle.fifo_generate(C_SERVERS, None, src, srv, jobs, 12345, threads, reps, p_bins=32, threads=threads)
    dt = time.perf_counter() - t0
    served = sum(r.served for r in res)
    return {"value": served / dt, "unit": "jobs/s", "cores": threads, "kind": "port",
            "sample": f"{reps} replications x {jobs} jobs of the same M/H2/8 workload, {dt:.1f} s, "
                      f"oracle/mq_oracle.c (fp64, libm) on {threads} threads",
            "seconds": dt, "w1": float(np.mean([r.w_sum[0] / r.taked for r in res]))}


def _reference_worker(task):
    """One process = one unmodified reference `QsSim(8, seed=s).run(jobs)` (most_queue/sim/base.py:29,488) imported
    from baseline/_ref; stdout / stderr (tqdm, colour banners) go to /dev/null."""
    seed, jobs = task
    import contextlib
    import io

    sys.path.insert(0, os.path.join(ROOT, "baseline", "_ref"))
    from most_queue.random.utils.params import H2Params
    from most_queue.sim.base import QsSim

    with contextlib.redirect_stdout(io.StringIO()), contextlib.redirect_stderr(io.StringIO()):
        qs = QsSim(C_SERVERS, verbose=False, seed=seed)
        qs.set_sources(ARRIVAL_RATE, "M")
        qs.set_servers(H2Params(**H2), "H")
        t0 = time.perf_counter()
        r = qs.run(jobs)
        dt = time.perf_counter() - t0
    return qs.served, dt, float(r.w[0]), float(r.v[0])


class ReferencePool:
    """All host cores running the unmodified reference class, one process per core, distinct seeds."""

    def __init__(self, cores: int | None = None):
        import multiprocessing as mp

        self.cores = cores or os.cpu_count() or 1
        # bench.py's reference arm never initialises CUDA, so fork is safe (and the GPU arm calls it in a subprocess)
        self.pool = mp.get_context("fork").Pool(self.cores)
        self.seed = 1000

    def step(self, jobs: int, rounds: int = 1):
        tasks = [(self.seed + i, jobs) for i in range(self.cores * rounds)]
        self.seed += len(tasks)
        t0 = time.perf_counter()
        rows = self.pool.map(_reference_worker, tasks, chunksize=1)
        dt = time.perf_counter() - t0
        return {"served": sum(r[0] for r in rows), "seconds": dt,
                "per_core": float(np.mean([r[0] / r[1] for r in rows])),
                "w1": float(np.mean([r[2] for r in rows])), "v1": float(np.mean([r[3] for r in rows])),
                "replications": len(rows)}

    def close(self):
        self.pool.close()
        self.pool.join()


def reference_available() -> bool:
    return os.path.isdir(os.path.join(ROOT, "baseline", "_ref", "most_queue"))


def run_reference(args):
    """`--impl reference`: the reference's own CPU implementation of the path -- the unmodified pure-Python `QsSim`
    (baseline/_ref, installed by baseline/install_reference.py) -- on all host cores, one process per core, each step a
    bounded sample of the cfg2 workload (one replication of `--ref-jobs` jobs per core).  If baseline/_ref is missing the
    C port is timed instead and the line says so (`kind: "port"`)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return  # the CPU arm runs on rank 0 only
    jobs = args.ref_jobs
    if reference_available() and not args.ref_port:
        pool = ReferencePool()
        rows = []
        for i in range(args.warmup + args.steps):
            r = pool.step(jobs)
            if i >= args.warmup:
                rows.append(r)
        pool.close()
        total_jobs = sum(r["served"] for r in rows)
        total_s = sum(r["seconds"] for r in rows)
        value = total_jobs / total_s
        base = {"value": value, "unit": "jobs/s", "cores": pool.cores, "kind": "reference",
                "per_core": float(np.mean([r["per_core"] for r in rows])),
                "sample": f"per step {pool.cores} replications x {jobs} jobs of the same M/H2/8 workload: the unmodified "
                          f"most_queue.sim.base.QsSim (baseline/_ref), one process per host core, distinct seeds",
                "w1": float(np.mean([r["w1"] for r in rows])), "v1": float(np.mean([r["v1"] for r in rows]))}
        dtype, gen = "f64", "numpy PCG64 (the reference's own generator)"
    else:
        vals = []
        for i in range(args.warmup + args.steps):
            b = cpu_port(args.ref_seconds, args.jobs)
            if i >= args.warmup:
                vals.append(b)
        total_jobs = sum(b["value"] * b["seconds"] for b in vals)
        total_s = sum(b["seconds"] for b in vals)
        value = total_jobs / total_s
        base = dict(vals[-1], value=value)
        base.pop("seconds", None)
        rows = vals
        dtype, gen = "f64", "Philox2x32-10"
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": "jobs/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * total_s / max(len(rows), 1),
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": dtype, "data": "synthetic",
        "config": {"workload": f"M/H2/{C_SERVERS} FIFO rho=0.7 CV=1.5 (BASELINE cfg2), bounded sample per step: "
                               f"{base['cores']} replications x {jobs} jobs", "generator": gen},
        "cpu_baseline": base,
        "e2e": {"value": value, "unit": "jobs/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def reference_subprocess(jobs: int, steps: int, warmup: int, port: bool = False, timeout: float = 600.0):
    """Runs the reference arm in a fresh interpreter (no CUDA context to fork) and returns its `cpu_baseline`."""
    cmd = [sys.executable, os.path.abspath(__file__), "--impl", "reference", "--steps", str(steps), "--warmup",
           str(warmup), "--ref-jobs", str(jobs)] + (["--ref-port"] if port else [])
    env = {k: v for k, v in os.environ.items() if k not in ("RANK", "LOCAL_RANK", "WORLD_SIZE")}
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=timeout, env=env)
    for ln in reversed(out.stdout.strip().splitlines()):
        if ln.startswith("{"):
            return json.loads(ln)["cpu_baseline"]
    raise RuntimeError(f"reference arm failed: {out.stderr[-400:]}")


def run_ours(args):
    import torch
    import torch.distributed as tdist

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    torch.cuda.set_device(local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        tdist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    from most_queue_b200 import _lib
    from most_queue_b200.random.utils.params import H2Params
    from most_queue_b200.sim.base import QsSim

    peaks, peak_src = _peaks()
    reps, jobs = args.reps, args.jobs
    p_bins = 128
    h2 = H2Params(**H2)
    ctx = _lib.default_context(local_rank)

    def barrier():
        if world > 1:
            tdist.barrier()
        torch.cuda.synchronize()

    def maxreduce(x: float) -> float:
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        tdist.all_reduce(t, op=tdist.ReduceOp.MAX)
        return float(t.item())

    src, srv = _lib.make_dist("M", ARRIVAL_RATE), _lib.make_dist("H", h2)

    # ---- device-resident throughput: the C-ABI call, timed with CUDA events on its stream ----
    strong = args.scaling == "strong"
    if strong:  # the configuration's 1e6 replications in total, this rank's contiguous share
        from most_queue_b200 import dist as mqdist
        rep0_rank, reps_rank = mqdist.shard(reps, rank, world)
    else:       # every rank simulates `reps` replications with its own global ids
        rep0_rank, reps_rank = rank * reps, reps

    def step(i):
        agg, _ = ctx.sim_fifo(C_SERVERS, None, src, srv, jobs, reps_rank, rep0_rank, 20260101 + i, p_bins=p_bins)
        return agg, ctx.last_timing()

    for i in range(args.warmup):
        step(i)
    barrier()
    dev_ms, sim_ms, launches = 0.0, 0.0, 0
    served = 0.0
    agg_sum = None
    with ClockSampler(local_rank) as clk:
        t0 = time.perf_counter()
        for i in range(args.steps):
            agg, tm = step(args.warmup + i)
            dev_ms += tm["sim_ms"] + tm["reduce_ms"]
            sim_ms += tm["sim_ms"]
            launches += tm["launches"]
            served += agg[_lib.A_SERVED]
            agg_sum = agg if agg_sum is None else agg_sum + agg
        barrier()
        wall_ms = 1e3 * (time.perf_counter() - t0)
    clocks = clk.summary()
    dev_ms = maxreduce(dev_ms)
    wall_ms = maxreduce(wall_ms)
    sim_ms_max = maxreduce(sim_ms)
    jobs_all = float(served)
    if world > 1:
        t = torch.tensor([jobs_all], dtype=torch.float64, device="cuda")
        tdist.all_reduce(t)
        jobs_all = float(t.item())
    value = jobs_all / (dev_ms * 1e-3)

    # ---- end to end through the public API (QsSim.run with replications=), host in / host out ----
    def api_step(i):
        qs = QsSim(C_SERVERS, seed=777 + i, p_bins=p_bins)
        qs.set_sources(ARRIVAL_RATE, "M")
        qs.set_servers(h2, "H")
        # under torchrun QsSim shards `replications` over the ranks and allreduces the aggregate
        res = qs.run(jobs, replications=reps if strong else reps * world)
        return qs, res

    api_step(0)
    barrier()
    t0 = time.perf_counter()
    e2e_served = 0
    for i in range(args.steps):
        qs, res = api_step(1 + i)
        e2e_served += qs.total_served_all
    barrier()
    e2e_s = maxreduce(time.perf_counter() - t0)
    e2e_value = e2e_served / e2e_s
    h2d = int(__import__("ctypes").sizeof(_lib.MqFifoCfg))
    d2h = 8 * _lib.agg_len(p_bins)

    # ---- moments against theory (Takahashi-Takami, tests/golden/theory.json) ----
    moment_err = None
    try:
        with open(os.path.join(ROOT, "tests", "golden", "theory.json")) as f:
            th = json.load(f)["mh2_8"]
        R = agg_sum[_lib.A_REPS]
        w1 = agg_sum[_lib.A_W] / R
        v1 = agg_sum[_lib.A_V] / R
        sd_w = np.sqrt(max(agg_sum[_lib.A_W2] / R - w1 * w1, 0.0) / R)
        sd_v = np.sqrt(max(agg_sum[_lib.A_V2] / R - v1 * v1, 0.0) / R)
        moment_err = {"w1": w1, "w1_theory": th["w"][0], "w1_ci99": 2.5758 * sd_w,
                      "v1": v1, "v1_theory": th["v"][0], "v1_ci99": 2.5758 * sd_v,
                      "note": "QsSim starts empty and keeps all jobs: O(1/N) start-up bias is part of the estimator"}
    except Exception:
        pass

    # the same workload with the additive warm-up (first 5 % of the jobs discarded, v not censored at the
    # stop): the estimator is then unbiased and must sit inside the plain 99 % CI of the theory
    if moment_err is not None and world == 1:
        try:
            aggw, _ = ctx.sim_fifo(C_SERVERS, None, src, srv, jobs, reps, 0, 424242, p_bins=p_bins,
                                   warmup=max(1, jobs // 20), v_at_start=True)
            Rw = aggw[_lib.A_REPS]
            w1, v1 = aggw[_lib.A_W] / Rw, aggw[_lib.A_V] / Rw
            ciw = 2.5758 * np.sqrt(max(aggw[_lib.A_W2] / Rw - w1 * w1, 0.0) / Rw)
            civ = 2.5758 * np.sqrt(max(aggw[_lib.A_V2] / Rw - v1 * v1, 0.0) / Rw)
            moment_err["with_warmup"] = {"warmup_jobs": max(1, jobs // 20), "w1": w1, "w1_ci99": ciw,
                                         "w1_minus_theory": w1 - moment_err["w1_theory"], "v1": v1, "v1_ci99": civ,
                                         "v1_minus_theory": v1 - moment_err["v1_theory"]}
        except Exception as e:
            moment_err["with_warmup"] = {"error": repr(e)}

    if rank == 0:
        sm_max = float(peaks.get("sm_max_mhz", 1965.0))
        sms = torch.cuda.get_device_properties(local_rank).multi_processor_count
        peak_issue = sms * 128 * sm_max * 1e6 / 1e9  # G lane-instructions / s
        per_gpu_jobs_s = (jobs_all / world) / (sim_ms_max * 1e-3)
        achieved = per_gpu_jobs_s * I_ALG / 1e9
        line = {
            "metric": METRIC, "value": value, "unit": "jobs/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": wall_ms / args.steps, "higher_is_better": True,
            "scaling": args.scaling, "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": f"M/H2/{C_SERVERS} FIFO rho=0.7 CV=1.5 (BASELINE cfg2): {reps} replications x "
                                   f"{jobs} jobs " + ("in total (strong scaling)" if strong else "per GPU"),
                       "replications_per_gpu": reps_rank, "jobs_per_replication": jobs,
                       "p_bins": p_bins, "generator": "Philox2x32-10, counter=(arrival index, global replication id)",
                       "cache": "generator mode has no input arrays; per-replication records "
                                f"({8 * _lib.rec_rows(p_bins) * reps / 1e6:.0f} MB) exceed L2",
                       "parallelism": f"replications sharded over {world} GPU(s), one allreduce of {d2h} B"},
            "clocks": clocks,
            "e2e": {"value": e2e_value, "unit": "jobs/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "api": "QsSim(8).set_sources/set_servers/run(total_served, replications=)"},
            "gpu_launches": launches,
            "roofline": {"bound": "issue", "achieved": achieved, "peak": peak_issue, "unit": "Glane-instr/s",
                         "frac": achieved / peak_issue,
                         # dram__bytes_read.sum + dram__bytes_write.sum of one full-size launch
                         # (profiles/r01_ncu_k2_full_traffic.csv, the final kernel: 0.82 GB read + 3.31 GB written); valid for
                         # the default 1e6 x 1e5 configuration
                         "traffic": 4.131e9 if (reps, jobs) == (1_000_000, 100_000) else None,
                         "kernel": "k2_fifo_gen<8,M,H2,infinite>",
                         "note": f"achieved = jobs/s/GPU x {I_ALG:.0f} algorithmic lane-instructions per job "
                                 f"(SURVEY 8d); peak = {sms} SMs x 128 lanes x {sm_max:.0f} MHz ({peak_src} clock); "
                                 "traffic in bytes per launch (fp64 per-replication records, ~4 KB per replication incl. "
                                 "flushes): HBM is idle, the kernel is issue bound"},
            "moments": moment_err,
        }
        if world == 1 and not args.no_cpu:
            # the reference's own class on all host cores (bounded sample), and the C port as a second figure
            try:
                if reference_available():
                    line["cpu_baseline"] = reference_subprocess(args.ref_jobs, steps=3, warmup=1)
            except Exception as e:
                line["cpu_baseline_error"] = repr(e)[:300]
            port = cpu_port(args.cpu_seconds, jobs)
            port.pop("seconds", None)
            line["cpu_port"] = port
            if "cpu_baseline" not in line:
                line["cpu_baseline"] = port
        if world == 1 and not args.no_replay:
            try:
                line["replay"] = replay_leg(ctx, peaks, peak_src)
            except Exception as e:  # the headline must not be lost to the secondary measurement
                line["replay"] = {"error": repr(e)}
            try:
                line["replay_jobs"] = replay_jobs_leg(ctx, peaks, peak_src)
            except Exception as e:
                line["replay_jobs"] = {"error": repr(e)[:200]}
        if world == 1 and not args.no_others:
            try:
                line["other_configs"] = other_configs_leg(peaks, peak_src, sms)
            except Exception as e:
                line["other_configs"] = {"error": repr(e)}
        if world == 1 and not args.no_fp64:
            try:
                line["k2_fp64"] = fp64_leg(ctx, src, srv, jobs, reps, p_bins, agg, args.warmup + args.steps - 1)
            except Exception as e:
                line["k2_fp64"] = {"error": repr(e)[:200]}
        print(json.dumps(line), flush=True)
    if world > 1:
        tdist.destroy_process_group()


def fp64_leg(ctx, src, srv, jobs, reps, p_bins, agg32, step_index):
    """The same loop in fp64 (MQ_FIFO_FP64: fp64 clocks, CUDA double-precision log, fp64 power sums) on the SAME seed and
    replication ids as the last timed fp32 step: throughput of the fp64 tier and the paired difference of every raw
    moment -- the measured cost of the fp32 production arithmetic ('nothing measures it directly')."""
    from most_queue_b200 import _lib

    agg64, _ = ctx.sim_fifo(C_SERVERS, None, src, srv, jobs, reps, 0, 20260101 + step_index, p_bins=p_bins, fp64=True)
    tm = ctx.last_timing()
    R = agg64[_lib.A_REPS]
    out = {"jobs_per_s": float(agg64[_lib.A_SERVED]) / (tm["sim_ms"] * 1e-3), "ms": tm["sim_ms"], "replications": int(R),
           "kernel": "k2_fifo_gen<double,8,runtime kinds,infinite>", "same_seed_as_step": step_index}
    for name, a, a2 in (("w", _lib.A_W, _lib.A_W2), ("v", _lib.A_V, _lib.A_V2)):
        m64 = agg64[a:a + 4] / R
        m32 = agg32[a:a + 4] / agg32[_lib.A_REPS]
        ci = 2.5758 * np.sqrt(np.maximum(agg64[a2:a2 + 4] / R - m64 * m64, 0.0) / R)
        out[name + "_f64"] = [float(x) for x in m64]
        out[name + "_f32_minus_f64"] = [float(x) for x in (m32 - m64)]
        out[name + "_rel_diff"] = [float(x) for x in (m32 - m64) / m64]
        out[name + "_ci99_of_the_run"] = [float(x) for x in ci]
    p64 = agg64[_lib.A_P:_lib.A_P + 16] / R
    p32 = agg32[_lib.A_P:_lib.A_P + 16] / agg32[_lib.A_REPS]
    out["p_max_abs_diff"] = float(np.max(np.abs(p32 - p64)))
    return out


def other_configs_leg(peaks, peak_src, sms):
    """The other BASELINE.json configurations through the public API (parity-test cases, reported for reference:
    device time of the simulation kernels from the library's CUDA events, one run each after a small warm-up), each with
    the roofline object SURVEY 8(d) defines for it."""
    from most_queue_b200.random.distributions import GammaDistribution, H2Distribution, ParetoDistribution
    from most_queue_b200.sim.base import QsSim
    from most_queue_b200.sim.lindley import run_chain
    from most_queue_b200.sim.networks.network import NetworkSimulator
    from most_queue_b200.sim.priority import PriorityQueueSimulator

    out = {}
    # cfg1: M/M/3/100 (README quick start), 1e4 replications x 1e5 jobs
    qs = QsSim(3, buffer=100, seed=1, replications=10_000)
    qs.set_sources(2.0, "M")
    qs.set_servers(1.0, "M")
    qs.run(1000)
    r = qs.run(100_000)
    out["cfg1_mm3_r100"] = {"jobs_per_s": 1e4 * 1e5 / (qs.timing["sim_ms"] * 1e-3), "w1": r.w[0], "w1_theory": 4.0 / 9.0}
    # cfg3: Gamma/Pareto/1, one chain of 1e10 jobs (max-plus Lindley scan)
    gam = GammaDistribution.get_params_by_mean_and_cv(1.0, 0.56)
    par = ParetoDistribution.get_params_by_mean_and_cv(0.7, 1.2)
    run_chain(("Gamma", gam), ("Pa", par), 1 << 24, seed=3, want_p=False)
    c3 = run_chain(("Gamma", gam), ("Pa", par), 10_000_000_000, seed=3, want_p=False)
    j3 = 1e10 / (c3.timing["sim_ms"] * 1e-3)
    out["cfg3_gamma_pareto_1_chain"] = {"jobs_per_s": j3, "w1": c3.w[0], "utilization": c3.utilization,
                                        "roofline": issue_roofline(j3, I_ALG_OTHER["cfg3"], "k3_summary + k3_walk (generator form, "
                                                                   "fp32 draw stash)", peaks, peak_src, sms)}
    # cfg3, replay form (2^28 jobs = 4 GiB of fp64 A, S resident in HBM, 16 B/job): the HBM-bound kernel of the set
    try:
        import torch

        from most_queue_b200 import _lib

        ctx = _lib.default_context()
        nr = 1 << 28
        g = torch.Generator(device="cuda").manual_seed(1)
        A = torch.empty(nr + 1, device="cuda", dtype=torch.float64).exponential_(1.0, generator=g)
        S = torch.empty(nr, device="cuda", dtype=torch.float64).exponential_(1.0 / 0.7, generator=g)
        torch.cuda.synchronize()
        ms = []
        for i in range(4):
            st, _ = ctx.scan_apply(nr, 0.0, A=A.data_ptr(), S=S.data_ptr(), device_ptrs=True)
            if i:
                ms.append(ctx.last_timing()["sim_ms"])
        del A, S
        t = float(np.mean(ms)) * 1e-3
        peaks, peak_src = _peaks()
        peak = float(peaks.get("hbm_gbs", 6650.0))
        out["cfg3_replay_2p28"] = {
            "jobs_per_s": nr / t, "ms": t * 1e3, "w1": float(st[_lib.S_W] / st[_lib.S_TAKED]), "w1_theory_mm1": 0.7 * 0.7 / 0.3,
            "roofline": {"bound": "hbm", "achieved": 16.0 * nr / t / 1e9, "peak": peak, "unit": "GB/s",
                         "frac": 16.0 * nr / t / 1e9 / peak, "traffic": 16.11 * nr,
                         "kernel": "k3_onepass (TMA stages, speculative walk in shared memory, decoupled look-back + repair: one HBM read)",
                         "note": "algorithmic bytes 16 B/job (A and S read once); traffic = dram__bytes_read + write of k3_onepass from "
                                 "profiles/r02_ncu_k3_onepass_summary.csv (2^26 jobs: 16.04 B/job read + 0.07 written), scaled; the three-launch "
                                 "path (MQ_SCAN_TWO_PASS=1) moves 33.75 B/job; peak " + peak_src}}
    except Exception as e:  # reported, never fatal for the headline line
        out["cfg3_replay_2p28"] = {"error": repr(e)[:200]}
    # cfg4: M/Gamma/4, two classes, NP and PR, 1e5 replications x 1e5 jobs
    g0 = GammaDistribution.get_params_by_mean_and_cv(2.0, 1.2)
    g1 = GammaDistribution.get_params_by_mean_and_cv(4.0, 1.2)
    for prty in ("NP", "PR"):
        ps = PriorityQueueSimulator(4, 2, prty, seed=7, replications=100_000, p_bins=32)
        ps.set_sources([{"type": "M", "params": 0.3}, {"type": "M", "params": 0.5}])
        ps.set_servers([{"type": "Gamma", "params": g0}, {"type": "Gamma", "params": g1}])
        ps.run(1000, replications=4096)
        r = ps.run(100_000)
        j4 = 1e5 * 1e5 / (ps.timing["sim_ms"] * 1e-3)
        out[f"cfg4_mg4_2class_{prty}"] = {"jobs_per_s": j4, "v1": [r.v[0][0], r.v[1][0]],
                                         "roofline": issue_roofline(j4, I_ALG_OTHER["cfg4"], "k4v2_prio<4 slots,2 classes,generator>",
                                                                    peaks, peak_src, sms)}
    # cfg5: 5-node open network of H2 nodes, 1e5 replications x 1e4 network jobs
    R = [[1, 0, 0, 0, 0, 0], [0, 0.4, 0.6, 0, 0, 0], [0, 0, 0.2, 0.4, 0.4, 0], [0, 0, 0, 0, 1, 0], [0, 0, 0, 0, 1, 0],
         [0, 0, 0, 0, 0, 1]]
    ch = [3, 2, 3, 4, 3]
    net = NetworkSimulator(seed=9, replications=100_000)
    net.set_sources(1.0, np.array(R, dtype=float))
    net.set_nodes([{"type": "H", "params": H2Distribution.get_params_by_mean_and_cv(0.7 * c, 1.2)} for c in ch], ch)
    net.run(500, replications=4096)
    r = net.run(10_000)
    j5 = 1e5 * 1e4 / (net.timing["sim_ms"] * 1e-3)
    out["cfg5_network_5_nodes"] = {"network_jobs_per_s": j5, "v1": r.v[0],
                                   "roofline": issue_roofline(j5, I_ALG_OTHER["cfg5"], "k5v2_net<16 server slots,generator>",
                                                              peaks, peak_src, sms)}
    return out


def replay_leg(ctx, peaks, peak_src, reps=65536, rows=4096, steps=3):
    """cfg 2r: K1 on HBM-resident fp64 arrays (16 B/job), timed with CUDA events; HBM roofline."""
    import torch

    from most_queue_b200 import _lib

    g = torch.Generator(device="cuda").manual_seed(1)
    A = -torch.log(1.0 - torch.rand((rows, reps), generator=g, device="cuda", dtype=torch.float64))
    br = torch.rand((rows, reps), generator=g, device="cuda", dtype=torch.float64) < H2["p1"]
    S = -torch.log(1.0 - torch.rand((rows, reps), generator=g, device="cuda", dtype=torch.float64))
    S = S / torch.where(br, H2["mu1"], H2["mu2"])
    del br
    n = int(rows * 0.9)  # departures per replication; leaves slack in the arrays
    rec = torch.empty((_lib.replay_rows(64), reps), device="cuda", dtype=torch.float64)
    torch.cuda.synchronize()
    ms = []
    for i in range(steps + 1):
        agg = ctx.replay_fifo_device(C_SERVERS, None, A.data_ptr(), S.data_ptr(), rows, rows, reps, n,
                                     rec.data_ptr(), p_bins=64)
        if i:
            ms.append(ctx.last_timing()["sim_ms"])
    jobs = float(agg[_lib.A_SERVED])
    used = float((rec[_lib.rec_rows(64) + _lib.RF_AUSED].sum() + rec[_lib.rec_rows(64) + _lib.RF_SUSED].sum()).item())
    t = float(np.mean(ms)) * 1e-3
    achieved = 8.0 * used / t / 1e9
    return {"kernel": "k1_fifo_replay<8,infinite>", "jobs_per_s": jobs / t, "ms": t * 1e3,
            "workload": f"{reps} replications x {n} jobs, fp64 A,S resident in HBM ([draw][replication])",
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": float(peaks.get("hbm_gbs", 6650.0)),
                         "unit": "GB/s", "frac": achieved / float(peaks.get("hbm_gbs", 6650.0)), "traffic": None,
                         "note": f"algorithmic bytes = 8 B x variates consumed (16 B/job); peak {peak_src}"}}


def run_other_config(args):
    """--config cfg3 | cfg5 under the same launch contract (one rank per GPU, barrier + synchronize around exactly K timed
    steps, max over ranks, rank 0 prints one JSON line).  cfg3: ONE Gamma/Pareto/1 chain of --chain-jobs jobs (1e10) cut
    into one contiguous range per rank: local max-plus summary -> all-gather of the (a, b) pairs -> carries folded in rank
    order -> walk from the carry -> allreduce of the sums (always strong scaling: it is one chain).  cfg5: the 5-node
    H2 network, 1e5 replications x 1e4 network jobs, sharded (strong) or per GPU (weak)."""
    import torch
    import torch.distributed as tdist

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    torch.cuda.set_device(local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        tdist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    from most_queue_b200 import _lib
    from most_queue_b200.random.distributions import GammaDistribution, H2Distribution, ParetoDistribution
    from most_queue_b200.sim.lindley import run_chain
    from most_queue_b200.sim.networks.network import NetworkSimulator

    peaks, peak_src = _peaks()
    sms = torch.cuda.get_device_properties(local_rank).multi_processor_count

    def barrier():
        if world > 1:
            tdist.barrier()
        torch.cuda.synchronize()

    def maxreduce(x):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        tdist.all_reduce(t, op=tdist.ReduceOp.MAX)
        return float(t.item())

    if args.config == "cfg3":
        gam = GammaDistribution.get_params_by_mean_and_cv(1.0, 0.56)
        par = ParetoDistribution.get_params_by_mean_and_cv(0.7, 1.2)
        n = int(args.chain_jobs)

        def step(i):
            r = run_chain(("Gamma", gam), ("Pa", par), n, seed=300 + i, want_p=False)
            return r, (r.timing["sim_ms"] if r.timing else 0.0), n

        metric, unit = "simulated jobs/sec (GI/G/1 Gamma/Pareto chain, max-plus Lindley scan)", "jobs/s"
        workload = f"Gamma/Pareto/1 rho=0.7, ONE chain of {n} jobs cut into {world} contiguous range(s) (BASELINE cfg3)"
        scaling, i_alg, kernel = "strong", I_ALG_OTHER["cfg3"], "k3_summary + k3_walk (generator form)"
        h2d, d2h = 0, 8 * 56
    else:
        strong = args.scaling == "strong"
        reps = 100_000 if strong else 100_000 * world
        jobs = 10_000

        def step(i):
            net = NetworkSimulator(seed=500 + i, replications=reps)
            net.set_sources(1.0, np.array(NET_R, dtype=float))
            net.set_nodes([{"type": "H", "params": H2Distribution.get_params_by_mean_and_cv(0.7 * c, 1.2)} for c in NET_CH], NET_CH)
            r = net.run(jobs)
            return r, (net.timing["sim_ms"] if net.timing else 0.0), reps * jobs

        metric, unit = "simulated network jobs/sec (5-node open network of M/H2/c nodes)", "network jobs/s"
        workload = (f"5-node open network, H2 nodes c={NET_CH}, {reps} replications x {jobs} network jobs "
                    + ("in total, sharded" if strong else "(100000 per GPU)") + " (BASELINE cfg5)")
        scaling, i_alg, kernel = args.scaling, I_ALG_OTHER["cfg5"], "k5v2_net<16 server slots,generator>"
        h2d, d2h = int(__import__("ctypes").sizeof(_lib.MqNetCfg)), 8 * _lib.net_agg_len(5)

    for i in range(args.warmup):
        step(i)
    barrier()
    dev_ms, units, res = 0.0, 0, None
    with ClockSampler(local_rank) as clk:
        t0 = time.perf_counter()
        for i in range(args.steps):
            res, ms, u = step(args.warmup + i)
            dev_ms += ms
            units += u
        barrier()
        wall_s = time.perf_counter() - t0
    clocks = clk.summary()
    dev_ms = maxreduce(dev_ms)
    wall_s = maxreduce(wall_s)
    if rank == 0:
        value = units / (dev_ms * 1e-3)       # units are whole-job counts (every rank returns the same total)
        e2e = units / wall_s
        line = {"metric": metric, "value": value, "unit": unit, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
                "ms_per_step": 1e3 * wall_s / args.steps, "higher_is_better": True, "scaling": scaling, "vs_baseline": None,
                "dtype": "f32 variates, f64 recursion" if args.config == "cfg3" else "f64", "data": "synthetic",
                "config": {"workload": workload, "generator": "Philox2x32-10",
                           "cache": "generator mode: no input arrays",
                           "parallelism": (f"{world} contiguous job ranges, all-gather of {world} (a, b) pairs + one allreduce"
                                           if args.config == "cfg3" else f"replications sharded over {world} GPU(s), one allreduce")},
                "clocks": clocks,
                "e2e": {"value": e2e, "unit": unit, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                        "api": "run_chain(source, server, n)" if args.config == "cfg3" else "NetworkSimulator().run(job_served, replications=)"},
                "gpu_launches": args.steps * (7 if args.config == "cfg3" and world > 1 else (6 if args.config == "cfg3" else 2)),
                "roofline": issue_roofline(value / world, i_alg, kernel, peaks, peak_src, sms),
                "result": {"w1": float(res.w[0]), "v1": float(res.v[0])} if args.config == "cfg3" else {"v1": float(res.v[0])}}
        print(json.dumps(line), flush=True)
    if world > 1:
        tdist.destroy_process_group()


def replay_jobs_leg(ctx, peaks, peak_src, reps=65536, rows=4096, steps=3):
    """cfg 2r through K1b, the job-indexed replay kernel (csrc/k1b_fifo_jobs.cu): fp64 A, S of `reps` replications x
    `rows` jobs resident in HBM ([job][replication], 2 x 2 GiB), every row read exactly once; HBM roofline."""
    import torch

    from most_queue_b200 import _lib

    g = torch.Generator(device="cuda").manual_seed(1)
    A = -torch.log(1.0 - torch.rand((rows, reps), generator=g, device="cuda", dtype=torch.float64))
    br = torch.rand((rows, reps), generator=g, device="cuda", dtype=torch.float64) < H2["p1"]
    S = -torch.log(1.0 - torch.rand((rows, reps), generator=g, device="cuda", dtype=torch.float64))
    S = S / torch.where(br, H2["mu1"], H2["mu2"])
    del br
    torch.cuda.synchronize()
    ms = []
    for i in range(steps + 1):
        agg = ctx.replay_fifo_jobs_device(C_SERVERS, A.data_ptr(), S.data_ptr(), rows, reps)
        if i:
            ms.append(ctx.last_timing()["sim_ms"])
    t = float(np.mean(ms)) * 1e-3
    jobs = float(rows) * reps
    peak = float(peaks.get("hbm_gbs", 6650.0))
    achieved = 16.0 * jobs / t / 1e9
    # the same tier end to end: arrays in PINNED HOST memory, through the C ABI's host-pointer form (H2D copies of A and
    # S, kernel, reduction, D2H of the aggregate inside the timed wall-clock span).  PCIe bound: 16 B per job.
    e2e = None
    try:
        reps_h = 8192
        Ah = A[:, :reps_h].contiguous().cpu().pin_memory()
        Sh = S[:, :reps_h].contiguous().cpu().pin_memory()
        ctx.replay_fifo_jobs(C_SERVERS, Ah.numpy(), Sh.numpy())
        ts = []
        for _ in range(3):
            t0 = time.perf_counter()
            agg_h = ctx.replay_fifo_jobs(C_SERVERS, Ah.numpy(), Sh.numpy())[0]
            ts.append(time.perf_counter() - t0)
        th = float(np.median(ts))
        e2e = {"value": rows * reps_h / th, "unit": "jobs/s", "h2d_bytes_per_step": 16 * rows * reps_h,
               "d2h_bytes_per_step": 8 * (len(agg_h) + 12 * reps_h), "gbytes_per_s_over_pcie": 16.0 * rows * reps_h / th / 1e9,
               "workload": f"{reps_h} replications x {rows} jobs from pinned host arrays through mq_replay_fifo_jobs (host pointers)"}
        del Ah, Sh
    except Exception as e:  # noqa: BLE001
        e2e = {"error": repr(e)[:200]}
    return {"kernel": "k1b_fifo_jobs<8>", "jobs_per_s": jobs / t, "ms": t * 1e3, "e2e": e2e,
            "workload": f"{reps} replications x {rows} jobs, fp64 A,S resident in HBM ([job][replication]), M/H2/8 draws",
            "w1": float(agg[0] / reps), "v1": float(agg[4] / reps),
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": 16.0 * jobs * 1.0,
                         "note": "algorithmic bytes 16 B/job (one gap and one service time, each read once; rows are "
                                 "consumed by all lanes of a warp at the same step: 256 contiguous bytes per row and array); "
                                 "traffic: dram bytes of one launch, profiles/r02_ncu_k1b_summary.csv (4.30 GB per 2.68e8 jobs); peak " + peak_src}}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--reps", type=int, default=1_000_000, help="replications per GPU (cfg2: 1e6)")
    ap.add_argument("--jobs", type=int, default=100_000, help="jobs per replication (cfg2: 1e5)")
    ap.add_argument("--cpu-seconds", type=float, default=12.0)
    ap.add_argument("--ref-seconds", type=float, default=8.0)
    ap.add_argument("--ref-jobs", type=int, default=100_000,
                    help="jobs per replication of the reference class (one replication per core and step)")
    ap.add_argument("--ref-port", action="store_true", help="time the C port instead of the reference class")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-others", action="store_true", help="skip the other BASELINE configurations")
    ap.add_argument("--no-replay", action="store_true")
    ap.add_argument("--no-fp64", action="store_true", help="skip the fp64 A/B run of K2")
    ap.add_argument("--config", default="cfg2", choices=["cfg2", "cfg3", "cfg5"])
    ap.add_argument("--scaling", default="weak", choices=["weak", "strong"])
    ap.add_argument("--chain-jobs", type=float, default=1e10, help="--config cfg3: jobs of the one chain")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    elif args.config == "cfg2":
        run_ours(args)
    else:
        run_other_config(args)


if __name__ == "__main__":
    main()
