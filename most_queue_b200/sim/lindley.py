"""One very long GI/G/1 FIFO chain on the GPU: the max-plus Lindley parallel scan (csrc/k3_lindley.cu).

For one channel and an infinite buffer the reference's QsSim.run(N) (most_queue/sim/base.py:488-528)
is the Lindley recursion W[n+1] = max(0, W[n] + S[n] - A[n+1]); the reference can only walk it job by
job (base.py:505).  `run_chain` evaluates N jobs (N up to ~1e11) with a parallel scan and returns the
same QueueResults fields (w, v raw moments, utilisation); state probabilities are limited to
p[0] = 1 - busy fraction (the number-in-system process needs the merged event order, which the
scan does not build).

Multi-GPU: the chain is cut into contiguous ranges, one per rank; each rank publishes the max-plus
summary (a, b) of its range, the summaries are all-gathered (16 bytes per rank) and folded in rank
order into each rank's carry, then every rank walks its range.  The counter of the generator is the
GLOBAL job index, so the variates do not depend on the number of ranks.
"""

from __future__ import annotations

import os
import time

import numpy as np

from .. import _lib, dist
from ..random.utils.load import calc_qs_load
from ..structs import QueueResults


def _gather_pairs(pair):
    d = dist._dist()
    if d is None or d.get_world_size() == 1:
        return [pair]
    import torch

    dev = "cuda" if d.get_backend() == "nccl" else "cpu"
    t = torch.tensor(pair, dtype=torch.float64, device=dev)
    out = [torch.empty_like(t) for _ in range(d.get_world_size())]
    d.all_gather(out, t)
    return [tuple(float(v) for v in x.cpu()) for x in out]


def fold_carries(pairs, w0=0.0):
    """Carry into each range from the (a, b) summaries of the ranges before it, in range order."""
    carries, w = [], w0
    for a, b in pairs:
        carries.append(w)
        w = max(a, w + b)
    return carries, w


def levels_to_p(levels, ttek, p_bins=None):
    """Time-in-state levels T[N >= m], m = 1..32 (+ the lumped excess) -> p[0..31] and the tail P(N >= 32)."""
    L = _lib
    tge = np.asarray(levels[:L.SCAN_NLEVEL - 1], dtype=np.float64)
    p = np.empty(L.SCAN_NLEVEL - 1)
    p[0] = 1.0 - tge[0] / ttek
    p[1:] = (tge[:-1] - tge[1:]) / ttek
    return [float(x) for x in p], float(tge[-1] / ttek)


def run_chain(source, server, total_served: int, seed=None, chain: int = 0, device=None,
              want_p: bool = True) -> QueueResults:
    """source / server: (kendall_notation, params) as for QsSim.set_sources / set_servers.
    want_p: also settle the time-in-state levels (p[0..31]); costs roughly one more pass of arithmetic."""
    start = time.process_time()
    L = _lib
    src, srv = L.make_dist(source[0], source[1]), L.make_dist(server[0], server[1])
    seed = dist.common_seed(seed)  # one key for every range of the chain (seed=None: drawn on rank 0, broadcast)
    rank, world = dist.world()
    job0, count = dist.shard(int(total_served))
    ctx = L.default_context(device)
    summary_ms = 0.0
    if world == 1:
        carries = [0.0]  # one range: no summary exchange, the apply call computes its own tile prefixes
    else:
        # the summary keeps its draws and prefixes on the device: the apply below only walks (MQ_SCAN_KEEP)
        ab = (ctx.scan_summary(count, job0=job0, src=src, srv=srv, seed=seed, chain=chain, keep=True, want_p=want_p,
                               jobs_ahead=int(total_served) - (job0 + count)) if count > 0 else (-np.inf, 0.0))
        summary_ms = ctx.last_timing()["sim_ms"] if count > 0 else 0.0
        carries, _ = fold_carries(_gather_pairs(ab))
    if count > 0:
        stats, _ = ctx.scan_apply(count, carries[rank], job0=job0, src=src, srv=srv, seed=seed, chain=chain,
                                  last_range=(rank == world - 1), want_p=want_p,
                                  jobs_ahead=int(total_served) - (job0 + count))
        timing = dict(ctx.last_timing())
        timing["summary_ms"] = summary_ms      # the range summary exchanged before the walk (multi-rank runs)
        timing["sim_ms"] += summary_ms
    else:
        stats, timing = np.zeros(L.SCAN_LEN), None
    # everything that sums; the tail belongs to the last range only
    tail = np.array([stats[L.S_TAILV]]) if rank == world - 1 else np.zeros(1)
    levels = stats[L.S_LEVELS:L.S_LEVELS + L.SCAN_NLEVEL]
    summed = dist.allreduce_sum(np.concatenate([stats[:L.S_A], levels, tail]))
    res = _publish(summed[:L.S_A], float(summed[-1]), source, server)
    if want_p:
        res.p, res.p_tail = levels_to_p(summed[L.S_A:L.S_A + L.SCAN_NLEVEL], res.ttek)
    res.duration = time.process_time() - start
    res.timing = timing
    return res


def _publish(s, tail_v, source, server) -> QueueResults:
    L = _lib
    taked, served = s[L.S_TAKED], s[L.S_SERVED]
    ttek = s[L.S_SUMA] + tail_v  # last arrival time + (W + S) of the last job = clock at the N-th departure
    w = [float(x) for x in s[L.S_W:L.S_W + 4] / taked]
    v = [float(x) for x in s[L.S_V:L.S_V + 4] / served]
    p0 = float(1.0 - s[L.S_SUMS] / ttek)
    res = QueueResults(v=v, w=w, p=[p0], utilization=calc_qs_load(source[0], source[1], server[0], server[1], 1))
    res.ttek = float(ttek)
    res.taked, res.served = int(round(taked)), int(round(served))
    res.measured_utilization = float(s[L.S_SUMS] / ttek)
    return res


def replay_chain(A, S, total_served: int | None = None, ranges: int = 1, want_w: bool = True, device=None,
                 want_p: bool = False):
    """Replay tier: A[n] = gap before job n (N + 1 entries), S[n] = service of job n.  `ranges` > 1
    evaluates the chain as that many contiguous ranges with summary / carry / apply, exactly what a
    multi-GPU run does, on one GPU.  Returns (stats dict, W array)."""
    L = _lib
    A = np.ascontiguousarray(A, dtype=np.float64)
    S = np.ascontiguousarray(S, dtype=np.float64)
    N = int(total_served if total_served is not None else len(S))
    ctx = L.default_context(device)
    bounds = [N * g // ranges for g in range(ranges + 1)]
    pairs = [ctx.scan_summary(bounds[g + 1] - bounds[g], job0=bounds[g], A=A[bounds[g]:], S=S[bounds[g]:])
             for g in range(ranges)]
    carries, w_out = fold_carries(pairs)
    tot = np.zeros(L.S_A)
    levels = np.zeros(L.SCAN_NLEVEL)
    Ws, tail_v, tail_raw, sim_ms = [], 0.0, 0.0, 0.0
    pairs_apply = []  # the range summaries as the apply call reports them (the one-pass kernel forms its own)
    for g in range(ranges):
        n = bounds[g + 1] - bounds[g]
        stats, W = ctx.scan_apply(n, carries[g], job0=bounds[g], A=A[bounds[g]:], S=S[bounds[g]:],
                                  last_range=(g == ranges - 1), want_w=want_w, want_p=want_p,
                                  jobs_ahead=N - bounds[g + 1])
        levels += stats[L.S_LEVELS:L.S_LEVELS + L.SCAN_NLEVEL]
        sim_ms += ctx.last_timing()["sim_ms"]
        tot += stats[:L.S_A]
        pairs_apply.append((float(stats[L.S_A]), float(stats[L.S_B])))
        if want_w:
            Ws.append(W)
        tail_v, tail_raw = stats[L.S_TAILV], stats[L.S_TAILRAW]
    out = {"w_sum": tot[L.S_W:L.S_W + 4], "v_sum": tot[L.S_V:L.S_V + 4], "taked": int(tot[L.S_TAKED]),
           "served": int(tot[L.S_SERVED]), "sum_a": tot[L.S_SUMA], "sum_s": tot[L.S_SUMS], "pairs": pairs,
           "pairs_apply": pairs_apply, "carries": carries, "w_out": w_out, "tail_raw": tail_raw, "ttek": tot[L.S_SUMA] + tail_v,
           "sim_ms": sim_ms, "levels": levels}
    if want_p:
        out["p"], out["p_tail"] = levels_to_p(levels, out["ttek"])
    return out, (np.concatenate(Ws) if want_w else None)
