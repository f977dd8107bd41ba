"""Run under torchrun with >= 2 GPUs: sharded runs must reproduce the single-process answer.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 \
        --master-port 29511 tests/mgpu/run_check.py
"""

import os
import sys

import numpy as np
import torch
import torch.distributed as tdist

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)


def main():
    rank, local, world = int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"]), int(os.environ["WORLD_SIZE"])
    torch.cuda.set_device(local)
    tdist.init_process_group("nccl", device_id=torch.device("cuda", local))
    from most_queue_b200 import _lib
    from most_queue_b200.random.utils.params import GammaParams, H2Params, ParetoParams
    from most_queue_b200.sim.base import QsSim
    from most_queue_b200.sim.lindley import run_chain

    h2 = H2Params(p1=0.447586, mu1=0.0950716, mu2=0.619214)
    R, N = 4096, 3000
    qs = QsSim(8, seed=99, p_bins=32)
    qs.set_sources(1.0, "M")
    qs.set_servers(h2, "H")
    sharded = qs.run(N, replications=R)           # shards over the ranks + NCCL allreduce
    agg_sharded = qs.agg.copy()
    # the same replications in one process on this rank's GPU (no process group involvement)
    ctx = _lib.default_context(local)
    agg_single, _ = ctx.sim_fifo(8, None, _lib.make_dist("M", 1.0), _lib.make_dist("H", h2), N, R, 0, 99, p_bins=32)
    np.testing.assert_allclose(agg_sharded, agg_single, rtol=1e-12, atol=1e-9)
    assert agg_sharded[_lib.A_REPS] == R and sharded.replications == R

    # Lindley chain: contiguous ranges + all-gather of (a, b) summaries + carry fold
    src, srv = ("Gamma", GammaParams(mu=3.18878, alpha=3.18878)), ("Pa", ParetoParams(alpha=2.30171, K=0.395878))
    Nj = 3_000_001
    multi = run_chain(src, srv, Nj, seed=5)
    lo = _lib.default_context(local)
    a, b = lo.scan_summary(Nj, src=_lib.make_dist(*src), srv=_lib.make_dist(*srv), seed=5)
    stats, _ = lo.scan_apply(Nj, 0.0, src=_lib.make_dist(*src), srv=_lib.make_dist(*srv), seed=5)
    w1 = stats[_lib.S_W] / stats[_lib.S_TAKED]
    v1 = stats[_lib.S_V] / stats[_lib.S_SERVED]
    assert multi.served == Nj
    np.testing.assert_allclose(multi.w[0], w1, rtol=1e-10)
    np.testing.assert_allclose(multi.v[0], v1, rtol=1e-10)
    # priority (cfg4 shape) and network (cfg5 shape): replications sharded over the ranks == one process
    from most_queue_b200.random.distributions import GammaDistribution, H2Distribution
    from most_queue_b200.sim.networks.network import NetworkSimulator
    from most_queue_b200.sim.priority import PriorityQueueSimulator

    g = [GammaDistribution.get_params_by_mean_and_cv(m, 1.2) for m in (2.0, 4.0)]
    ps = PriorityQueueSimulator(4, 2, "PR", seed=17, replications=2048, p_bins=32)
    ps.set_sources([{"type": "M", "params": 0.3}, {"type": "M", "params": 0.5}])
    ps.set_servers([{"type": "Gamma", "params": x} for x in g])
    ps.run(2000)
    pa = ctx.sim_prio(4, 2, "PR", None, [_lib.make_dist("M", 0.3), _lib.make_dist("M", 0.5)],
                      [_lib.make_dist("Gamma", x) for x in g], 2000, 2048, 0, 17, p_bins=32)[0]
    np.testing.assert_allclose(ps.agg, pa, rtol=1e-12, atol=1e-9)
    Rm = [[1, 0, 0, 0, 0, 0], [0, 0.4, 0.6, 0, 0, 0], [0, 0, 0.2, 0.4, 0.4, 0], [0, 0, 0, 0, 1, 0], [0, 0, 0, 0, 1, 0],
          [0, 0, 0, 0, 0, 1]]
    ch = [3, 2, 3, 4, 3]
    net = NetworkSimulator(seed=23, replications=2048)
    net.set_sources(1.0, np.array(Rm, dtype=float))
    hs = [H2Distribution.get_params_by_mean_and_cv(0.7 * c, 1.2) for c in ch]
    net.set_nodes([{"type": "H", "params": h} for h in hs], ch)
    net.run(1000)
    na = ctx.sim_net(Rm, ch, _lib.make_dist("M", 1.0), [_lib.make_dist("H", h) for h in hs], 1000, 2048, 0, 23)[0]
    np.testing.assert_allclose(net.agg, na, rtol=1e-12, atol=1e-9)
    # seed=None: rank 0's seed is broadcast, so every rank reports the SAME result
    q2 = QsSim(3, seed=None, p_bins=16)
    q2.set_sources(2.0, "M")
    q2.set_servers(1.0, "M")
    r2 = q2.run(500, replications=256)
    t = torch.tensor([r2.w[0]], dtype=torch.float64, device="cuda")
    lo_, hi_ = t.clone(), t.clone()
    tdist.all_reduce(lo_, op=tdist.ReduceOp.MIN)
    tdist.all_reduce(hi_, op=tdist.ReduceOp.MAX)
    assert float(lo_) == float(hi_)
    tdist.barrier()
    if rank == 0:
        print(f"MGPU_CHECK_OK world={world} w1={multi.w[0]:.6f} v1={multi.v[0]:.6f}")
    tdist.destroy_process_group()


if __name__ == "__main__":
    main()
