"""Launches tests/mgpu/run_check.py under torchrun when at least two GPUs are visible."""

import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_two_rank_sharding_matches_single_process():
    import torch

    n = torch.cuda.device_count()
    if n < 2:
        pytest.skip("needs >= 2 GPUs")
    r = subprocess.run(
        [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr",
         "127.0.0.1", "--master-port", "29511", os.path.join(ROOT, "tests", "mgpu", "run_check.py")],
        capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert "MGPU_CHECK_OK" in r.stdout


def test_multi_context_call_equals_one_context():
    """mq_sim_fifo_multi: two contexts (here on the same GPU, one host thread each inside the library) share the
    replications of one call; the aggregate equals the single-context call (global replication ids in the counters)."""
    import numpy as np

    from most_queue_b200 import _lib
    from most_queue_b200.random.utils.params import H2Params

    h2 = H2Params(p1=0.447586, mu1=0.0950716, mu2=0.619214)
    src, srv = _lib.make_dist("M", 1.0), _lib.make_dist("H", h2)
    a, b, c = _lib.Context(0), _lib.Context(0), _lib.Context(0)
    try:
        one, _ = a.sim_fifo(8, None, src, srv, 2000, 5000, 7, 42, p_bins=32, busy=True)
        multi, _ = a.sim_fifo_multi([b, c], 8, None, src, srv, 2000, 5000, 7, 42, p_bins=32, busy=True)
    finally:
        for x in (a, b, c):
            x.close()
    np.testing.assert_allclose(multi, one, rtol=1e-12, atol=1e-9)
    assert multi[_lib.A_REPS] == 5000


def test_single_process_uses_every_gpu():
    """QsSim / PriorityQueueSimulator / NetworkSimulator with devices="all": one Python process, all visible GPUs."""
    import numpy as np
    import torch

    if torch.cuda.device_count() < 2:
        pytest.skip("needs >= 2 GPUs")
    from most_queue_b200.sim.base import QsSim
    from most_queue_b200.sim.networks.network import NetworkSimulator
    from most_queue_b200.sim.priority import PriorityQueueSimulator

    def fifo(**kw):
        qs = QsSim(3, buffer=30, seed=3, replications=6000, p_bins=48, **kw)
        qs.set_sources(1.0, "M")
        qs.set_servers(1.0 / 2.1, "M")
        return qs.run(3000)

    r1, rn = fifo(), fifo(devices="all")
    np.testing.assert_allclose(rn.w, r1.w, rtol=1e-12)
    np.testing.assert_allclose(rn.p[:20], r1.p[:20], rtol=1e-12, atol=1e-15)

    def prio(**kw):
        ps = PriorityQueueSimulator(3, 2, "PR", seed=4, replications=3000, p_bins=32, **kw)
        ps.set_sources([{"type": "M", "params": 0.4}, {"type": "M", "params": 0.5}])
        ps.set_servers([{"type": "M", "params": 0.6}, {"type": "M", "params": 0.5}])
        return ps.run(2000)

    p1, pn = prio(), prio(devices="all")
    np.testing.assert_allclose(pn.v, p1.v, rtol=1e-12)

    def net(**kw):
        n = NetworkSimulator(seed=5, replications=3000, **kw)
        n.set_sources(1.0, np.array([[1, 0, 0], [0, 0.5, 0.5], [0, 0, 1]], dtype=float))
        n.set_nodes([{"type": "M", "params": 1.5}, {"type": "M", "params": 2.5}], [2, 1])
        return n.run(1500)

    n1, nn = net(), net(devices="all")
    np.testing.assert_allclose(nn.v, n1.v, rtol=1e-12)
