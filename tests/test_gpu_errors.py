"""Error behaviour at the boundary: bad arguments raise ValueError like the reference, library limits
are reported (never silently wrong), short replay arrays are detected."""

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def test_invalid_arguments_raise_value_error():
    from most_queue_b200 import _lib
    from most_queue_b200.sim.base import QsSim

    qs = QsSim(0, seed=1)
    qs.set_sources(1.0, "M")
    qs.set_servers(1.0, "M")
    with pytest.raises(ValueError):
        qs.run(10)
    qs = QsSim(2, seed=1)
    qs.set_sources(-1.0, "M")       # negative rate
    qs.set_servers(1.0, "M")
    with pytest.raises(ValueError):
        qs.run(10)
    qs = QsSim(40, seed=1)          # more channels than the register-resident sorter covers
    qs.set_sources(1.0, "M")
    qs.set_servers(0.05, "M")
    with pytest.raises(_lib.MqError) as e:
        qs.run(10)
    assert e.value.code == _lib.MQ_E_UNSUPPORTED
    with pytest.raises(RuntimeError):
        QsSim(2).run(10)            # sources / servers not set


def test_priority_queue_overflow_is_flagged():
    from most_queue_b200 import _lib
    from most_queue_b200.sim.priority import PriorityQueueSimulator

    # overloaded system with a tiny engine-side waiting line: must be reported, not silently wrong
    qs = PriorityQueueSimulator(1, 2, "NP", seed=3, replications=64, queue_cap=4)
    qs.set_sources([{"type": "M", "params": 2.0}, {"type": "M", "params": 2.0}])
    qs.set_servers([{"type": "M", "params": 1.0}, {"type": "M", "params": 1.0}])
    with pytest.raises(_lib.MqError) as e:
        qs.run(2000)
    assert e.value.code == _lib.MQ_E_OVERFLOW


def test_short_replay_arrays_are_detected():
    from most_queue_b200 import _lib
    from most_queue_b200.sim.networks.network import replay as net_replay
    from most_queue_b200.sim.priority import replay as prio_replay

    with pytest.raises(_lib.MqError):
        prio_replay(2, 2, "PR", [np.ones((5, 2)), np.ones((5, 2))], [np.ones((5, 2)), np.ones((5, 2))], 100)
    R = [[1, 0], [0, 1]]
    with pytest.raises(_lib.MqError):
        net_replay(R, [2], np.ones((5, 3)), np.full((5, 3), 0.5), [np.ones((5, 3))], 100)
    with pytest.raises(ValueError):
        net_replay([[1, 0, 0], [0, 1, 0]], [2], np.ones((5, 3)), np.full((5, 3), 0.5), [np.ones((5, 3))], 1)


def test_network_source_routed_out_is_rejected():
    from most_queue_b200.sim.networks.network import NetworkSimulator

    sim = NetworkSimulator(seed=1)
    sim.set_sources(1.0, [[0.5, 0.5], [0, 1]])   # R[0][m] > 0: the reference would index qs[m]
    sim.set_nodes([{"type": "M", "params": 2.0}], [1])
    with pytest.raises(ValueError):
        sim.run(10)


def test_zero_jobs_and_single_replication_shapes():
    from most_queue_b200.sim.base import QsSim

    qs = QsSim(2, seed=5, p_bins=16)
    qs.set_sources(1.0, "M")
    qs.set_servers(1.0, "M")
    res = qs.run(1)
    assert qs.served == 1 and res.replications == 1 and len(res.w) == 4 and len(res.v) == 4
    assert all(np.isnan(res.ci["w"]))      # no across-replication spread with one replication
    assert abs(sum(res.p[:16]) + qs.p_over - 1.0) < 1e-6
