"""The C-ABI library loads on a machine without a GPU and exports every entry point that
include/mqb200.h declares; the product path refuses to run without a CUDA device (no CPU fallback)."""

import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    src = open(os.path.join(ROOT, "include", "mqb200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(mq_[a-z0-9_]+)\s*\(", src)))


def test_header_declares_the_hot_path_entry_points():
    names = _declared()
    for must in ("mq_create", "mq_destroy", "mq_sim_fifo", "mq_replay_fifo", "mq_sim_prio", "mq_replay_prio",
                 "mq_sim_net", "mq_replay_net", "mq_sim_tandem", "mq_replay_tandem", "mq_sim_prionet",
                 "mq_replay_prionet", "mq_sim_vacation", "mq_replay_vacation", "mq_sim_negatives", "mq_replay_negatives", "mq_sim_size", "mq_replay_size", "mq_scan_gg1", "mq_scan_gg1_summary", "mq_scan_gg1_apply",
                 "mq_last_error", "mq_version", "mq_device_count"):
        assert must in names


def test_library_exports_every_declared_symbol():
    from most_queue_b200 import _lib

    L = ctypes.CDLL(_lib.LIB_PATH)
    for name in _declared():
        assert hasattr(L, name), f"{name} declared in include/mqb200.h but not exported"
    assert _lib.load().mq_version() == 1


def test_struct_sizes_match_header_layout():
    from most_queue_b200 import _lib

    assert ctypes.sizeof(_lib.MqDist) == 40
    assert ctypes.sizeof(_lib.MqFifoCfg) == 8 + 8 + 80 + 8 * 4 + 8
    assert ctypes.sizeof(_lib.MqPrioCfg) == 16 + 8 * 5 + 8 + 40 * 16
    assert ctypes.sizeof(_lib.MqScanCfg) == 80 + 24 + 16 + 8
    # sizeof() printed by gcc for include/mqb200.h
    assert ctypes.sizeof(_lib.MqNetCfg) == 3104
    assert ctypes.sizeof(_lib.MqTandemCfg) == 856
    assert ctypes.sizeof(_lib.MqPrionetCfg) == 4272
    assert ctypes.sizeof(_lib.MqVacCfg) == 336
    assert ctypes.sizeof(_lib.MqNegCfg) == 184
    assert ctypes.sizeof(_lib.MqSizeCfg) == 144


def test_no_cpu_fallback_without_device():
    import torch

    from most_queue_b200 import _lib

    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    assert _lib.load().mq_device_count() == 0
    with pytest.raises(_lib.MqError) as e:
        _lib.Context(0)
    assert e.value.code == _lib.MQ_E_NODEVICE
    from most_queue_b200.sim.base import QsSim

    qs = QsSim(3, seed=1)
    qs.set_sources(2.0, "M")
    qs.set_servers(1.0, "M")
    with pytest.raises(_lib.MqError):
        qs.run(100)


def test_unknown_kendall_code_raises_value_error():
    from most_queue_b200.sim.base import QsSim

    qs = QsSim(2)
    with pytest.raises(ValueError):
        qs.set_sources(1.0, "XYZ")
    with pytest.raises(ValueError):
        QsSim(2, buffer_type="stack")
