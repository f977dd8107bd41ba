"""Known-answer tests for the counter-based generator of the oracle (Random123 KAT
vectors for philox2x32-10 and philox4x32-10) and distribution sanity checks."""

import math

import numpy as np

from oracle import oracle


def test_philox2x32_kat():
    assert oracle.philox2x32(0, 0, 0) == (0xFF1DAE59, 0x6CD10DF2)
    assert oracle.philox2x32(0xFFFFFFFF, 0xFFFFFFFF, 0xFFFFFFFF) == (0x2C3F628B, 0xAB4FD7AD)
    assert oracle.philox2x32(0x243F6A88, 0x85A308D3, 0x13198A2E) == (0xDD7CE038, 0xF62A4C12)


def test_philox4x32_kat():
    assert oracle.philox4x32((0, 0, 0, 0), (0, 0)) == (0x6627E8D5, 0xE169C58D, 0xBC57AC4C, 0x9B00DBD8)
    assert oracle.philox4x32((0xFFFFFFFF,) * 4, (0xFFFFFFFF,) * 2) == (
        0x408F276D, 0x41C83B0E, 0xA20BC7C6, 0x6D5451FD)
    assert oracle.philox4x32((0x243F6A88, 0x85A308D3, 0x13198A2E, 0x03707344), (0xA4093822, 0x299F31D0)) == (
        0xD16CFE09, 0x94FDCCEB, 0x5001E420, 0x24126EA1)


class _P:
    def __init__(self, **k):
        self.__dict__.update(k)


def _moments(dist, n=40000, which=1):
    x = np.array([oracle.variate(dist, 7, 3, i, which) for i in range(n)])
    return x.mean(), x.std() / x.mean()


def test_variate_means_and_cv():
    m, cv = _moments(("M", 2.0))
    assert abs(m - 0.5) < 0.01 and abs(cv - 1.0) < 0.03
    m, cv = _moments(("H", _P(p1=0.447586, mu1=0.0950716, mu2=0.619214)))
    assert abs(m - 5.6) < 0.15 and abs(cv - 1.5) < 0.06
    m, cv = _moments(("E", _P(r=4, mu=2.0)))
    assert abs(m - 2.0) < 0.03 and abs(cv - 0.5) < 0.02
    m, cv = _moments(("Gamma", _P(mu=3.18878, alpha=3.18878)), which=0)
    assert abs(m - 1.0) < 0.015 and abs(cv - 0.56) < 0.02
    m, cv = _moments(("Gamma", _P(mu=0.5, alpha=0.5)))
    assert abs(m - 1.0) < 0.03 and abs(cv - math.sqrt(2)) < 0.06
    m, _ = _moments(("Pa", _P(alpha=2.30171, K=0.395878)))
    assert abs(m - 0.7) < 0.03
    m, cv = _moments(("D", 1.25))
    assert m == 1.25 and cv == 0.0
    m, cv = _moments(("C", _P(p1=0.3, mu1=1.0, mu2=0.5)))
    assert abs(m - (1.0 + 0.3 * 2.0)) < 0.04
    m, cv = _moments(("Uniform", _P(mean=2.0, half_interval=1.0)))
    assert abs(m - 2.0) < 0.01 and abs(cv - (1 / math.sqrt(3)) / 2.0) < 0.01
    m, cv = _moments(("Norm", _P(mean=5.0, std_dev=0.5)))
    assert abs(m - 5.0) < 0.01 and abs(cv - 0.1) < 0.003
