"""Distribution parameter records.

Same names and field order as the reference's most_queue/random/utils/params.py:8-89, so code
written against the reference constructs them unchanged.  The engine also accepts the
reference's own dataclasses (duck-typed on field names).
"""

from dataclasses import dataclass


@dataclass
class H2Params:
    mu1: float
    mu2: float
    p1: float


@dataclass
class Cox2Params:
    mu1: float
    mu2: float
    p1: float


@dataclass
class GammaParams:
    mu: float
    alpha: float
    g: list | None = None


@dataclass
class WeibullParams:
    k: float
    W: float


@dataclass
class GaussianParams:
    mean: float
    std_dev: float


@dataclass
class UniformParams:
    mean: float
    half_interval: float


@dataclass
class ParetoParams:
    alpha: float
    K: float


@dataclass
class ErlangParams:
    r: int
    mu: float
