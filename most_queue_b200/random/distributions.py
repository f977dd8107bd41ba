"""Host-side distribution helpers with the reference's class and method names.

The static parameter helpers the hot-path configurations use are provided
(get_params_by_mean_and_cv / get_params / calc_theory_moments).  Variates are drawn on the GPU
(csrc/mq_device.cuh): inside the simulation kernels, or -- for code that holds a distribution object
and calls .generate() as the reference's simulators do (distributions.py:104,1002, ...) -- in batches
through the C ABI (`generate_batch`), served one at a time by `generate()`.
Reference: most_queue/random/distributions.py.
"""

import math
import os

from .utils import fit
from .utils.params import (Cox2Params, ErlangParams, GammaParams, GaussianParams, H2Params, ParetoParams,
                           UniformParams, WeibullParams)


class Distribution:
    """Base of the distribution objects (distributions.py:30-73 in the reference): `params`, `type`,
    `generate()`.  The stream is counter based: variate i of (seed, stream) is a pure function of i, so
    a batch and a sequence of single draws give the same numbers."""

    type = None
    _BATCH = 1 << 16

    def __init__(self, params, generator=None, seed=None, stream=0, device=None):
        self.params = params
        self.generator = generator  # accepted for signature compatibility; the engine is counter based
        self.seed = seed if seed is not None else int.from_bytes(os.urandom(8), "little")
        self.stream = int(stream)
        self.device = device
        self._next = 0
        self._buf = None
        self._buf0 = 0

    def generate_batch(self, n: int, first: int | None = None):
        """Variates first .. first+n-1 of this object's stream as a float32 array (drawn on the GPU)."""
        from .. import _lib

        ctx = _lib.default_context(self.device)
        start = self._next if first is None else int(first)
        out = ctx.variates(_lib.make_dist(self.type, self.params, weibull_ok=True), self.seed, self.stream, 0, start + n)[start:]
        if first is None:
            self._next = start + n
        return out

    def generate(self) -> float:
        i = self._next
        if self._buf is None or not self._buf0 <= i < self._buf0 + len(self._buf):
            self._buf0 = i
            self._buf = self.generate_batch(self._BATCH, first=i)
        self._next = i + 1
        return float(self._buf[i - self._buf0])


class H2Distribution(Distribution):
    type = "H"

    @staticmethod
    def get_params(moments) -> H2Params:
        return fit.fit_h2(moments)

    @staticmethod
    def get_params_by_mean_and_cv(f1: float, cv: float) -> H2Params:
        return fit.h2_by_mean_and_cv(f1, cv)

    @staticmethod
    def calc_theory_moments(params: H2Params, num: int = 3):
        y2 = 1.0 - params.p1
        return [math.factorial(i + 1) * (params.p1 / params.mu1 ** (i + 1) + y2 / params.mu2 ** (i + 1))
                for i in range(num)]


class GammaDistribution(Distribution):
    type = "Gamma"

    @staticmethod
    def get_params(moments) -> GammaParams:
        return fit.fit_gamma(moments)

    @staticmethod
    def get_params_by_mean_and_cv(f1: float, cv: float) -> GammaParams:
        return fit.gamma_by_mean_and_cv(f1, cv)

    @staticmethod
    def calc_theory_moments(params: GammaParams, num: int = 3):
        out, prod = [], 1.0
        for i in range(num):
            prod *= params.alpha + i
            out.append(prod / params.mu ** (i + 1))
        return out


class ParetoDistribution(Distribution):
    type = "Pa"

    @staticmethod
    def get_params_by_mean_and_cv(f1: float, cv: float) -> ParetoParams:
        return fit.pareto_by_mean_and_cv(f1, cv)

    @staticmethod
    def calc_theory_moments(params: ParetoParams, num: int = 3):
        out = []
        for i in range(num):
            if params.alpha > i + 1:
                out.append(params.alpha * params.K ** (i + 1) / (params.alpha - i - 1))
            else:
                break
        return out


class ErlangDistribution(Distribution):
    type = "E"

    @staticmethod
    def get_params(moments) -> ErlangParams:
        return fit.fit_erlang(moments)

    @staticmethod
    def get_params_by_mean_and_cv(f1: float, cv: float) -> ErlangParams:
        return fit.erlang_by_mean_and_cv(f1, cv)

    @staticmethod
    def calc_theory_moments(params: ErlangParams, num: int = 3):
        out, prod = [], 1.0
        for i in range(num):
            prod *= params.r + i
            out.append(prod / params.mu ** (i + 1))
        return out


class ExpDistribution(Distribution):
    type = "M"

    @staticmethod
    def get_params_by_mean_and_cv(f1: float, cv: float = 1.0) -> float:
        return 1.0 / f1

    @staticmethod
    def calc_theory_moments(params: float, num: int = 3):
        return [math.factorial(i + 1) / params ** (i + 1) for i in range(num)]


class DeterministicDistribution(Distribution):
    type = "D"

    @staticmethod
    def get_params(moments):
        return moments[0]

    @staticmethod
    def get_params_by_mean_and_cv(f1: float, cv: float = 0.0) -> float:
        return f1

    @staticmethod
    def calc_theory_moments(params: float, num: int = 3):
        return [params ** (i + 1) for i in range(num)]


class UniformDistribution(Distribution):
    type = "Uniform"

    @staticmethod
    def get_params(moments) -> UniformParams:
        var = moments[1] - moments[0] * moments[0]
        return UniformParams(mean=moments[0], half_interval=math.sqrt(3.0 * var))

    @staticmethod
    def get_params_by_mean_and_cv(f1: float, cv: float) -> UniformParams:
        return UniformParams(mean=f1, half_interval=math.sqrt(3.0) * cv * f1)

    @staticmethod
    def calc_theory_moments(params: UniformParams, num: int = 3):
        lo, hi = params.mean - params.half_interval, params.mean + params.half_interval
        return [(hi ** (i + 2) - lo ** (i + 2)) / ((hi - lo) * (i + 2)) for i in range(num)]


class NormalDistribution(Distribution):
    type = "Norm"

    @staticmethod
    def get_params(moments) -> GaussianParams:
        return GaussianParams(moments[0], math.sqrt(moments[1] - moments[0] ** 2))

    @staticmethod
    def get_params_by_mean_and_cv(f1: float, cv: float) -> GaussianParams:
        return GaussianParams(f1, cv * f1)

    @staticmethod
    def calc_theory_moments(params: GaussianParams, num: int = 3):
        m, s2 = params.mean, params.std_dev ** 2
        full = [m, m * m + s2, m ** 3 + 3 * m * s2, m ** 4 + 6 * m * m * s2 + 3 * s2 * s2]
        return full[:num]


class CoxDistribution(Distribution):
    type = "C"

    @staticmethod
    def calc_theory_moments(params: Cox2Params, num: int = 3):
        # phase 1 ~ Exp(mu1); with probability p1 a second phase ~ Exp(mu2) follows (distributions.py:542-560)
        a, b, p = 1.0 / params.mu1, 1.0 / params.mu2, params.p1
        full = [a + p * b,
                2 * a * a + p * (2 * a * b + 2 * b * b),
                6 * a ** 3 + p * (6 * a * a * b + 6 * a * b * b + 6 * b ** 3),
                24 * a ** 4 + p * (24 * a ** 3 * b + 24 * a * a * b * b + 24 * a * b ** 3 + 24 * b ** 4)]
        return full[:num]


class Weibull(Distribution):
    """The reference's parametrisation (distributions.py:94-102): x = (-ln(u) * W) ** (-1 / k)."""

    type = "Weibull"
