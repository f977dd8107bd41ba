"""Host-side distribution helpers with the reference's class and method names.

Only the static parameter helpers the hot-path configurations use are provided
(get_params_by_mean_and_cv / get_params / calc_theory_moments); variates themselves are drawn
on the GPU (csrc/mq_device.cuh).  Reference: most_queue/random/distributions.py.
"""

import math

from .utils import fit
from .utils.params import ErlangParams, GammaParams, H2Params, ParetoParams


class H2Distribution:
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


class GammaDistribution:
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


class ParetoDistribution:
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


class ErlangDistribution:
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


class ExpDistribution:
    @staticmethod
    def get_params_by_mean_and_cv(f1: float, cv: float = 1.0) -> float:
        return 1.0 / f1

    @staticmethod
    def calc_theory_moments(params: float, num: int = 3):
        return [math.factorial(i + 1) / params ** (i + 1) for i in range(num)]
