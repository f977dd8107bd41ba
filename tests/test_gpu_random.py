"""most_queue.random drop-in: distribution objects draw on the GPU (same counter-based streams and samplers as
the simulation kernels); sample moments against the laws' theoretical moments, and generate() == generate_batch()."""

import math

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _cases():
    from most_queue_b200.random import distributions as D
    from most_queue_b200.random.utils.params import Cox2Params, WeibullParams

    h2 = D.H2Distribution.get_params_by_mean_and_cv(5.6, 1.5)
    gam = D.GammaDistribution.get_params_by_mean_and_cv(2.0, 1.2)
    par = D.ParetoDistribution.get_params_by_mean_and_cv(0.7, 0.2)
    erl = D.ErlangDistribution.get_params_by_mean_and_cv(2.0, 0.5)
    uni = D.UniformDistribution.get_params_by_mean_and_cv(2.0, 0.4)
    nor = D.NormalDistribution.get_params_by_mean_and_cv(5.0, 0.2)
    cox = Cox2Params(p1=0.7, mu1=2.0, mu2=0.5)
    wb = WeibullParams(k=2.0, W=1.5)
    # the reference's Weibull: x = (-ln(u) W)^(-1/k); E[x^j] = W^(-j/k) Gamma(1 - j/k), finite for j < k
    wb_m1 = 1.5 ** (-0.5) * math.gamma(0.5)
    return [
        (D.ExpDistribution, 0.8, D.ExpDistribution.calc_theory_moments(0.8, 2)),
        (D.H2Distribution, h2, D.H2Distribution.calc_theory_moments(h2, 2)),
        (D.GammaDistribution, gam, D.GammaDistribution.calc_theory_moments(gam, 2)),
        (D.ParetoDistribution, par, D.ParetoDistribution.calc_theory_moments(par, 2)),
        (D.ErlangDistribution, erl, D.ErlangDistribution.calc_theory_moments(erl, 2)),
        (D.UniformDistribution, uni, D.UniformDistribution.calc_theory_moments(uni, 2)),
        (D.NormalDistribution, nor, D.NormalDistribution.calc_theory_moments(nor, 2)),
        (D.CoxDistribution, cox, D.CoxDistribution.calc_theory_moments(cox, 2)),
        (D.DeterministicDistribution, 1.7, [1.7, 1.7 ** 2]),
        (D.Weibull, wb, [wb_m1]),
    ]


def test_sample_moments_match_the_laws():
    n = 2_000_000
    for cls, params, moments in _cases():
        x = cls(params, seed=123).generate_batch(n).astype(np.float64)
        assert x.shape == (n,)
        for j, m in enumerate(moments):
            s = x ** (j + 1)
            half = 4.0 * s.std() / math.sqrt(n) + 1e-6 * abs(m)
            if cls.__name__ == "Weibull":
                half = 0.02 * m  # infinite variance at k = 2: loose check of the mean only
            assert abs(s.mean() - m) < half, (cls.__name__, j, s.mean(), m)


def test_generate_is_the_batch_one_at_a_time():
    from most_queue_b200.random.distributions import GammaDistribution

    p = GammaDistribution.get_params_by_mean_and_cv(2.0, 1.2)
    a = GammaDistribution(p, seed=9, stream=3)
    b = GammaDistribution(p, seed=9, stream=3)
    batch = a.generate_batch(1000)
    singles = np.array([b.generate() for _ in range(1000)], dtype=np.float32)
    assert np.array_equal(batch, singles)
    assert not np.array_equal(batch, GammaDistribution(p, seed=9, stream=4).generate_batch(1000))
    # a second batch continues the stream
    more = a.generate_batch(10)
    assert np.array_equal(more, np.array([b.generate() for _ in range(10)], dtype=np.float32))
