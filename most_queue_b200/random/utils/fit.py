"""Moments -> parameters helpers used to build the benchmark / test configurations on the host.

Only the closed forms and the one bisection the hot-path configurations need are restated here
(reference: most_queue/random/utils/fit.py:35-82 fit_h2, :183-208 Pareto, :211-217 Erlang,
:266-270 Gamma, :318-329 gamma_moments_by_mean_and_cv).  They are pinned against the reference's
outputs in tests/golden/theory.json ("fits").
"""

import math

from .params import ErlangParams, GammaParams, H2Params, ParetoParams


def gamma_moments_by_mean_and_cv(mean: float, cv: float) -> list[float]:
    """First four raw moments of the Gamma law with this mean and coefficient of variation."""
    shape = 1.0 / (cv * cv)
    m1 = mean
    m2 = m1 * m1 * (cv * cv + 1.0)
    m3 = m2 * m1 * (1.0 + 2.0 / shape)
    m4 = m3 * m1 * (1.0 + 3.0 / shape)
    return [m1, m2, m3, m4]


def fit_h2(moments) -> H2Params:
    """Real-parameter H2 matched to three raw moments (Aliev's bisection on the branch weight)."""
    m1, m2, m3 = moments[0], moments[1], moments[2]
    cv = math.sqrt(m2 - m1 * m1) / m1
    if cv < 1.0:
        return H2Params(p1=0, mu1=0, mu2=0)
    c2m1 = cv * cv - 1.0

    def means(qv):
        hi = (1.0 + math.sqrt((1.0 - qv) * c2m1 / (2.0 * qv))) * m1
        lo = (1.0 - math.sqrt(qv * c2m1 / (2.0 * (1.0 - qv)))) * m1
        return hi, lo

    q_hi = 2.0 / (1.0 + cv * cv)
    q_lo = 0.0
    if 1.5 * (1.0 + cv * cv) ** 2 * m1**3 > m3:
        # third moment below the H2-feasible range: degenerate second phase
        _, lo = means(q_hi)
        return H2Params(p1=q_hi, mu1=1e10 if math.isclose(lo, 0.0) else 1.0 / lo, mu2=1e6)
    third = 0.0
    t_hi = t_lo = 0.0
    for _ in range(10000):
        if abs(third - m3) <= 1e-8:
            break
        qv = 0.5 * (q_hi + q_lo)
        t_hi, t_lo = means(qv)
        third = 6.0 * (qv * t_hi**3 + (1.0 - qv) * t_lo**3)
        if third - m3 > 0:
            q_lo = qv
        else:
            q_hi = qv
    return H2Params(p1=q_hi, mu1=1.0 / t_hi, mu2=1.0 / t_lo)


def h2_by_mean_and_cv(mean: float, cv: float) -> H2Params:
    return fit_h2(gamma_moments_by_mean_and_cv(mean, cv))


def gamma_by_mean_and_cv(mean: float, cv: float) -> GammaParams:
    var = (mean * cv) ** 2
    mu = mean / var
    return GammaParams(mu=mu, alpha=mu * mean)


def fit_gamma(moments) -> GammaParams:
    var = moments[1] - moments[0] * moments[0]
    mu = moments[0] / var
    return GammaParams(mu=mu, alpha=mu * moments[0])


def pareto_by_mean_and_cv(mean: float, cv: float) -> ParetoParams:
    ratio = mean * mean / (mean * cv) ** 2
    alpha = (2.0 + math.sqrt(4.0 * (1.0 + ratio))) / 2.0
    return ParetoParams(alpha=alpha, K=(alpha - 1.0) * mean / alpha)


def fit_erlang(moments) -> ErlangParams:
    r = int(math.floor(moments[0] ** 2 / (moments[1] - moments[0] ** 2) + 0.5))
    return ErlangParams(r=r, mu=r / moments[0])


def erlang_by_mean_and_cv(mean: float, cv: float) -> ErlangParams:
    return fit_erlang([mean, (cv * cv + 1.0) * mean * mean])
