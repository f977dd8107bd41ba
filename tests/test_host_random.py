"""Host-side mirror of most_queue.random: the parameter fitters and theoretical moments reproduce the
reference's numbers (tests/golden/theory.json, written by tests/golden/make_golden.py from the unmodified
reference), and the Kendall-code validation behaves like create_distribution (random/utils/create.py:48-77)."""

import json
import os

import numpy as np
import pytest

from most_queue_b200.random import distributions as D
from most_queue_b200.random.utils.params import Cox2Params, ErlangParams, GaussianParams, UniformParams

with open(os.path.join(os.path.dirname(__file__), "golden", "theory.json")) as f:
    TH = json.load(f)


def test_fitters_match_reference():
    for key, want in TH["fits"].items():
        kind, mean, cv = key.split("_")
        mean, cv = float(mean), float(cv)
        if kind == "h2":
            p = D.H2Distribution.get_params_by_mean_and_cv(mean, cv)
            got = [p.p1, p.mu1, p.mu2]
        elif kind == "gamma":
            p = D.GammaDistribution.get_params_by_mean_and_cv(mean, cv)
            got = [p.mu, p.alpha]
        else:
            p = D.ParetoDistribution.get_params_by_mean_and_cv(mean, cv)
            got = [p.alpha, p.K]
        np.testing.assert_allclose(np.real(got), want, rtol=1e-12, err_msg=key)


def test_theoretical_moments_match_reference():
    L = TH["laws"]
    u = D.UniformDistribution.get_params_by_mean_and_cv(2.0, 0.4)
    np.testing.assert_allclose([u.mean, u.half_interval], L["uniform"]["params"], rtol=1e-14)
    np.testing.assert_allclose(D.UniformDistribution.calc_theory_moments(u, 3), L["uniform"]["moments"], rtol=1e-13)
    r = D.UniformDistribution.get_params(L["uniform"]["moments"])
    np.testing.assert_allclose([r.mean, r.half_interval], L["uniform"]["refit"], rtol=1e-12)
    n = D.NormalDistribution.get_params_by_mean_and_cv(5.0, 0.2)
    assert [n.mean, n.std_dev] == L["normal"]["params"]
    np.testing.assert_allclose(D.NormalDistribution.calc_theory_moments(n, 2), L["normal"]["moments"], rtol=1e-14)
    c = Cox2Params(p1=0.7, mu1=2.0, mu2=0.5)
    np.testing.assert_allclose(D.CoxDistribution.calc_theory_moments(c, 3), L["cox"]["moments"], rtol=1e-13)
    np.testing.assert_allclose(D.DeterministicDistribution.calc_theory_moments(1.7, 3), L["det"]["moments"], rtol=1e-14)
    e = D.ErlangDistribution.get_params_by_mean_and_cv(2.0, 0.5)
    assert [e.r, e.mu] == L["erlang"]["params"]
    np.testing.assert_allclose(D.ErlangDistribution.calc_theory_moments(e, 3), L["erlang"]["moments"], rtol=1e-13)
    np.testing.assert_allclose(D.ExpDistribution.calc_theory_moments(0.8, 3), L["exp"]["moments"], rtol=1e-13)


def test_kendall_codes_are_validated_like_the_reference():
    from most_queue_b200 import _lib

    with pytest.raises(ValueError):
        _lib.make_dist("X", 1.0)
    with pytest.raises(ValueError):  # no Kendall code for Weibull in create_distribution
        _lib.make_dist("Weibull", {"k": 2.0, "W": 1.0})
    d = _lib.make_dist("Weibull", {"k": 2.0, "W": 1.0}, weibull_ok=True)
    assert d.kind == _lib.KINDS["Weibull"] and list(d.p)[:2] == [2.0, 1.0]
    d = _lib.make_dist("Uniform", UniformParams(mean=2.0, half_interval=0.5))
    assert list(d.p)[:2] == [2.0, 0.5]
    d = _lib.make_dist("Norm", GaussianParams(5.0, 1.0))
    assert list(d.p)[:2] == [5.0, 1.0]
    d = _lib.make_dist("E", ErlangParams(r=3, mu=2.0))
    assert list(d.p)[:2] == [3.0, 2.0]


def test_create_distribution_factory_mirrors_the_reference():
    """random/utils/create.py:19-80: one class per Kendall code, ValueError for an unknown code."""
    from most_queue_b200.random.utils.create import create_distribution
    from most_queue_b200.random.utils.params import GammaParams, H2Params, ParetoParams

    cases = [("M", 1.5, D.ExpDistribution), ("H", H2Params(p1=0.4, mu1=1.0, mu2=3.0), D.H2Distribution),
             ("E", ErlangParams(r=3, mu=2.0), D.ErlangDistribution), ("Gamma", GammaParams(mu=2.0, alpha=1.5), D.GammaDistribution),
             ("C", Cox2Params(p1=0.7, mu1=2.0, mu2=0.5), D.CoxDistribution), ("Pa", ParetoParams(alpha=2.5, K=1.0), D.ParetoDistribution),
             ("Uniform", UniformParams(mean=2.0, half_interval=0.5), D.UniformDistribution),
             ("Norm", GaussianParams(mean=5.0, std_dev=1.0), D.NormalDistribution), ("D", 1.7, D.DeterministicDistribution)]
    for code, prm, cls in cases:
        d = create_distribution(prm, code, None)
        assert type(d) is cls and d.params is prm and d.type == code
    with pytest.raises(ValueError):
        create_distribution(1.0, "Q", None)
    with pytest.raises(NotImplementedError):
        create_distribution(None, "MAP", None)
