"""Theoretical utilisation lambda * b1 / n reported in the result struct.

Restates most_queue/random/utils/load.py:6-100 (QsSim.calc_load, sim/base.py:138-153): the value
is computed from the distribution PARAMETERS, it is not measured by the simulation.
"""


def mean_of(kind: str, params) -> float | None:
    """First moment of a distribution given its Kendall code and reference-style params."""
    if kind == "M":
        return 1.0 / params
    if kind == "D":
        return float(params)
    if kind in ("Uniform", "Norm"):
        return params.mean
    if kind == "H":
        return params.p1 / params.mu1 + (1.0 - params.p1) / params.mu2
    if kind == "E":
        return params.r / params.mu
    if kind == "Gamma":
        return params.alpha / params.mu
    if kind == "C":
        return (1.0 - params.p1) / params.mu1 + params.p1 * (1.0 / params.mu1 + 1.0 / params.mu2)
    if kind == "Pa":
        return params.alpha * params.K / (params.alpha - 1.0)
    return None


def calc_qs_load(source_kendall_notation, source_params, server_kendall_notation, server_params, n):
    a1 = mean_of(source_kendall_notation, source_params)
    b1 = mean_of(server_kendall_notation, server_params)
    lam = 0 if a1 is None else 1.0 / a1
    b1 = 0 if b1 is None else b1
    return lam * b1 / n
