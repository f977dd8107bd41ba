"""Distribution factory: mirror of most_queue/random/utils/create.py:19-80.

`create_distribution(params, kendall_notation, generator)` returns the distribution object of the Kendall code, with
the reference's signature and its `ValueError` for an unknown code (create.py:76-78).  The objects are the engine's own
(most_queue_b200.random.distributions): `.generate()` is served from batches drawn on the GPU.  "PH" and "MAP" sources
(create.py:63-71) are outside the hot path this engine covers (M, H, E, Gamma, Pa, D, then C, Uniform, Norm)
and raise NotImplementedError, which says so.
"""

from ..distributions import (CoxDistribution, DeterministicDistribution, ErlangDistribution, ExpDistribution,
                             GammaDistribution, H2Distribution, NormalDistribution, ParetoDistribution,
                             UniformDistribution)

_BY_CODE = {
    "M": ExpDistribution, "H": H2Distribution, "E": ErlangDistribution, "Gamma": GammaDistribution,
    "C": CoxDistribution, "Pa": ParetoDistribution, "Uniform": UniformDistribution, "Norm": NormalDistribution,
    "D": DeterministicDistribution,
}


def create_distribution(params, kendall_notation: str, generator=None):
    """params: as for the reference (a float for "M" / "D", the parameter dataclass otherwise); generator: accepted for
    signature compatibility -- an object with a `.seed`-like integer state is not needed, the engine's streams are
    counter based (seed drawn from os.urandom unless the distribution is built with seed=)."""
    cls = _BY_CODE.get(kendall_notation)
    if cls is None:
        if kendall_notation in ("PH", "MAP"):
            raise NotImplementedError(f"'{kendall_notation}' sources are outside the hot path of this engine "
                                      "(most_queue/random/map_ph.py); use the reference's own classes")
        raise ValueError("Incorrect distribution type specified. See most_queue.random.distributions."
                         "print_supported_distributions() for list of supported distributions")
    return cls(params, generator=generator)
