"""Small driver for ncu captures of the generator form of the Lindley scan (cfg3 laws, 2^28 jobs)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from most_queue_b200.random.distributions import GammaDistribution, ParetoDistribution
from most_queue_b200.sim.lindley import run_chain

n = 1 << int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 28
gam = GammaDistribution.get_params_by_mean_and_cv(1.0, 0.56)
par = ParetoDistribution.get_params_by_mean_and_cv(0.7, 1.2)
for _ in range(2):
    c = run_chain(("Gamma", gam), ("Pa", par), n, seed=3, want_p=False)
print(n, c.timing, n / (c.timing["sim_ms"] * 1e-3), c.w[0])
