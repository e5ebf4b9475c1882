"""Short single-GPU target for ncu: the CTA-cooperative d = 40 backward filter on 4 intervals x 100 steps (unblocked)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import dmcmc_pkg
pkg = dmcmc_pkg.load()
from diffusionmcmc_jl_b200.engine import Engine
spec = pkg.problems.l96_spec(n_obs=4, C=1, dt=1e-4)
e = Engine(spec, seed=5)
e.backward_filter()
print("done")
