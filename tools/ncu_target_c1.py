"""ncu target: BASELINE config 1 (FitzHugh-Nagumo, 1 chain x 50 obs, unblocked) -- warp-per-unit kernel and backward filter."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import dmcmc_pkg
pkg = dmcmc_pkg.load()
from diffusionmcmc_jl_b200.engine import Engine
spec = pkg.problems.fhn_spec(n_obs=50, C=1, dt=1e-3, rho=0.96)
e = Engine(spec, seed=1)
e.init_paths()
e.run_sweeps(0, 6, 0)
e.backward_filter(0)
print("ok")
