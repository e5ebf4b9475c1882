"""Short, single-GPU target for ncu: the C3 bench workload, a handful of sweeps."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import dmcmc_pkg
pkg = dmcmc_pkg.load()
from diffusionmcmc_jl_b200.engine import Engine
from diffusionmcmc_jl_b200.backend import ImputationRunner

chains = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
sweeps = int(sys.argv[2]) if len(sys.argv) > 2 else 6
half = int(sys.argv[3]) if len(sys.argv) > 3 else 1
spec = pkg.problems.fhn_spec(n_obs=100, C=chains, dt=1e-3, rho=0.96, seed=1)
e = Engine(spec, seed=1)
e.init_paths()
r = ImputationRunner(e, half_len=half)
r.run_device(0, sweeps)
print(r.kernel_breakdown())
