"""Short single-GPU target for ncu: Lorenz-96 (C5 shape at reduced length), a few blocked sweeps."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import dmcmc_pkg
pkg = dmcmc_pkg.load()
from diffusionmcmc_jl_b200.engine import Engine
obs = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
chains = int(sys.argv[2]) if len(sys.argv) > 2 else 256
spec = pkg.problems.l96_spec(n_obs=obs, C=chains, dt=1e-4, rho=0.9)
e = Engine(spec, seed=5, store_noise=False)
e.init_paths()
e.run_sweeps(0, 2, 16)
e.timer_start()
e.run_sweeps(10, 4, 16)
print("ms per sweep", e.timer_stop() / 4)
