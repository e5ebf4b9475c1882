"""Device-time probe of the C3 sweep (1024 chains x 100 obs, half_len 1): CUDA events around n back-to-back sweeps."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import dmcmc_pkg
pkg = dmcmc_pkg.load()
from diffusionmcmc_jl_b200.engine import Engine
n = int(sys.argv[1]) if len(sys.argv) > 1 else 400
threads = int(sys.argv[2]) if len(sys.argv) > 2 else 64
spec = pkg.problems.fhn_spec(n_obs=100, C=1024, dt=1e-3)
e = Engine(spec, seed=1)
e.lib.dmcmc_set_threads(e.h, threads)
e.init_paths()
e.run_sweeps(0, 50, 1)
out = []
for rep in range(4):
    e.timer_start()
    e.run_sweeps(100 + rep * n, n, 1)
    out.append(e.timer_stop() / n * 1e3)
print("us/sweep:", " ".join(f"{x:.2f}" for x in out), f"-> best {min(out):.2f} us = {32*1.024e7/min(out)/1e3/6556.2*100:.1f}% HBM")
