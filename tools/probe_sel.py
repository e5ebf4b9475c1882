import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import dmcmc_pkg
pkg = dmcmc_pkg.load()
from diffusionmcmc_jl_b200.engine import Engine
spec = pkg.problems.fhn_spec(n_obs=100, C=1024, dt=1e-3)
e = Engine(spec, seed=1)
e.init_paths()
e.set_rho(np.full(e.J, float(sys.argv[1])))
e.run_sweeps(0, 8, 1)
ms, n = e.kernel_times()
a, p = e.accept_counts()
print("rho", sys.argv[1], "fused us", ms[0] / n[0] * 1e3, "acc", a.sum() / p.sum())
