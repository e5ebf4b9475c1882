"""Kernel time by mode at the C3 blocked workload: relayout (accepted side only), stored-noise
imputation (proposal side only), fused."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import dmcmc_pkg
pkg = dmcmc_pkg.load()
from diffusionmcmc_jl_b200.engine import Engine
P = pkg.problems
spec = P.fhn_spec(n_obs=100, C=1024, dt=1e-3)
e = Engine(spec, seed=1)
if len(sys.argv) > 1:
    e.lib.dmcmc_set_threads(e.h, int(sys.argv[1]))
e.init_paths()
e.run_sweeps(0, 4, 1)
e.kernel_times()
for it in range(10):
    lay = P.chequerboard_layout(e.rec_nobs, 1, it % 2)
    e.relayout(*lay)          # MODE_RELAYOUT
    e.impute(it, want=False)  # stored-noise MODE_IMPUTE
ms, n = e.kernel_times()
print("relayout us", ms[2] / n[2] * 1e3, " stored-noise impute us", ms[0] / n[0] * 1e3)
e.run_sweeps(100, 20, 1)
ms, n = e.kernel_times()
print("fused us", ms[0] / n[0] * 1e3)
e.init_ll()
e.init_ll()
ms, n = e.kernel_times()
print("loglik us", ms[3] / n[3] * 1e3)
