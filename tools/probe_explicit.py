"""Kernel time of the fused blocked sweep with explicit (HBM) noise vs in-register Philox."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import dmcmc_pkg
pkg = dmcmc_pkg.load()
from diffusionmcmc_jl_b200.engine import Engine
P = pkg.problems
spec = P.fhn_spec(n_obs=100, C=1024, dt=1e-3)
e = Engine(spec, seed=1)
e.init_paths()
rng = np.random.default_rng(0)
Z = rng.standard_normal((e.C, e.S, e.m))
for mode in ("explicit", "philox"):
    e.kernel_times()
    for it in range(6):
        lay = P.chequerboard_layout(e.rec_nobs, 1, it % 2)
        e.set_layout(*lay)
        e.backward_filter()
        if mode == "explicit":
            E = rng.exponential(size=(e.C, len(lay[0])))
            e.impute(it, Z, E, want=False)
        else:
            e.impute(it, want=False)
    ms, n = e.kernel_times()
    print(mode, "impute kernel us:", ms[0] / n[0] * 1e3, "launches", n[0])
