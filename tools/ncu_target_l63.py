"""Short, single-GPU ncu target for the generic small-d kernel: Lorenz-63 (d = 3, m = 3), blocked sweeps."""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import dmcmc_pkg
pkg = dmcmc_pkg.load()
from diffusionmcmc_jl_b200.engine import Engine

chains = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
nobs = int(sys.argv[2]) if len(sys.argv) > 2 else 400
sweeps = int(sys.argv[3]) if len(sys.argv) > 3 else 6
half = int(sys.argv[4]) if len(sys.argv) > 4 else 1
model = sys.argv[5] if len(sys.argv) > 5 else "l63"
mk = {"l63": pkg.problems.l63_spec, "lv": pkg.problems.lv_spec}[model]
spec = mk(n_obs=nobs, C=chains, dt=1e-3 if model == "l63" else 1e-2, rho=0.9)
e = Engine(spec, seed=1)
e.init_paths()
e.run_sweeps(0, 2, half)
e.kernel_times()
e.timer_start()
e.run_sweeps(10, sweeps, half)
region = e.timer_stop() / sweeps
ms, n = e.kernel_times()
steps = e.C * e.S
bps = 8 * (2 * e.m + e.d)
a, p = e.accept_counts()
print(json.dumps({"model": model, "chains": e.C, "obs": nobs, "half_len": half, "d": e.d, "m": e.m,
                  "ms_per_sweep": region, "path_steps_per_sec": steps / region * 1e3,
                  "algorithmic_GBps": bps * steps / region / 1e6, "frac_of_6556": bps * steps / region / 1e6 / 6556.2,
                  "accept_rate": float(a.sum() / max(1, p.sum()))}))
