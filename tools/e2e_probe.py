"""Where does the end-to-end (host buffers) path lose time?  Times run_sweeps on the C3 workload with
and without the per-sweep rho H2D and the history D2H."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import dmcmc_pkg
pkg = dmcmc_pkg.load()
from diffusionmcmc_jl_b200.engine import Engine
from diffusionmcmc_jl_b200.backend import ImputationRunner

spec = pkg.problems.fhn_spec(n_obs=100, C=1024, dt=1e-3, rho=0.96, seed=1)
e = Engine(spec, seed=1)
e.init_paths()
r = ImputationRunner(e, half_len=1)
n = 500
r.reserve_host(n)
rho, llp, ll, acc = (a[:n] for a in r._host)
rho[...] = 0.96
e.run_sweeps(0, 20, 1, 0)
for name, rs, hist in [("device only", None, None), ("rho H2D only", rho, None),
                       ("hist D2H only", None, (llp, ll, acc)), ("both", rho, (llp, ll, acc)),
                       ("llp only", None, (llp, None, None))]:
    for rep in range(2):
        t0 = time.perf_counter()
        if hist is not None and hist[1] is None:
            from diffusionmcmc_jl_b200.engine import _dp, _bp
            e._ck(e.lib.dmcmc_run_sweeps(e.h, 1000, n, 1, 0, None, llp.shape[2], _dp(llp), None, None))
        else:
            e.run_sweeps(1000, n, 1, 0, rs, hist)
        dt = (time.perf_counter() - t0) / n
    ms, k = e.kernel_times()
    print(f"{name:14s}: {dt*1e6:7.1f} us/sweep wall, kernel {ms[0]/max(1,k[0])*1e3:6.1f} us", flush=True)
