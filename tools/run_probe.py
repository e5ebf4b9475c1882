"""MCMC iterations per second through the host mirror's run_ (the reference's run!) on the reference's own shapes."""
import sys, os, json, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import dmcmc_pkg
pkg = dmcmc_pkg.load()
P = pkg.problems

def timed(name, mcmc, n, data, theta=None, **kw):
    pkg.run_(mcmc(), 20, data, theta, seed=1, **kw)          # warm: context, tables, page faults
    t0 = time.perf_counter()
    gws, lws = pkg.run_(mcmc(), n, data, theta, seed=2, **kw)
    dt = time.perf_counter() - t0
    steps = int(np.asarray(data["N"]).sum()) * (kw.get("n_chains") or data["x0"].shape[0])
    print(json.dumps({"case": name, "mcmc_iters": n, "wall_s": round(dt, 3), "iters_per_sec": round(n / dt, 1),
                      "path_steps_per_sec": steps * n / dt,
                      "imputation_accept_rate": float(np.mean(lws[0].acceptance_history))}), flush=True)

c1 = P.fhn_spec(n_obs=50, C=1, dt=1e-3, rho=0.96)
timed("C1 smoothing, 1 chain, unblocked", lambda: pkg.MCMC([pkg.PathImputation(0.96, "FitzHughNagumoAux")]), 2000, c1)
timed("C1 smoothing, 1 chain, BlockingChequerboardSwitch(1)",
      lambda: pkg.MCMC([pkg.BlockingChequerboardSwitch(1), pkg.PathImputation(0.96, "FitzHughNagumoAux")]), 2000, c1)
timed("C1 + random-walk update of gamma (critical), unblocked",
      lambda: pkg.MCMC([pkg.PathImputation(0.96, "FitzHughNagumoAux"), pkg.RandomWalkUpdate(0.1, [2])]), 1000, c1, c1["theta"].copy())
c2 = P.lv_spec(n_obs=200, C=1)
timed("C2 LV smoothing + random-walk update, 8 blocks",
      lambda: pkg.MCMC([pkg.BlockingChequerboardSwitch(12), pkg.PathImputation(0.9), pkg.RandomWalkUpdate(0.05, [0], positive=True)]),
      500, c2, c2["theta"].copy())
