"""Device time of the backward filter (a8) and of a critical parameter step on the reference's shapes."""
import sys, os, json, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import dmcmc_pkg
pkg = dmcmc_pkg.load()
from diffusionmcmc_jl_b200.engine import Engine
P = pkg.problems
for name, spec, half in [("C1 FHN 1 x 50 obs unblocked", P.fhn_spec(n_obs=50, C=1, dt=1e-3), 0),
                         ("C2 LV 1 x 200 obs, 8 blocks", P.lv_spec(n_obs=200, C=1), 12),
                         ("C3 FHN 1024 x 100 obs unblocked", P.fhn_spec(n_obs=100, C=1024, dt=1e-3), 0),
                         ("C4 L63 1 x 10000 obs half_len 5", P.l63_spec(n_obs=10000, C=1), 5)]:
    e = Engine(spec, seed=2)
    e.init_paths()
    if half:
        e.relayout(*P.chequerboard_layout(e.rec_nobs, half, 0))
    e.impute(0)
    out = {"case": name}
    for rep in range(2):
        e.timer_start(); e.backward_filter(0); out["bf_ms"] = e.timer_stop()
    th = np.array(spec["theta"], float)
    for rep in range(2):
        t0 = time.perf_counter()
        llp, ok = e.param_loglik(th * (1 + 1e-3), critical=True)
        e.param_commit(False)
        out["param_step_wall_ms"] = (time.perf_counter() - t0) * 1e3
    for rep in range(2):
        t0 = time.perf_counter(); e.impute(5 + rep); out["impute_wall_ms"] = (time.perf_counter() - t0) * 1e3
    print(json.dumps(out), flush=True)
    e.close()
