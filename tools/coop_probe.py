"""A/B of the warp-per-unit (cooperative) imputation kernel against the thread-per-unit kernel on the
few-unit shapes of BASELINE configs 1-4: bitwise equality of the sampled state, device time per sweep."""
import sys, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import dmcmc_pkg
pkg = dmcmc_pkg.load()
from diffusionmcmc_jl_b200.engine import Engine
P = pkg.problems

CASES = [
    ("C1 FHN 1 chain x 50 obs, unblocked", lambda: P.fhn_spec(n_obs=50, C=1, dt=1e-3), 0),
    ("C1 FHN 1 chain x 50 obs, half_len 1", lambda: P.fhn_spec(n_obs=50, C=1, dt=1e-3), 1),
    ("C2 LV 1 chain x 200 obs, 8 blocks", lambda: P.lv_spec(n_obs=200, C=1), 12),
    ("C3 FHN 1024 chains x 100 obs, unblocked", lambda: P.fhn_spec(n_obs=100, C=1024, dt=1e-3), 0),
    ("C4 L63 1 chain x 10000 obs, half_len 5", lambda: P.l63_spec(n_obs=10000, C=1), 5),
    ("L63 8 chains x 2000 obs, unblocked", lambda: P.l63_spec(n_obs=2000, C=8), 0),
    ("FHN 1024 chains x 100 obs, half_len 25 (3072 units)", lambda: P.fhn_spec(n_obs=100, C=1024, dt=1e-3), 25),
    ("FHN 1024 chains x 100 obs, half_len 10 (6144 units)", lambda: P.fhn_spec(n_obs=100, C=1024, dt=1e-3), 10),
    ("FHN 1024 chains x 100 obs, half_len 5 (11264 units)", lambda: P.fhn_spec(n_obs=100, C=1024, dt=1e-3), 5),
    ("L63 64 chains x 10000 obs, half_len 50 (6464 units)", lambda: P.l63_spec(n_obs=10000, C=64), 50),
    ("L63 64 chains x 10000 obs, half_len 5 (64064 units)", lambda: P.l63_spec(n_obs=10000, C=64), 5),
    ("LVM 3 chains x 40 obs, half_len 2", lambda: P.lvm_spec(n_obs=40, C=3) if hasattr(P, "lvm_spec") else None, 2),
]
only = sys.argv[1:] and [int(a) for a in sys.argv[1:]]
for ci, (name, mk, half) in enumerate(CASES):
    if only and ci not in only:
        continue
    spec = mk()
    if spec is None:
        continue
    res, t = {}, {}
    for coop in (0, 1):
        e = Engine(spec, seed=3)
        e.lib.dmcmc_set_cooperative(e.h, coop)
        e.init_paths()
        e.run_sweeps(0, 6, half)
        X, W = e.download_state(0)
        res[coop] = (X, e.get_ll().copy(), e.accept_counts()[0].copy())
        n = 50
        e.timer_start()
        e.run_sweeps(100, n, half)
        t[coop] = e.timer_stop() / n
        steps = e.C * e.S
        e.close()
    a, b = res[0], res[1]
    print(json.dumps({"case": name, "units": int(len(b[1].ravel())), "steps_per_sweep": int(steps),
                      "us_thread_per_unit": round(t[0] * 1e3, 1), "us_warp_per_unit": round(t[1] * 1e3, 1),
                      "speedup": round(t[0] / t[1], 2), "path_steps_per_sec_coop": steps / t[1] * 1e3,
                      "bitwise_X": bool(np.array_equal(a[0], b[0])), "bitwise_ll": bool(np.array_equal(a[1], b[1])),
                      "bitwise_counts": bool(np.array_equal(a[2], b[2])),
                      "max_dX": float(np.max(np.abs(a[0] - b[0])))}), flush=True)
