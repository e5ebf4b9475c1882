"""BASELINE config 5: Lorenz-96 (d = 40), partially observed (every other coordinate), chequerboard
blocking; noise never stored (store_noise = 0) so the state is 2 x X only."""
import argparse, json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import dmcmc_pkg
pkg = dmcmc_pkg.load()
from diffusionmcmc_jl_b200.engine import Engine

ap = argparse.ArgumentParser()
ap.add_argument("--obs", type=int, default=4096)
ap.add_argument("--chains", type=int, default=256)
ap.add_argument("--half-len", type=int, default=16)
ap.add_argument("--sweeps", type=int, default=6)
ap.add_argument("--dt", type=float, default=1e-4)
args = ap.parse_args()
t0 = time.perf_counter()
spec = pkg.problems.l96_spec(n_obs=args.obs, C=args.chains, dt=args.dt, rho=0.9)
t1 = time.perf_counter()
e = Engine(spec, seed=5, store_noise=False)
t2 = time.perf_counter()
e.init_paths()
t3 = time.perf_counter()
e.run_sweeps(0, 2, args.half_len)     # builds both layouts' tables
t4 = time.perf_counter()
e.kernel_times()
e.run_sweeps(10, args.sweeps, args.half_len)
ms, n = e.kernel_times()
a, p = e.accept_counts()
steps = e.C * e.S
per = ms[0] / n[0]
print(json.dumps({"config": "C5 Lorenz-96 d=40", "chains": e.C, "obs": args.obs, "steps_per_interval": int(e.N[0]),
                  "half_len": args.half_len, "ms_per_sweep": per, "path_steps_per_sec": steps / per * 1e3,
                  "algorithmic_GBps": 960 * steps / per / 1e6, "accept_rate": float(a.sum() / max(1, p.sum())),
                  "device_GB": e.device_bytes() / 1e9, "setup_s": {"data": t1 - t0, "engine+bf": t2 - t1,
                  "init_paths": t3 - t2, "first_sweeps+bf": t4 - t3}}))
