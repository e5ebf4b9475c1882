"""A/B of the warp-specialised FitzHugh-Nagumo kernel against the single-role kernel: bitwise equality of
the sampled state after a few blocked sweeps, and device time per sweep (C3 workload)."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import dmcmc_pkg
pkg = dmcmc_pkg.load()
from diffusionmcmc_jl_b200.engine import Engine
P = pkg.problems

C = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
nobs = int(sys.argv[2]) if len(sys.argv) > 2 else 100
half_len = int(sys.argv[3]) if len(sys.argv) > 3 else 1
check = "--no-check" not in sys.argv
spec = P.fhn_spec(n_obs=nobs, C=C, dt=1e-3)
steps = C * nobs * 100
res = {}
for level in (1, 2):
    e = Engine(spec, seed=1)
    e.lib.dmcmc_set_fast_step(e.h, level)
    e.init_paths()
    e.run_sweeps(0, 6, half_len)
    if check:
        X, _ = e.download_state(0)
        res[level] = (X, e.get_ll().copy(), e.accept_counts())
    e.kernel_times()
    n = 200
    e.timer_start()
    e.run_sweeps(100, n, half_len)
    region = e.timer_stop() / n
    ms, k = e.kernel_times()
    imp = ms[0] / max(1, k[0])
    print(f"fast_step={level}: impute {imp*1e3:.1f} us/launch, region {region*1e3:.1f} us/sweep -> "
          f"{32*steps/region/1e6:.0f} GB/s ({32*steps/region/1e6/6556.2*100:.1f}% HBM)", flush=True)
    e.close()
if check:
    a, b = res[1], res[2]
    print("bitwise X:", np.array_equal(a[0], b[0]), " ll:", np.array_equal(a[1], b[1]),
          " counts:", np.array_equal(a[2][0], b[2][0]), " max|dX|:", float(np.max(np.abs(a[0] - b[0]))))
