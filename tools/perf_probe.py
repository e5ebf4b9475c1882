"""Quick device-time probe of the imputation kernel on the C3 workload (not the bench)."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import dmcmc_pkg
pkg = dmcmc_pkg.load()
from diffusionmcmc_jl_b200.engine import Engine
P = pkg.problems

C = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
nobs = int(sys.argv[2]) if len(sys.argv) > 2 else 100
hls = [int(x) for x in sys.argv[3].split(",")] if len(sys.argv) > 3 else [0, 1, 2, 5]
ths = [int(x) for x in sys.argv[4].split(",")] if len(sys.argv) > 4 else [32, 64, 128]
spec = P.fhn_spec(n_obs=nobs, C=C, dt=1e-3)
steps = C * nobs * 100
for half_len in hls:
    for threads in (ths if half_len else ths[:1]):
        e = Engine(spec, seed=1)
        e.lib.dmcmc_set_threads(e.h, threads)
        e.init_paths()
        e.run_sweeps(0, 4, half_len)
        e.kernel_times()
        n = 20
        t0 = time.perf_counter()
        e.run_sweeps(100, n, half_len)
        wall = (time.perf_counter() - t0) / n * 1e3
        ms, k = e.kernel_times()
        imp = ms[0] / max(1, k[0])
        a, p = e.accept_counts()
        print(f"half_len={half_len} threads={threads} nblk={e.nblk}: impute {imp*1e3:.1f} us/sweep "
              f"(wall {wall*1e3:.1f}, other kernels {ms[1:].sum()*1e3/n:.1f}) -> {steps/imp/1e6:.1f} Gsteps/s, "
              f"{32*steps/imp/1e6:.0f} GB/s ({32*steps/imp/1e6/6556.2*100:.1f}% HBM); acc {a.sum()/max(1,p.sum()):.2f}",
              flush=True)
        e.close()
