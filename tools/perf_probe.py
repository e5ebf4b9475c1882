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
spec = P.fhn_spec(n_obs=nobs, C=C, dt=1e-3)
steps = C * nobs * 100
for half_len in [0, 1, 2, 5]:
    for threads in ([32, 64, 128] if half_len else [32]):
        e = Engine(spec, seed=1)
        e.lib.dmcmc_set_threads(e.h, threads) if hasattr(e.lib, "dmcmc_set_threads") else None
        e.init_paths()
        lays = [P.chequerboard_layout(e.rec_nobs, half_len, p) for p in (0, 1)]
        e.relayout(*lays[0])
        e.kernel_ms()
        for it in range(3):
            e.impute(it, want=False)
        e.kernel_ms()
        n = 10
        t0 = time.perf_counter()
        e.run_imputation(100, n, histories=False)
        wall = (time.perf_counter() - t0) / n * 1e3
        ms, k = e.kernel_ms()
        ms /= k
        # relayout cost
        e.relayout(*lays[1]); e.kernel_ms()
        for i in range(4):
            e.relayout(*lays[i % 2])
        rms, rk = e.kernel_ms(); rms /= rk
        a, p = e.accept_counts()
        print(f"half_len={half_len} threads={threads} nblk={e.nblk}: impute {ms*1e3:.1f} us/sweep "
              f"(wall {wall*1e3:.1f}) -> {steps/ms/1e6:.1f} Gsteps/s, {32*steps/ms/1e6:.0f} GB/s "
              f"({32*steps/ms/1e6/6556.2*100:.1f}% HBM); relayout {rms*1e3:.1f} us; acc {a.sum()/max(1,p.sum()):.2f}",
              flush=True)
        e.close()
