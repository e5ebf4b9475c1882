"""Generate the golden input/output vectors under tests/golden/ from the CPU oracle.

The reference holds no golden vectors for the arithmetic (test/runtests.jl:56-58 is empty) and
cannot run here (no Julia), so these fixtures are ORACLE-generated ("parity unpinned" against
upstream; pinned against regressions of the oracle and used as fixed inputs for the CUDA path).
Run from the repo root:  python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import dmcmc_pkg  # noqa: E402

pkg = dmcmc_pkg.load()
from oracle import oracle as orc  # noqa: E402

P = pkg.problems
HERE = os.path.dirname(os.path.abspath(__file__))

CASES = {
    "fhn_c1": (lambda: P.fhn_spec(n_obs=5, C=1, dt=0.01), 0),              # config 1 shape, reduced
    "fhn_blocked": (lambda: P.fhn_spec(n_obs=6, C=3, dt=0.01), 1),        # chequerboard half_len=1
    "lv_blocks": (lambda: P.lv_spec(n_obs=8, C=2, dt=0.01), 2),           # config 2 shape, reduced
    "l63": (lambda: P.l63_spec(n_obs=4, C=2, dt=2e-3), 0),
}


def run_case(name):
    mk, half_len = CASES[name]
    spec = mk()
    o = orc.Oracle(spec)
    rng = np.random.default_rng(2024)
    scale = 0.2 if name.startswith("lv") else 1.0
    Z0 = scale * rng.standard_normal((o.C, o.S, o.m))
    o.init_paths(Z=Z0)
    out = {"Z0": Z0, "ll_init": o.ll.copy(), "half_len": half_len}
    for k in ("model", "d", "m", "theta", "N", "t", "rec_first", "obs_L", "obs_Sigma", "obs_v", "x0",
              "rho", "pin_eps", "obs_times", "t0"):
        out["spec_" + k] = np.asarray(spec[k])
    for it in range(3):
        if half_len:
            o.relayout(*P.chequerboard_layout(o.rec_nobs, half_len, it % 2))
            out[f"ll_relayout_{it}"] = o.ll.copy()
        Z = rng.standard_normal((o.C, o.S, o.m))
        E = rng.exponential(size=(o.C, o.nblk))
        llp, ll, acc = o.impute(Z, E)
        out[f"Z_{it}"], out[f"E_{it}"] = Z, E
        out[f"llp_{it}"], out[f"ll_{it}"], out[f"acc_{it}"] = llp, ll, acc
        out[f"Xacc_{it}"], out[f"Wacc_{it}"] = o.accepted_X(), o.accepted_W()
        out[f"Xprop_{it}"] = o.proposal_X()
    out["glued_chain0"] = o.glue_path(0)
    return out


if __name__ == "__main__":
    for name in CASES:
        np.savez_compressed(os.path.join(HERE, name + ".npz"), **run_case(name))
        print("wrote", name)
