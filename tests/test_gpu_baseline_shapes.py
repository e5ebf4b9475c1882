"""Parity at the REAL shapes of BASELINE.json's five configs: the CUDA path (through the C ABI, device backward
filter, in-register Philox noise -- nothing uploaded from the oracle) side by side with the CPU oracle on the same
seed.  Compared at RTOL = 1e-10 (north_star): per-sweep proposal and accepted log-likelihoods of every (chain, block),
the accept sequence (exact), the accepted paths, and for the parameter-inference configs the parameter step.  Every
comparison prints the achieved max-rel (helpers.check).

  C1  FitzHugh-Nagumo, 50 obs, dt = 1e-3 (100 steps per interval), 1 chain, unblocked, pCN rho = 0.96
      (docs/src/tutorials/simple_inference.md:77-105)                       -- warp-per-unit kernel
  C2  Lotka-Volterra, 200 obs, 1 chain, 8 interval blocks of 25 (+ the shifted pattern), random-walk
      parameter step                                                        -- warp-per-unit kernel
  C3  FitzHugh-Nagumo, 1024 chains x 100 obs, chequerboard half_len = 1, through dmcmc_run_sweeps:
      solve_kernel<ModelFHN, PSI, MODE_IMPUTE, PHILOX, REINV, FAST>, THE instantiation bench.py times
  C4  Lorenz-63, 10^4 obs (10 steps per interval), half_len = 5, a few chains, parameter step
  C5  Lorenz-96 d = 40, 64 obs x 100 steps (dt = 1e-4), 8 chains, half_len = 16, store_noise = 0
      (the full-size configuration's mode: W never stored)

Reference rows: update!(::PathImputation) src/run!_alterations.jl:5-26; init_ll! src/workspaces.jl:425-431;
set_proposal! src/run!_alterations.jl:177-211.
"""
import os

import numpy as np
import pytest

from helpers import check, check_scaled, interior_mask

pytestmark = pytest.mark.gpu
NCPU = os.cpu_count() or 1


def _engine(spec, **kw):
    from diffusionmcmc_jl_b200.engine import Engine
    return Engine(spec, **kw)


def _compare_tables(tag, o, e, psi):
    t = e.download_tables()
    keys = ("H", "F0", "c0") + (("Psi", "blk_g", "blk_Qc") if psi else ())
    for k in keys:
        check_scaled(f"{tag} device filter {k}", t[k], getattr(o, k))


def _owned_X(o):
    """accepted X of every interval without the carried start point: [C][S][d] (dmcmc_download_segment's shape)."""
    return o.accepted_X()[:, interior_mask(o.N)]


def _sweeps_side_by_side(tag, o, e, layouts, n, seed, it0=0):
    """n PathImputation updates alternating `layouts`: per-sweep API (set_layout + impute = fused layout switch) on the
    device, relayout + impute on the oracle."""
    n_acc = 0
    for it in range(n):
        lay = layouts[it % len(layouts)]
        if len(layouts) > 1 or it == 0:
            o.relayout(*lay)
            e.set_layout(*lay)
        llp_o, ll_o, acc_o = o.impute(seed=seed, it=it0 + it)
        llp_g, ll_g, acc_g = e.impute(it0 + it)
        check(f"{tag} sweep {it} ll deg", llp_g, llp_o)
        check(f"{tag} sweep {it} ll", ll_g, ll_o)
        assert np.array_equal(acc_g, acc_o), f"{tag} sweep {it}: accept sequence differs in {int((acc_g != acc_o).sum())} units"
        n_acc += int(acc_o.sum())
    check(f"{tag} accepted X after {n} sweeps", e.download_segment(0, e.J), _owned_X(o))
    return n_acc


def test_c1_fhn_single_chain_unblocked(pkg, orc):
    seed = 101
    spec = pkg.problems.fhn_spec(n_obs=50, C=1, dt=1e-3, rho=0.96)
    o, e = orc.Oracle(spec), _engine(spec, seed=seed)
    assert e.S == 5000 and e.nblk == 1
    _compare_tables("C1", o, e, psi=False)
    o.init_paths(seed=seed)
    e.init_paths()
    check("C1 init ll", e.get_ll(), o.ll)
    n = 8
    hist = (np.empty((n, 1, 1)), np.empty((n, 1, 1)), np.empty((n, 1, 1), np.uint8))
    e.run_sweeps(0, n, hist=hist)       # the production loop: stored noise, no layout switch
    ref = [o.impute(seed=seed, it=it) for it in range(n)]
    check("C1 run_sweeps ll deg history", hist[0], np.stack([r[0] for r in ref]))
    check("C1 run_sweeps ll history", hist[1], np.stack([r[1] for r in ref]))
    assert np.array_equal(hist[2].astype(bool), np.stack([r[2] for r in ref]))
    X, W = e.download_state(0)
    im = interior_mask(o.N)
    check("C1 accepted X", X[:, im], o.accepted_X()[:, im])
    check("C1 accepted W", W[:, im], o.accepted_W()[:, im])
    check("C1 glued path", e.download_path(0), o.glue_path(0))
    assert np.stack([r[2] for r in ref]).any()


def test_c2_lv_eight_blocks_and_parameter_step(pkg, orc):
    seed = 202
    spec = pkg.problems.lv_spec(n_obs=200, C=1, dt=0.01, rho=0.9)
    o, e = orc.Oracle(spec), _engine(spec, seed=seed)
    o.init_paths(seed=seed)
    e.init_paths()
    check("C2 init ll", e.get_ll(), o.ll)
    J = 200
    a1 = np.arange(0, J, 25, dtype=np.int32)                        # 8 interval blocks of 25
    lay_a = (a1, a1 + 24, (a1 + 24 < J - 1).astype(np.int32))
    b1 = np.concatenate([[0], np.arange(12, J, 25)]).astype(np.int32)  # the same blocks shifted by half a block
    b2 = np.concatenate([b1[1:] - 1, [J - 1]]).astype(np.int32)
    lay_b = (b1, b2, (b2 < J - 1).astype(np.int32))
    assert len(lay_a[0]) == 8 and len(lay_b[0]) == 9
    for lay in (lay_a, lay_b):
        o.set_layout(*lay)
        e.set_layout(*lay)
        e.backward_filter()
        _compare_tables(f"C2 {len(lay[0])} blocks", o, e, psi=True)
    n_acc = _sweeps_side_by_side("C2", o, e, [lay_a, lay_b], 6, seed)
    assert n_acc > 0
    # random-walk update of one drift parameter (critical: the aux law changes), fixed accepted noise
    o.relayout(*lay_a)
    e.relayout(*lay_a)
    check("C2 relayout ll", e.get_ll(), o.ll)
    for k, accept in enumerate((True, False)):
        th = np.array(o.theta, float)
        th[0] *= 1.0 + 0.01 * (k + 1)
        llp_o, ok_o = o.param_solve(th)
        llp_g, ok_g = e.param_loglik(th, critical=True)
        check(f"C2 parameter step {k} ll deg", llp_g, llp_o)
        assert np.array_equal(ok_g, ok_o)
        o.param_commit(accept)
        e.param_commit(accept)
        check(f"C2 parameter step {k} ll", e.get_ll(), o.ll)
    check("C2 accepted X after the parameter steps", e.download_segment(0, e.J), _owned_X(o))


def test_c3_fhn_1024_chains_blocked_run_sweeps_full_size(pkg, orc):
    """The bench workload, the bench call (dmcmc_run_sweeps, chequerboard half_len = 1, Philox, fused layout switch),
    the bench kernel -- against the oracle at full size."""
    seed = 303
    P = pkg.problems
    spec = P.fhn_spec(n_obs=100, C=1024, dt=1e-3, rho=0.96)
    o, e = orc.Oracle(spec, nthreads=NCPU), _engine(spec, seed=seed)
    assert e.C * e.S == 10_240_000
    o.init_paths(seed=seed)
    e.init_paths()
    check("C3 init ll", e.get_ll(), o.ll)
    lays = [P.chequerboard_layout(o.rec_nobs, 1, p) for p in (0, 1)]
    nb = [len(l[0]) for l in lays]
    assert nb == [50, 51] and e.C * nb[0] == 51_200      # units per sweep: far above the warp-per-unit kernel's range
    n, stride = 4, max(nb)
    hist = (np.empty((n, e.C, stride)), np.empty((n, e.C, stride)), np.empty((n, e.C, stride), np.uint8))
    rho = np.full((n, e.J), 0.96)
    e.run_sweeps(0, n, half_len=1, first_pattern=0, rho_sched=rho, hist=hist)
    n_acc = 0
    for it in range(n):
        o.relayout(*lays[it % 2])
        llp_o, ll_o, acc_o = o.impute(seed=seed, it=it)
        k = nb[it % 2]
        check(f"C3 sweep {it} ll deg", hist[0][it, :, :k], llp_o)
        check(f"C3 sweep {it} ll", hist[1][it, :, :k], ll_o)
        acc_g = hist[2][it, :, :k].astype(bool)
        assert np.array_equal(acc_g, acc_o), f"C3 sweep {it}: accept sequence differs in {int((acc_g != acc_o).sum())} of {acc_o.size} units"
        n_acc += int(acc_o.sum())
    assert 0.02 < n_acc / (n * e.C * 50.5) < 0.98
    check("C3 accepted X after 4 sweeps", e.download_segment(0, e.J), _owned_X(o))
    for pat in (0, 1):   # the tables the sweeps ran on
        o.set_layout(*lays[pat])
        e.set_layout(*lays[pat])
        _compare_tables(f"C3 pattern {pat}", o, e, psi=True)


def test_c4_l63_ten_thousand_obs_blocked(pkg, orc):
    seed = 404
    P = pkg.problems
    spec = P.l63_spec(n_obs=10_000, C=3, dt=1e-3, rho=0.9)
    o, e = orc.Oracle(spec, nthreads=NCPU), _engine(spec, seed=seed)
    assert e.J == 10_000 and e.S == 100_000
    o.init_paths(seed=seed)
    e.init_paths()
    check("C4 init ll", e.get_ll(), o.ll)
    lays = [P.chequerboard_layout(o.rec_nobs, 5, p) for p in (0, 1)]
    assert [len(l[0]) for l in lays] == [1000, 1001]
    n_acc = _sweeps_side_by_side("C4", o, e, lays, 4, seed)
    assert n_acc > 0
    # parameter step under the blocked layout (every block's ll deg; the sum is what the all-reduce carries)
    o.relayout(*lays[0])
    e.relayout(*lays[0])
    check("C4 relayout ll", e.get_ll(), o.ll)
    th = np.array(o.theta, float)
    th[1] += 0.05
    llp_o, ok_o = o.param_solve(th)
    llp_g, ok_g = e.param_loglik(th, critical=True)
    check("C4 parameter step ll deg (per block)", llp_g, llp_o)
    check("C4 parameter step ll deg (sum over blocks)", llp_g.sum(axis=1), llp_o.sum(axis=1))
    assert np.array_equal(ok_g, ok_o)
    o.param_commit(True)
    e.param_commit(True)
    check("C4 accepted X after the parameter step", e.download_segment(0, e.J), _owned_X(o))


def test_c5_l96_d40_blocked(pkg, orc):
    seed = 505
    P = pkg.problems
    spec = P.l96_spec(n_obs=64, C=8, d=40, dt=1e-4, rho=0.9)
    o, e = orc.Oracle(spec, nthreads=NCPU), _engine(spec, seed=seed, store_noise=False)
    assert e.d == 40 and int(e.N[0]) == 100 and e.S == 6400
    o.init_paths(seed=seed)
    e.init_paths()
    check("C5 init ll", e.get_ll(), o.ll)
    lays = [P.chequerboard_layout(o.rec_nobs, 16, p) for p in (0, 1)]
    assert [len(l[0]) for l in lays] == [2, 3]
    n_acc = _sweeps_side_by_side("C5", o, e, lays, 3, seed)
    assert n_acc > 0
