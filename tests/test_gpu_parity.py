"""Parity tests proper: the CUDA path (through the C ABI) against the CPU oracle on the same
seeded inputs.  Tolerance: 1e-10 relative (north_star) on imputed paths, Wiener paths and
per-block log-likelihoods; accept decisions identical except ties within that tolerance."""
import numpy as np
import pytest

from helpers import RTOL, accept_mismatch_is_tie, check, check_scaled, close, interior_mask, maxrel

pytestmark = pytest.mark.gpu


@pytest.fixture(autouse=True, params=["thread_per_unit", "warp_per_unit"])
def imputation_kernel(request, monkeypatch):
    """Every parity test runs against both imputation kernels (dmcmc_set_cooperative): the problems here are
    small enough that the automatic choice would always be the warp-per-unit kernel."""
    monkeypatch.setenv("DMCMC_COOP", "0" if request.param == "thread_per_unit" else "1")
    return request.param


def _mk(pkg, name, **kw):
    P = pkg.problems
    return {"fhn": P.fhn_spec, "lv": P.lv_spec, "l63": P.l63_spec, "lin2": P.lin2_spec,
            "lvm": lambda **k: P.lv_spec(multiplicative=True, **k)}[name](**kw)


def _pair(pkg, orc, spec, upload=True):
    from diffusionmcmc_jl_b200.engine import Engine
    o = orc.Oracle(spec)
    e = Engine(spec)
    if upload:
        e.upload_tables(o.H, o.F0, None, o.c0)
    return o, e


def _compare_state(o, e, which, what=""):
    Xg, Wg = e.download_state(which)
    Xo = o.accepted_X() if which == 0 else o.proposal_X()
    Wo = o.accepted_W() if which == 0 else o.proposal_W()
    im = interior_mask(o.N)
    check(f"{what} X[{which}]", Xg[:, im], Xo[:, im])
    check(f"{what} W[{which}]", Wg[:, im], Wo[:, im])


CASES = [
    ("fhn", dict(n_obs=6, C=37, dt=0.01)),      # ragged chain count, N=10 -> partial tiles
    ("fhn", dict(n_obs=3, C=5, dt=0.004)),      # N=25: 3 full tiles + 1 step
    ("lv", dict(n_obs=5, C=33, dt=0.01)),
    ("l63", dict(n_obs=7, C=9, dt=1e-3)),
    ("lin2", dict(n_obs=4, C=3, dt=0.01)),
    ("lvm", dict(n_obs=5, C=8, dt=0.005)),
]


@pytest.mark.parametrize("name,kw", CASES)
def test_impute_parity_explicit_noise(pkg, orc, name, kw):
    spec = _mk(pkg, name, **kw)
    o, e = _pair(pkg, orc, spec)
    rng = np.random.default_rng(123)
    Z0 = rng.standard_normal((o.C, o.S, o.m)) * (0.2 if name in ("lv", "lvm") else 1.0)
    o.init_paths(Z=Z0)
    e.init_paths(Z=Z0)
    check(f"{name} init ll", e.get_ll(), o.ll)
    _compare_state(o, e, 0, "init")
    n_acc = 0
    for it in range(4):
        Z = rng.standard_normal((o.C, o.S, o.m))
        E = rng.exponential(size=(o.C, o.nblk))
        ll_before = o.ll.copy()
        llp_o, ll_o, acc_o = o.impute(Z, E)
        llp_g, ll_g, acc_g = e.impute(it, Z, E)
        check(f"{name} it {it} ll deg", llp_g, llp_o)
        assert accept_mismatch_is_tie(acc_g, acc_o, E, llp_o, ll_before)
        assert np.array_equal(acc_g, acc_o)
        check(f"{name} it {it} ll", ll_g, ll_o)
        _compare_state(o, e, 0, f"it {it}")
        _compare_state(o, e, 1, f"it {it}")
        n_acc += acc_o.sum()
    assert n_acc > 0


def test_lv_bound_violation_rejected(pkg, orc):
    """A path that leaves the positive quadrant reports ll = -Inf and is rejected
    (src/run!_alterations.jl:18: `(success, ll)`; SURVEY section 5 failure handling)."""
    spec = _mk(pkg, "lv", n_obs=4, C=6, dt=0.01)
    o, e = _pair(pkg, orc, spec)
    rng = np.random.default_rng(5)
    Z0 = 0.1 * rng.standard_normal((o.C, o.S, o.m))
    o.init_paths(Z=Z0)
    e.init_paths(Z=Z0)
    Z = rng.standard_normal((o.C, o.S, o.m))
    Z[::2] *= 400.0  # huge noise on even chains: certain to cross zero
    o.rho[:] = 0.0
    e.set_rho(o.rho)
    E = rng.exponential(size=(o.C, o.nblk))
    llp_o, ll_o, acc_o = o.impute(Z, E)
    llp_g, ll_g, acc_g = e.impute(0, Z, E)
    assert np.all(np.isneginf(llp_o[::2])) and np.all(np.isneginf(llp_g[::2]))
    assert not acc_g[::2].any()
    assert np.array_equal(acc_g, acc_o)
    check("lv bound violation ll", ll_g, ll_o)
    _compare_state(o, e, 0)


@pytest.mark.parametrize("name,kw", [("fhn", dict(n_obs=5, C=40, dt=0.01)),
                                     ("l63", dict(n_obs=4, C=3, dt=2e-3))])
def test_impute_parity_philox(pkg, orc, name, kw):
    """Production noise: Philox4x32-10 + Box-Muller in-register vs the oracle's restatement."""
    spec = _mk(pkg, name, **kw)
    o, e = _pair(pkg, orc, spec)
    from diffusionmcmc_jl_b200.engine import Engine
    e.close()
    e = Engine(spec, seed=0xC0FFEE1234)
    e.upload_tables(o.H, o.F0, None, o.c0)
    o.init_paths(seed=0xC0FFEE1234)
    e.init_paths()
    check(f"{name} philox init ll", e.get_ll(), o.ll)
    for it in range(3):
        llp_o, ll_o, acc_o = o.impute(seed=0xC0FFEE1234, it=it)
        llp_g, ll_g, acc_g = e.impute(it)
        check(f"{name} philox it {it} ll deg", llp_g, llp_o)
        assert np.array_equal(acc_g, acc_o)
        _compare_state(o, e, 0, f"philox it {it}")


def test_device_backward_filter_vs_oracle(pkg, orc):
    """a8: device recursion vs oracle recursion (same scheme, FMA contraction differs)."""
    for name, kw in [("fhn", dict(n_obs=8, C=2, dt=0.01)), ("l63", dict(n_obs=5, C=1, dt=1e-3)),
                     ("lv", dict(n_obs=6, C=1, dt=0.01))]:
        spec = _mk(pkg, name, **kw)
        o, e = _pair(pkg, orc, spec, upload=False)
        t = e.download_tables()
        for key, ref in (("H", o.H), ("F0", o.F0), ("c0", o.c0)):
            check_scaled(f"device filter {name} {key}", t[key], ref)


def test_device_backward_filter_pinned_layout_vs_oracle(pkg, orc):
    """a8 under chequerboard blocking: the device filter (parallel in time inside an interval: every grid point from
    the interval's right-end value and the transition over tau_k) against the oracle's step-by-step recursion,
    including the pinned blocks' Psi, g, Qc.  The oracle runs the same scheme as the device (every grid point from the
    interval's right end, orc_backward_filter_scheme 1), so what is left is FMA contraction."""
    P = pkg.problems
    for name, kw, half in [("fhn", dict(n_obs=9, C=2, dt=0.004), 1), ("l63", dict(n_obs=12, C=1, dt=1e-3), 2),
                           ("lv", dict(n_obs=10, C=1, dt=0.01), 3),
                           ("fhn", dict(n_obs=3, C=1, dt=4e-4), 1)]:   # 250 grid points per interval: more than one per thread
        spec = _mk(pkg, name, **kw)
        o, e = _pair(pkg, orc, spec, upload=False)
        for pat in (0, 1):
            lay = P.chequerboard_layout(o.rec_nobs, half, pat)
            o.set_layout(*lay) if hasattr(o, "set_layout") else o.relayout(*lay)
            e.set_layout(*lay)
            e.backward_filter(0)
            t = e.download_tables()
            for key, ref in (("H", o.H), ("F0", o.F0), ("Psi", o.Psi), ("c0", o.c0), ("blk_g", o.blk_g), ("blk_Qc", o.blk_Qc)):
                check_scaled(f"device filter pinned {name} h={half} pattern {pat} {key}", t[key], ref)


def test_blocked_sweeps_parity(pkg, orc):
    """Chequerboard blocking: relayout (W re-inversion + ll recompute) then imputation, both
    patterns, pinned end points."""
    spec = _mk(pkg, "fhn", n_obs=9, C=35, dt=0.01)
    o, e = _pair(pkg, orc, spec)
    P = pkg.problems
    rng = np.random.default_rng(9)
    Z0 = rng.standard_normal((o.C, o.S, o.m))
    o.init_paths(Z=Z0)
    e.init_paths(Z=Z0)
    for it in range(6):
        lay = P.chequerboard_layout(o.rec_nobs, 2, it % 2)
        o.relayout(*lay)
        e.relayout(*lay)
        e.upload_tables(o.H, o.F0, o.Psi, o.c0, o.blk_g, o.blk_Qc)
        e.relayout(*lay)  # recompute with the oracle's tables
        check(f"blocked it {it} relayout ll", e.get_ll(), o.ll)
        Xg, Wg = e.download_state(0)
        im = interior_mask(o.N)
        check(f"blocked it {it} re-inverted W", Wg[:, im], o.accepted_W()[:, im])
        e.set_ll(o.ll)
        Z = rng.standard_normal((o.C, o.S, o.m))
        E = rng.exponential(size=(o.C, o.nblk))
        ll_before = o.ll.copy()
        llp_o, ll_o, acc_o = o.impute(Z, E)
        llp_g, ll_g, acc_g = e.impute(it, Z, E)
        check(f"blocked it {it} ll deg", llp_g, llp_o)
        assert accept_mismatch_is_tie(acc_g, acc_o, E, llp_o, ll_before)
        assert np.array_equal(acc_g, acc_o)
        Xg, _ = e.download_state(0)
        check(f"blocked it {it} X", Xg[:, im], o.accepted_X()[:, im])
        # glued path (PathSavingBuffer format) is continuous and equals the oracle's
        check(f"blocked it {it} glued path", e.download_path(3), o.glue_path(3))


def test_param_step_parity(pkg, orc):
    """a9: set_proposal! with fixed W under theta deg, then commit / reject."""
    spec = _mk(pkg, "fhn", n_obs=6, C=4, dt=0.01)
    o, e = _pair(pkg, orc, spec)
    rng = np.random.default_rng(11)
    Z0 = rng.standard_normal((o.C, o.S, o.m))
    o.init_paths(Z=Z0)
    e.init_paths(Z=Z0)
    im = interior_mask(o.N)
    for k, accept in enumerate([True, False, True]):
        th = spec["theta"].copy()
        th[2] += 0.05 * (k + 1)   # gamma: critical (enters the aux law)
        th[1] -= 0.02 * (k + 1)
        llp_o, ok_o = o.param_solve(th)
        llp_g, ok_g = e.param_loglik(th, critical=True)
        check(f"param step {k} ll deg", llp_g, llp_o)
        assert np.array_equal(ok_g, ok_o)
        Xg, _ = e.download_state(1)
        check(f"param step {k} X deg", Xg[:, im], o.proposal_X()[:, im])
        o.param_commit(accept)
        e.param_commit(accept)
        check(f"param step {k} ll after commit", e.get_ll(), o.ll)
        Xg, Wg = e.download_state(0)
        check(f"param step {k} X", Xg[:, im], o.accepted_X()[:, im])
        check(f"param step {k} W", Wg[:, im], o.accepted_W()[:, im])
        # an imputation sweep after the parameter step still agrees
        Z = rng.standard_normal((o.C, o.S, o.m))
        E = rng.exponential(size=(o.C, o.nblk))
        llp_o2, _, acc_o = o.impute(Z, E)
        llp_g2, _, acc_g = e.impute(k, Z, E)
        check(f"param step {k} next imputation ll deg", llp_g2, llp_o2)
        assert np.array_equal(acc_g, acc_o)


def test_rho_one_is_identity_full_size(pkg):
    """Size-independent property at BASELINE size (FHN, 1024 chains x 100 obs): rho = 1 =>
    the proposal reproduces the accepted path bit for bit and ll deg == ll (SURVEY section 4 iv)."""
    from diffusionmcmc_jl_b200.engine import Engine
    spec = pkg.problems.fhn_spec(n_obs=100, C=1024, dt=1e-3)
    e = Engine(spec, seed=7)
    e.init_paths()
    ll0 = e.get_ll()
    e.set_rho(np.ones(e.J))
    llp, ll, acc = e.impute(1)
    assert np.array_equal(llp, ll0) and acc.all()
    e.set_rho(np.full(e.J, 0.96))
    _, _, acc = e.impute(2)
    p0 = e.download_path(17)
    assert np.isfinite(p0).all() and p0.shape == (100 * 100 + 1, 2)
    assert 0.02 < acc.mean() < 0.98


def test_chain_sharding_is_invariant(pkg):
    """Partition invariance: chains [0,24) on one context equal chains
    [0,12) + [12,24) on two contexts bit for bit (Philox streams are per global chain)."""
    from diffusionmcmc_jl_b200.engine import Engine
    spec = pkg.problems.fhn_spec(n_obs=6, C=24, dt=0.01)
    whole = Engine(spec, seed=99)
    parts = [Engine(spec, seed=99, chain_offset=o, n_chains=12) for o in (0, 12)]
    for en in [whole] + parts:
        en.init_paths()
    for it in range(3):
        a = whole.impute(it)
        b = [p.impute(it) for p in parts]
        for k in range(3):
            assert np.array_equal(a[k], np.concatenate([b[0][k], b[1][k]]))
    Xw, Ww = whole.download_state(0)
    Xp = np.concatenate([p.download_state(0)[0] for p in parts])
    assert np.array_equal(Xw, Xp)


@pytest.mark.parametrize("name,kw,h", [("fhn", dict(n_obs=9, C=70, dt=0.01), 1),
                                       ("fhn", dict(n_obs=7, C=3, dt=0.004), 2),
                                       ("lv", dict(n_obs=8, C=5, dt=0.01), 2),
                                       ("l63", dict(n_obs=6, C=2, dt=2e-3), 1)])
def test_fused_layout_switch_parity(pkg, orc, name, kw, h):
    """Production blocked sweep: the layout switch is fused into the imputation kernel (accepted
    noise and ll are recovered from the accepted X on the fly, no W stored) and must equal the
    oracle's relayout followed by its imputation."""
    spec = _mk(pkg, name, **kw)
    o, e = _pair(pkg, orc, spec)
    P = pkg.problems
    rng = np.random.default_rng(21)
    Z0 = rng.standard_normal((o.C, o.S, o.m)) * (0.2 if name == "lv" else 1.0)
    o.init_paths(Z=Z0)
    e.init_paths(Z=Z0)
    im = interior_mask(o.N)
    for it in range(5):
        lay = P.chequerboard_layout(o.rec_nobs, h, it % 2)
        o.relayout(*lay)
        e.set_layout(*lay)  # no relayout kernel: the next impute re-inverts on the fly
        e.upload_tables(o.H, o.F0, o.Psi, o.c0, o.blk_g, o.blk_Qc)
        Z = rng.standard_normal((o.C, o.S, o.m))
        E = rng.exponential(size=(o.C, o.nblk))
        ll_before = o.ll.copy()
        llp_o, ll_o, acc_o = o.impute(Z, E)
        llp_g, ll_g, acc_g = e.impute(it, Z, E)
        check(f"fused switch {name} h={h} it {it} ll deg", llp_g, llp_o)
        assert accept_mismatch_is_tie(acc_g, acc_o, E, llp_o, ll_before)
        assert np.array_equal(acc_g, acc_o)
        check(f"fused switch {name} h={h} it {it} ll", ll_g, ll_o)
        Xg, _ = e.download_state(0)
        check(f"fused switch {name} h={h} it {it} X", Xg[:, im], o.accepted_X()[:, im])


def test_run_sweeps_histories_and_counts(pkg):
    """dmcmc_run_sweeps: chequerboard production loop; histories land with the host stride and
    agree with the device-side accept counters; ll history is consistent with the accepts."""
    from diffusionmcmc_jl_b200.engine import Engine
    spec = pkg.problems.fhn_spec(n_obs=10, C=96, dt=0.01)
    e = Engine(spec, seed=5)
    e.init_paths()
    n, stride = 8, 6
    hist = (e.pinned((n, e.C, stride), np.float64), e.pinned((n, e.C, stride), np.float64),
            e.pinned((n, e.C, stride), np.uint8))
    rho = e.pinned((n, e.J), np.float64)
    rho[...] = 0.9
    e.accept_counts()
    e.run_sweeps(0, n, half_len=1, first_pattern=0, rho_sched=rho, hist=hist)
    llp, ll, acc = hist
    nb = [5, 6]  # 10 obs, half_len 1: pattern 0 -> 5 blocks, pattern 1 -> 6 blocks
    for it in range(n):
        k = nb[it % 2]
        a = acc[it, :, :k].astype(bool)
        assert np.all(np.isfinite(ll[it, :, :k]))
        assert np.array_equal(ll[it, :, :k][a], llp[it, :, :k][a])
    # a second engine stepping sweep by sweep through the blocking API gives the same thing
    e2 = Engine(spec, seed=5)
    e2.init_paths()
    P = pkg.problems
    for it in range(n):
        e2.set_rho(np.full(e2.J, 0.9))
        e2.set_layout(*P.chequerboard_layout(e2.rec_nobs, 1, it % 2))
        e2.backward_filter()
        l2p, l2, a2 = e2.impute(it)
        k = nb[it % 2]
        assert np.array_equal(a2, acc[it, :, :k].astype(bool))
        assert np.array_equal(l2p, llp[it, :, :k])
    X1, _ = e.download_state(0)
    X2, _ = e2.download_state(0)
    assert np.array_equal(X1, X2)


def _ragged_fhn(pkg, C):
    """Irregular observation times => ragged grids: N_j in {1, 3, 17, 16, 33, 8, 40, 5} (single
    steps, odd counts, exact / one-over multiples of the 8-step W tile and the 16-step staged tile)."""
    P = pkg.problems
    rng = np.random.default_rng(77)
    steps = np.array([1, 3, 17, 16, 33, 8, 40, 5])
    dt = 0.004
    times = np.cumsum(steps * dt)
    x0 = np.array([-1.0, -1.0])
    tt, X = P.simulate(P.MODEL_FHN, P.FHN_THETA, x0, times[-1] + 1e-3, 1e-4, rng)
    idx = np.round(times / 1e-4).astype(int)
    v = X[idx, 0] + 0.1 * rng.standard_normal(len(steps))
    spec = P.make_spec(P.MODEL_FHN, P.FHN_THETA, x0, times, v[:, None], [[1.0, 0.0]], [[0.01]], dt, C=C, rho=0.9)
    assert list(spec["N"]) == list(steps)
    return spec


def test_ragged_grid_parity(pkg, orc):
    """Ragged inputs (edge cases of the record layout): stored-noise sweeps, fused layout switches in
    both chequerboard patterns, explicit noise and Philox noise, against the oracle."""
    from diffusionmcmc_jl_b200.engine import Engine
    spec = _ragged_fhn(pkg, C=45)
    o, e = _pair(pkg, orc, spec)
    P = pkg.problems
    rng = np.random.default_rng(31)
    Z0 = rng.standard_normal((o.C, o.S, o.m))
    o.init_paths(Z=Z0)
    e.init_paths(Z=Z0)
    check("ragged init ll", e.get_ll(), o.ll)
    _compare_state(o, e, 0, "init")
    for it in range(2):  # unblocked, stored noise
        Z = rng.standard_normal((o.C, o.S, o.m))
        E = rng.exponential(size=(o.C, o.nblk))
        llp_o, ll_o, acc_o = o.impute(Z, E)
        llp_g, ll_g, acc_g = e.impute(it, Z, E)
        check(f"ragged unblocked {it} ll deg", llp_g, llp_o)
        assert np.array_equal(acc_g, acc_o)
        _compare_state(o, e, 0, f"unblocked {it}")
        _compare_state(o, e, 1, f"unblocked {it}")
    im = interior_mask(o.N)
    for it in range(4):  # fused layout switch, both patterns
        lay = P.chequerboard_layout(o.rec_nobs, 1 + it // 2, it % 2)
        o.relayout(*lay)
        e.set_layout(*lay)
        e.upload_tables(o.H, o.F0, o.Psi, o.c0, o.blk_g, o.blk_Qc)
        Z = rng.standard_normal((o.C, o.S, o.m))
        E = rng.exponential(size=(o.C, o.nblk))
        ll_before = o.ll.copy()
        llp_o, ll_o, acc_o = o.impute(Z, E)
        llp_g, ll_g, acc_g = e.impute(10 + it, Z, E)
        check(f"ragged blocked {it} ll deg", llp_g, llp_o)
        assert accept_mismatch_is_tie(acc_g, acc_o, E, llp_o, ll_before)
        assert np.array_equal(acc_g, acc_o)
        Xg, _ = e.download_state(0)
        check(f"ragged blocked {it} X", Xg[:, im], o.accepted_X()[:, im])
        check(f"ragged blocked {it} glued path", e.download_path(7), o.glue_path(7))
    # Philox noise on the ragged grid
    o2 = orc.Oracle(spec)
    e2 = Engine(spec, seed=4242)
    e2.upload_tables(o2.H, o2.F0, None, o2.c0)
    o2.init_paths(seed=4242)
    e2.init_paths()
    for it in range(3):
        llp_o, ll_o, acc_o = o2.impute(seed=4242, it=it)
        llp_g, ll_g, acc_g = e2.impute(it)
        check(f"ragged philox {it} ll deg", llp_g, llp_o)
        assert np.array_equal(acc_g, acc_o)
        _compare_state(o2, e2, 0, f"philox {it}")


def test_block_ll_additivity_full_size(pkg):
    """Size-independent property at BASELINE size: after a sweep the per-block log-likelihoods of the
    blocked layout are finite for every (chain, block), the glued path is continuous across block
    edges (pinned end points), and an unblocked re-evaluation of the same accepted path is
    reproducible bit for bit (init_ll! twice)."""
    from diffusionmcmc_jl_b200.engine import Engine
    spec = pkg.problems.fhn_spec(n_obs=100, C=1024, dt=1e-3)
    e = Engine(spec, seed=11)
    e.init_paths()
    e.run_sweeps(0, 4, half_len=1)
    ll = e.get_ll()
    assert ll.shape[0] == 1024 and np.isfinite(ll).all()
    p = e.download_path(1023)
    assert np.isfinite(p).all()
    # consecutive points stay within a few Euler steps' worth of motion: no jump at block edges
    assert np.abs(np.diff(p, axis=0)).max() < 0.5
    P = pkg.problems
    e.relayout(*P.chequerboard_layout(e.rec_nobs, 0, 0))
    ll1 = e.get_ll().copy()
    e.init_ll()
    assert np.array_equal(ll1, e.get_ll())


@pytest.mark.parametrize("name,kw,half", [("fhn", dict(n_obs=9, C=70, dt=0.004), 0), ("fhn", dict(n_obs=9, C=70, dt=0.004), 1),
                                          ("lv", dict(n_obs=12, C=5, dt=0.01), 2), ("l63", dict(n_obs=40, C=3, dt=1e-3), 3),
                                          ("l63", dict(n_obs=10, C=2, dt=2e-4), 0), ("lvm", dict(n_obs=6, C=4, dt=0.005), 1)])
def test_warp_per_unit_kernel_is_bit_identical(pkg, monkeypatch, name, kw, half):
    """The kernel choice (by unit count) must never change a sample: stored-noise sweeps (half = 0) and fused
    layout-switch sweeps, Philox noise, ragged tiles (N = 25, 50, 10, 5 steps per interval)."""
    from diffusionmcmc_jl_b200.engine import Engine
    spec = _mk(pkg, name, **kw)
    out = []
    for mode in (0, 1):
        monkeypatch.setenv("DMCMC_COOP", str(mode))
        e = Engine(spec, seed=11)
        e.init_paths()
        e.run_sweeps(0, 5, half)
        X, W = e.download_state(0)
        th = np.array(spec["theta"], float)
        llp, ok = e.param_loglik(th * (1 + 1e-3), critical=True)   # a9 under a perturbed law, fixed noise
        Xp, _ = e.download_state(1)
        e.param_commit(False)
        e.relayout(*pkg.problems.chequerboard_layout(e.rec_nobs, max(half, 1), 1))   # layout switch: W re-inversion + ll
        _, Wr = e.download_state(0)
        out.append((X, W, e.get_ll().copy(), e.accept_counts()[0].copy(), llp, ok, Xp, Wr))
        e.close()
    for a, b, what in zip(out[0], out[1], ("X", "W", "ll", "accept counts", "parameter-step ll", "parameter-step success",
                                           "parameter-step X", "W after a layout switch")):
        assert np.array_equal(a, b), f"{what} differs between the two imputation kernels"



@pytest.mark.parametrize("name,kw,half", [("l63", dict(n_obs=40, C=3, dt=1e-3), 3), ("lv", dict(n_obs=24, C=2, dt=0.01), 2),
                                          ("fhn", dict(n_obs=30, C=2, dt=0.01), 5), ("lvm", dict(n_obs=12, C=4, dt=0.01), 1)])
def test_tiles_of_several_short_intervals_are_bit_identical(pkg, monkeypatch, name, kw, half):
    """Warp-per-unit kernel: a tile may hold up to four WHOLE short intervals (lanes of one tile then work on different
    intervals; phase S walks them one after another).  Forced on and off (DMCMC_COOP_MULTI) it must never change a
    sample: fused layout-switch sweeps, parameter-step solve, layout switch -- 10 Euler steps per interval, so three
    intervals share a tile, blocks of 2-10 intervals."""
    from diffusionmcmc_jl_b200.engine import Engine
    spec = _mk(pkg, name, **kw)
    assert max(spec["N"]) <= 16
    monkeypatch.setenv("DMCMC_COOP", "1")
    out = []
    for multi in ("0", "1"):
        monkeypatch.setenv("DMCMC_COOP_MULTI", multi)
        e = Engine(spec, seed=21)
        e.init_paths()
        e.run_sweeps(0, 4, half)
        X, W = e.download_state(0)
        th = np.array(spec["theta"], float)
        llp, ok = e.param_loglik(th * (1 + 1e-3), critical=True)
        Xp, _ = e.download_state(1)
        e.param_commit(False)
        e.relayout(*pkg.problems.chequerboard_layout(e.rec_nobs, max(half - 1, 1), 1))
        _, Wr = e.download_state(0)
        out.append((X, W, e.get_ll().copy(), e.accept_counts()[0].copy(), llp, ok, Xp, Wr))
        e.close()
    for a, b, what in zip(out[0], out[1], ("X", "W", "ll", "accept counts", "parameter-step ll", "parameter-step success",
                                           "parameter-step X", "W after a layout switch")):
        assert np.array_equal(a, b), f"{what} differs between one interval per tile and several"
