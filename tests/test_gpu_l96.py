"""Lorenz-96 (d = 40, BASELINE config 5 at reduced size): the FP64 tensor-core solve kernel (eight chains per warp,
dmcmc_solve_l96.cu) and the CTA-cooperative backward filter against the CPU oracle."""
import numpy as np
import pytest

from helpers import accept_mismatch_is_tie, check, check_scaled, close, interior_mask, maxrel

pytestmark = pytest.mark.gpu


@pytest.fixture(autouse=True, params=["auto_cta", "8_warps_per_cta"])
def l96_mapping(request, monkeypatch):
    """Every Lorenz-96 test runs with the launcher's own CTA size (one or two warps for these chain counts) and with
    eight warps per CTA (BASELINE C5's shape: the warps of a CTA share the table stages behind full / empty mbarriers,
    warp 0 refills a stage once every warp has released it; surplus warps shadow the last chain)."""
    if request.param == "8_warps_per_cta":
        monkeypatch.setenv("DMCMC_L96_WARPS", "8")
    else:
        monkeypatch.delenv("DMCMC_L96_WARPS", raising=False)
    return request.param


def _pair(pkg, orc, spec, upload=True, **kw):
    from diffusionmcmc_jl_b200.engine import Engine
    o = orc.Oracle(spec)
    e = Engine(spec, **kw)
    if upload:
        e.upload_tables(o.H, o.F0, None, o.c0)
    return o, e


def _state_ok(o, e, tol, what):
    Xg, Wg = e.download_state(0)
    im = interior_mask(o.N)
    check(f"l96 {what} X", Xg[:, im], o.accepted_X()[:, im], tol)
    check(f"l96 {what} W", Wg[:, im], o.accepted_W()[:, im], tol)


def test_l96_impute_parity_explicit_noise(pkg, orc):
    spec = pkg.problems.l96_spec(n_obs=3, C=5, dt=1e-3, rho=0.9)   # N = 10 steps per interval
    o, e = _pair(pkg, orc, spec)
    rng = np.random.default_rng(40)
    Z0 = rng.standard_normal((o.C, o.S, o.m))
    o.init_paths(Z=Z0)
    e.init_paths(Z=Z0)
    check("l96 init ll", e.get_ll(), o.ll)
    _state_ok(o, e, 1e-10, "init")
    for it in range(3):
        Z = rng.standard_normal((o.C, o.S, o.m))
        E = rng.exponential(size=(o.C, o.nblk))
        ll_before = o.ll.copy()
        llp_o, ll_o, acc_o = o.impute(Z, E)
        llp_g, ll_g, acc_g = e.impute(it, Z, E)
        check(f"l96 it {it} ll deg", llp_g, llp_o)
        assert accept_mismatch_is_tie(acc_g, acc_o, E, llp_o, ll_before)
        assert np.array_equal(acc_g, acc_o)
        _state_ok(o, e, 1e-10, f"it {it}")


def test_l96_device_backward_filter(pkg, orc):
    spec = pkg.problems.l96_spec(n_obs=4, C=1, dt=2e-3)
    o, e = _pair(pkg, orc, spec, upload=False)
    t = e.download_tables()
    for k, ref in (("H", o.H), ("F0", o.F0), ("c0", o.c0)):
        check_scaled(f"l96 device filter {k}", t[k], ref)
    # pinned layout: Psi, g, Qc
    lay = pkg.problems.chequerboard_layout(o.rec_nobs, 1, 0)
    o.set_layout(*lay)
    e.set_layout(*lay)
    e.backward_filter()
    t = e.download_tables()
    for k, ref in (("H", o.H), ("F0", o.F0), ("Psi", o.Psi), ("blk_g", o.blk_g), ("blk_Qc", o.blk_Qc)):
        check_scaled(f"l96 device filter pinned {k}", t[k], ref)


def test_l96_fused_blocked_and_philox(pkg, orc):
    spec = pkg.problems.l96_spec(n_obs=4, C=3, dt=2e-3, rho=0.8)
    o, e = _pair(pkg, orc, spec, seed=1234)
    P = pkg.problems
    o.init_paths(seed=1234)
    e.init_paths()
    check("l96 philox init ll", e.get_ll(), o.ll)
    im = interior_mask(o.N)
    for it in range(4):
        lay = P.chequerboard_layout(o.rec_nobs, 1, it % 2)
        o.relayout(*lay)
        e.set_layout(*lay)
        e.upload_tables(o.H, o.F0, o.Psi, o.c0, o.blk_g, o.blk_Qc)
        llp_o, ll_o, acc_o = o.impute(seed=1234, it=it)
        llp_g, ll_g, acc_g = e.impute(it)
        check(f"l96 fused blocked it {it} ll deg", llp_g, llp_o)
        assert np.array_equal(acc_g, acc_o)
        check(f"l96 fused blocked it {it} ll", ll_g, ll_o)
        Xg, _ = e.download_state(0)
        check(f"l96 fused blocked it {it} X", Xg[:, im], o.accepted_X()[:, im])


def test_l96_cta_sizes_are_bit_identical(pkg, monkeypatch):
    """The number of warps per CTA must never change a sample (the per-chain arithmetic is the same, operation for
    operation): stored-noise sweeps, fused blocked sweeps, parameter-step solve, layout switch -- ragged chain count."""
    from diffusionmcmc_jl_b200.engine import Engine
    spec = pkg.problems.l96_spec(n_obs=6, C=11, dt=1e-3, rho=0.8)
    out = []
    for nw in ("1", "2", "8"):
        monkeypatch.setenv("DMCMC_L96_WARPS", nw)
        e = Engine(spec, seed=77)
        e.init_paths()
        e.run_sweeps(0, 2, 0)
        e.run_sweeps(10, 3, 1)
        X, W = e.download_state(0)
        llp, ok = e.param_loglik(np.array(spec["theta"]) * (1 + 1e-3), critical=True)
        e.param_commit(False)
        e.relayout(*pkg.problems.chequerboard_layout(e.rec_nobs, 2, 1))
        _, Wr = e.download_state(0)
        out.append((X, W, e.get_ll().copy(), llp, ok, Wr, e.accept_counts()[0].copy()))
        e.close()
    for other in out[1:]:
        for a, b, what in zip(out[0], other, ("X", "W", "ll", "parameter-step ll", "success", "W after a layout switch", "accept counts")):
            assert np.array_equal(a, b), f"{what} differs between CTA sizes"
