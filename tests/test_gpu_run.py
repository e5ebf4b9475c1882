"""`run_` (the reference's `run!`) end to end on the GPU: smoothing, adaptation, blocking,
parameter inference.  Statistical sanity mirrors the docs' informal acceptance tests
(docs/src/tutorials/simple_inference.md:105-117, :160-161, :189-192)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def test_smoothing_like_the_tutorial(pkg):
    data = pkg.problems.fhn_spec(n_obs=100, C=1, dt=1e-3, rho=0.96)
    mcmc = pkg.MCMC([pkg.PathImputation(0.96, "FitzHughNagumoAux")], backend=pkg.DiffusionMCMCBackend())
    gws, lws = pkg.run_(mcmc, 1000, data, None, path_buffer_size=10, seed=3)
    acc = lws[0].acceptance_history           # [iter][block]
    assert acc.shape == (1000, 1)
    assert 0.05 < acc.mean() < 0.95
    ll = lws[0].ll_history
    assert np.all(np.isfinite(ll)) and ll[-1, 0] > ll[0, 0] - 50
    xs = gws.XX_buffer.XX                     # ring of 10 glued paths, saved every 100 iterations
    assert len(xs) == 10 and xs[0][0]["x"].shape == (100 * 100 + 1, 2)
    assert np.allclose(xs[0][0]["t"][[0, -1]], [0.0, 10.0])
    # the smoothed first coordinate follows the observations (sd 0.1)
    v = data["obs_v"][:, 0]
    at_obs = xs[9][0]["x"][100::100, 0]
    assert np.sqrt(np.mean((at_obs - v) ** 2)) < 0.3
    # consecutive saved paths differ (the chain moves)
    assert not np.array_equal(xs[8][0]["x"], xs[9][0]["x"])


def test_adaptation_moves_rho_towards_the_target(pkg):
    """docs: rho = 0.1 never accepts (0.0); adaptation drives rho up to ~0.95 and the rate to ~0.23."""
    data = pkg.problems.fhn_spec(n_obs=50, C=1, dt=1e-3)
    fixed = pkg.MCMC([pkg.PathImputation(0.1, "FitzHughNagumoAux")])
    _, lws = pkg.run_(fixed, 300, data, None, seed=1)
    assert lws[0].acceptance_history.mean() < 0.02
    adpt = pkg.AdaptationPathImputation(adapt_every_k_steps=50, scale=1.0, offset=0.0)
    u = pkg.PathImputation(0.1, "FitzHughNagumoAux", adpt=adpt)
    _, lws = pkg.run_(pkg.MCMC([u]), 3000, data, None, seed=1)
    assert u.rho_s[0][0] > 0.8
    assert 0.08 < lws[0].acceptance_history[-500:].mean() < 0.6


def test_blocked_smoothing_matches_unblocked_in_distribution(pkg):
    """Chequerboard blocking (ours; the reference has only the decorator) targets the same
    smoothing distribution: posterior means of the path agree with the unblocked sampler."""
    data = pkg.problems.fhn_spec(n_obs=20, C=256, dt=2e-3, rho=0.9, pin_eps=1e-4)
    un = pkg.MCMC([pkg.PathImputation(0.9, "FitzHughNagumoAux")])
    bl = pkg.MCMC([pkg.BlockingChequerboardSwitch(2), pkg.PathImputation(0.9, "FitzHughNagumoAux"),
                   pkg.BlockingChequerboardSwitch(2), pkg.PathImputation(0.9, "FitzHughNagumoAux")])
    g1, l1 = pkg.run_(un, 400, data, None, n_chains=256, seed=11)
    g2, l2 = pkg.run_(bl, 400, data, None, n_chains=256, seed=12)
    assert l2[0].acceptance_history.shape == (400, 256, 5) and l2[1].acceptance_history.shape == (400, 256, 6)
    assert l2[0].acceptance_history.mean() > 0.1
    X1, _ = g1.engine.download_state(0)
    X2, _ = g2.engine.download_state(0)
    m1, m2 = X1.mean(axis=0), X2.mean(axis=0)
    s1 = X1.std(axis=0)
    # means over 256 independent chains: agree within a few standard errors (+ pin_eps softness)
    err = np.abs(m1 - m2) / (s1 / np.sqrt(256) * 2 ** 0.5 + 0.02)
    assert np.quantile(err, 0.99) < 6.0, np.quantile(err, [0.5, 0.9, 0.99, 1.0])


def test_parameter_inference_recovers_gamma(pkg):
    """PathImputation + RandomWalkUpdate on one drift parameter (config 2 pattern): the chain
    moves from a wrong start towards the data-generating value."""
    data = pkg.problems.fhn_spec(n_obs=100, C=1, dt=2e-3, rho=0.95)
    th0 = data["theta"].copy()
    th0[2] = 1.0  # gamma, truth 1.5
    mcmc = pkg.MCMC([pkg.PathImputation(0.95, "FitzHughNagumoAux"),
                     pkg.RandomWalkUpdate(0.1, [2])])
    gws, lws = pkg.run_(mcmc, 1500, data, th0, seed=5)
    th = np.array(gws.state_history)
    assert th.shape == (1500, 5)
    assert lws[1].acceptance_history.mean() > 0.02
    assert abs(th[-500:, 2].mean() - 1.5) < 0.35, th[-500:, 2].mean()
    assert np.all(th[:, [0, 1, 3, 4]] == th0[[0, 1, 3, 4]])


def test_single_update_with_switch_alternates_patterns(pkg):
    """[BlockingChequerboardSwitch, PathImputation]: the pattern alternates from one MCMC iteration to the next, and the
    shifted pattern has one block more -- the histories must hold both (regression: the row stride was that of the
    first pattern only)."""
    data = pkg.problems.fhn_spec(n_obs=10, C=3, dt=0.01, rho=0.9)
    mcmc = pkg.MCMC([pkg.BlockingChequerboardSwitch(1), pkg.PathImputation(0.9, "FitzHughNagumoAux")])
    gws, lws = pkg.run_(mcmc, 7, data, seed=3)
    acc, ll = lws[0].acceptance_history, lws[0].ll_history
    assert acc.shape == (7, 3, 6) and ll.shape == (7, 3, 6)      # 5 blocks (pattern 0) / 6 blocks (pattern 1)
    assert np.isfinite(ll[:, :, :5]).all()
    assert np.isfinite(ll[1::2, :, 5]).all()                     # the sixth block exists in every other iteration


def test_odd_switch_count_with_a_parameter_update_moves_the_block_ends(pkg):
    """[Switch, PathImputation, RandomWalkUpdate] walks the general path (update by update).  The pattern must alternate
    from one MCMC iteration to the next there as well, otherwise the points at the pattern-0 block ends are pinned
    forever and never re-imputed."""
    data = pkg.problems.fhn_spec(n_obs=12, C=1, dt=0.01, rho=0.5)
    mcmc = pkg.MCMC([pkg.BlockingChequerboardSwitch(1), pkg.PathImputation(0.5, "FitzHughNagumoAux"),
                     pkg.RandomWalkUpdate(0.05, [2])])
    saved = []

    def cb(gws, lws, step):
        if step.pidx == 2:
            saved.append(gws.engine.download_path(0))
    gws, lws = pkg.run_(mcmc, 40, data, data["theta"].copy(), callbacks=[cb], seed=4)
    P = np.array(saved)                        # [iter][point][d]; 10 steps per interval
    ends0 = np.arange(2, 12, 2) * 10           # pattern-0 block ends (pinned in pattern 0, interior in pattern 1)
    moved = np.abs(np.diff(P[:, ends0, 1], axis=0)).max(axis=0)
    assert (moved > 0).all(), moved
    acc = lws[0].acceptance_history
    assert acc.shape == (40, 7)                # pattern 0: 6 blocks, pattern 1: 7 blocks
    assert np.isnan(lws[0].ll_history[0::2, 6]).all() and np.isfinite(lws[0].ll_history[1::2, 6]).all()
    assert acc[1::2].any() and acc[0::2].any()


def test_two_adaptive_imputations_both_adapt_and_save_several_chains(pkg):
    """Fast path with two PathImputation updates: each keeps its own acceptance counters (from its own history rows),
    so both rho's move; the path ring holds the chosen chains."""
    data = pkg.problems.fhn_spec(n_obs=10, C=8, dt=0.01)
    mk = lambda r: pkg.PathImputation(r, "FitzHughNagumoAux",
                                      adpt=pkg.AdaptationPathImputation(adapt_every_k_steps=20, scale=1.0, offset=0.0))
    u1, u2 = mk(0.05), mk(0.999999)
    mcmc = pkg.MCMC([pkg.BlockingChequerboardSwitch(2), u1, pkg.BlockingChequerboardSwitch(2), u2])
    gws, lws = pkg.run_(mcmc, 200, data, n_chains=8, seed=6, save_every=50, path_buffer_size=3, save_chains=(0, 5))
    assert u1.rho_s[0][0] != 0.05 and u2.rho_s[0][0] < 0.999999      # both adapted (the second one used to see 0/0)
    assert len(u1.rho_s) == 3 and len(u2.rho_s) == 3          # 10 obs, half_len 2: 3 blocks in either pattern
    x = gws.XX_buffer.XX[0][0]["x"]
    assert x.shape == (2, 101, 2) and np.isfinite(x).all() and not np.array_equal(x[0], x[1])
    last = gws.XX_buffer.XX[0][0]["x"]           # ring of 3, 4 saves: slot 0 holds the 4th (iteration 200 = final state)
    assert np.array_equal(last[0], gws.engine.download_path(0)) and np.array_equal(last[1], gws.engine.download_path(5))


def test_device_side_adaptation_matches_the_host_rule(pkg):
    """dmcmc_set_adaptation: the kernel's logit step equals AdaptationPathImputation.readjust (same formula, CUDA vs
    libm exp / log: 1e-12), and run_ with the adaptation on the device drives rho to the same region as with the
    host-side adaptation."""
    from diffusionmcmc_jl_b200.engine import Engine
    data = pkg.problems.fhn_spec(n_obs=10, C=64, dt=0.01)
    e = Engine(data, seed=9)
    e.init_paths()
    cfg = pkg.AdaptationPathImputation(adapt_every_k_steps=2 * 64, scale=0.7, offset=0.0)
    rho0 = np.full(e.J, 0.6)
    e.set_adaptation(cfg, 1, rho0, rho0)
    n, stride = 4, 6
    hist = (np.empty((n, 64, stride)), np.empty((n, 64, stride)), np.empty((n, 64, stride), np.uint8))
    e.run_sweeps(0, n, half_len=1, first_pattern=0, hist=hist)       # sweeps 0, 2: pattern 0 (5 blocks); 1, 3: pattern 1 (6)
    for pat, nb in ((0, 5), (1, 6)):
        lay = pkg.problems.chequerboard_layout(e.rec_nobs, 1, pat)
        acc = hist[2][pat::2, :, :nb].astype(np.int64).sum(axis=(0, 1))   # two sweeps of 64 proposals per block = k
        got = e.get_rho(pat)
        for b in range(nb):
            a = pkg.AdaptationPathImputation(adapt_every_k_steps=128, scale=0.7, offset=0.0)
            a.register(int(acc[b]), 128)
            assert a.time_to_update()
            want = a.readjust(0.6, 3 + pat)       # MCMC iteration of the adapting sweep: sweep index (2 + pat) + 1
            assert np.all(np.abs(got[lay[0][b]:lay[1][b] + 1] - want) <= 1e-12), (pat, b, got[lay[0][b]], want)
    e.close()
    # end to end: both ways of adapting end up in the same region
    out = {}
    for dev in (True, False):
        u = pkg.PathImputation(0.3, "FitzHughNagumoAux", adpt=pkg.AdaptationPathImputation(adapt_every_k_steps=50, scale=1.0, offset=0.0))
        gws, lws = pkg.run_(pkg.MCMC([u]), 1500, pkg.problems.fhn_spec(n_obs=50, C=1, dt=1e-3), None, seed=1, device_adaptation=dev)
        out[dev] = (u.rho_s[0][0], lws[0].acceptance_history[-500:].mean())
    assert out[True][0] > 0.8 and out[False][0] > 0.8, out
    assert abs(out[True][1] - out[False][1]) < 0.2, out
