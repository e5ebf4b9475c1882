"""Pins for the CPU oracle (the reference has no golden vectors for the arithmetic, so the
oracle is pinned by published/analytic answers and by an independent NumPy/SciPy restatement)."""
import numpy as np
import pytest
from scipy.integrate import solve_ivp

from helpers import close, interior_mask


def test_philox_known_answers(orc):
    # Random123 kat_vectors, philox4x32-10
    kat = [([0, 0, 0, 0], [0, 0], [0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8]),
           ([0xffffffff] * 4, [0xffffffff] * 2, [0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd]),
           ([0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344], [0xa4093822, 0x299f31d0],
            [0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1])]
    for ctr, key, out in kat:
        assert list(orc.philox4x32_10(ctr, key)) == out


def test_philox_normals_are_standard(orc):
    z = np.array([orc.philox_normal4(42, 3, j, c, 0, q) for j in range(4) for c in range(50)
                  for q in range(25)]).ravel()
    assert abs(z.mean()) < 0.02 and abs(z.std() - 1.0) < 0.02
    # the single-precision Box-Muller is a faithful normal: compare quantiles with the exact ones
    from scipy.stats import norm, kstest
    assert kstest(z, "norm").pvalue > 1e-3
    assert abs(np.mean(z ** 4) - 3.0) < 0.25
    e = np.array([orc.philox_exponential(42, it, 0, c) for it in range(40) for c in range(50)])
    assert abs(e.mean() - 1.0) < 0.06 and e.min() > 0


def _ode_rhs(t, y, B, beta, a, d):
    H = y[:d * d].reshape(d, d)
    F = y[d * d:d * d + d]
    dH = -B.T @ H - H @ B + H @ a @ H
    dF = -B.T @ F + H @ a @ F + H @ beta
    dc = beta @ F + 0.5 * F @ a @ F - 0.5 * np.trace(H @ a)
    return np.concatenate([dH.ravel(), dF, [dc]])


@pytest.mark.parametrize("name", ["fhn", "l63"])
def test_backward_filter_solves_the_HFc_odes(pkg, orc, name):
    """SURVEY a8 / App. B state the backward ODEs; the oracle's exact Gaussian recursion must
    agree with a high-accuracy ODE integration of them (independent SciPy restatement)."""
    P = pkg.problems
    spec = P.fhn_spec(n_obs=4, C=1, dt=1e-2) if name == "fhn" else P.l63_spec(n_obs=3, C=1, dt=2e-3)
    o = orc.Oracle(spec)
    d = o.d
    y = np.zeros(d * d + d + 1)
    worst = 0.0
    for j in range(o.J - 1, -1, -1):
        y[:d * d] += o.Hobs[j].ravel()
        y[d * d:d * d + d] += o.Fobs[j]
        y[-1] += o.cobs[j]
        tg = o.t[o.pt_off[j]:o.pt_off[j + 1]]
        sol = solve_ivp(_ode_rhs, (tg[-1], tg[0]), y, t_eval=tg[::-1], method="DOP853",
                        args=(o.auxB[j], o.auxbeta[j], o.auxatil[j], d), rtol=1e-12, atol=1e-13)
        Y = sol.y[:, ::-1]
        sl = slice(o.pt_off[j], o.pt_off[j + 1])
        worst = max(worst, np.abs(Y[:d * d].T.reshape(-1, d, d) - o.H[sl]).max(),
                    np.abs(Y[d * d:d * d + d].T - o.F0[sl]).max(), np.abs(Y[-1] - o.c0[sl]).max())
        y = Y[:, 0].copy()
    assert worst < 5e-9 * max(1.0, np.abs(o.H).max())


def test_linear_target_gives_kalman_marginal_likelihood(pkg, orc):
    """Target == aux law => G == 0, ll = log h~(t_0, x_0) = log p(v_1..v_n | x_0), which a Kalman
    filter gives independently; every proposal is accepted (SURVEY section 4 iii)."""
    spec = pkg.problems.lin2_spec(C=3)
    o = orc.Oracle(spec)
    o.init_paths(seed=1)
    B, beta, a = o.auxB[0], o.auxbeta[0], o.auxatil[0]
    m, Pm, ll_kf = spec["x0"][0, 0].copy(), np.zeros((2, 2)), 0.0
    for j in range(o.J):
        T = o.t[o.pt_off[j + 1] - 1] - o.t[o.pt_off[j]]
        Phi, Q, mv = orc.transition(B, beta, a, T)
        m, Pm = Phi @ m + mv, Phi @ Pm @ Phi.T + Q
        L = o.obs_L[j]
        S = L @ Pm @ L.T + o.obs_Sigma[j]
        r = o.obs_v[j] - L @ m
        ll_kf += -0.5 * r @ np.linalg.solve(S, r) - 0.5 * np.log(np.linalg.det(2 * np.pi * S))
        K = Pm @ L.T @ np.linalg.inv(S)
        m, Pm = m + K @ r, Pm - K @ L @ Pm
    assert np.allclose(o.ll, ll_kf, rtol=0, atol=1e-10)
    rng = np.random.default_rng(3)
    for _ in range(3):
        llp, ll, acc = o.impute(rng.standard_normal((o.C, o.S, o.m)), rng.exponential(size=(o.C, 1)))
        assert np.allclose(llp, ll_kf, atol=1e-10) and acc.all()


def _numpy_solve(o, j, W, y1):
    """Independent NumPy restatement of solve_and_ll for FHN, un-pinned."""
    th = o.theta
    x, ll = y1.copy(), 0.0
    X = [x.copy()]
    base = o.pt_off[j]
    for k in range(o.N[j]):
        dt = o.t[base + k + 1] - o.t[base + k]
        r = o.F0[base + k] - o.H[base + k] @ x
        b = np.array([(x[0] - x[0] ** 3 - x[1] + th[1]) / th[0], th[2] * x[0] - x[1] + th[3]])
        bt = o.auxB[j] @ x + o.auxbeta[j]
        sig = np.array([0.0, th[4]])
        ll += (b - bt) @ r * dt
        x = x + (b + np.outer(sig, sig) @ r) * dt + sig * (W[k + 1, 0] - W[k, 0])
        X.append(x.copy())
    return np.array(X), ll


def test_numpy_twin_of_the_solve(pkg, orc):
    spec = pkg.problems.fhn_spec(n_obs=3, C=2, dt=1e-2)
    o = orc.Oracle(spec)
    rng = np.random.default_rng(0)
    Z0 = rng.standard_normal((o.C, o.S, o.m))
    o.init_paths(Z=Z0)
    Z = rng.standard_normal((o.C, o.S, o.m))
    E = rng.exponential(size=(o.C, 1))
    Wacc = o.accepted_W().copy()
    Xacc = o.accepted_X().copy()
    llp, _, _ = o.impute(Z, E)
    Xp = o.proposal_X() if False else None
    for c in range(o.C):
        y1 = Xacc[c, 0]
        ll = -o.c0[0] - 0.5 * y1 @ o.H[0] @ y1 + o.F0[0] @ y1
        for j in range(o.J):
            sl = slice(o.pt_off[j], o.pt_off[j + 1])
            tg = o.t[sl]
            Wf = np.concatenate([[0.0], np.cumsum(np.sqrt(np.diff(tg)) * Z[c, o.st_off[j]:o.st_off[j + 1], 0])])
            Wo = np.sqrt(1 - o.rho[j] ** 2) * Wf + o.rho[j] * Wacc[c, sl, 0]
            X, l = _numpy_solve(o, j, Wo[:, None], y1)
            ll += l
            y1 = X[-1]
        assert abs(ll - llp[c, 0]) <= 1e-11 * max(1.0, abs(ll))


def test_rho_one_reproduces_the_accepted_path(pkg, orc):
    spec = pkg.problems.fhn_spec(n_obs=5, C=4, dt=1e-2)
    o = orc.Oracle(spec)
    o.init_paths(seed=3)
    o.rho[:] = 1.0
    X0, ll0 = o.accepted_X().copy(), o.ll.copy()
    llp, ll, acc = o.impute(seed=3, it=1)
    assert np.array_equal(llp, ll0) and acc.all()
    assert np.array_equal(o.accepted_X(), X0)


def test_swap_is_a_selector_flip(pkg, orc):
    """Accept = swap of container references (src/run!_alterations.jl:81-91): rejected units keep
    their accepted containers bit for bit."""
    spec = pkg.problems.fhn_spec(n_obs=5, C=16, dt=1e-2, rho=0.3)
    o = orc.Oracle(spec)
    o.init_paths(seed=3)
    X0, W0 = o.accepted_X().copy(), o.accepted_W().copy()
    llp, ll, acc = o.impute(seed=3, it=1)
    assert acc.any() and (~acc).any()
    rej = ~acc[:, 0]
    assert np.array_equal(o.accepted_X()[rej], X0[rej]) and np.array_equal(o.accepted_W()[rej], W0[rej])
    assert not np.array_equal(o.accepted_X()[~rej], X0[~rej])


def test_pinned_tables_match_an_explicit_exact_observation(pkg, orc):
    """F0 + Psi x_b, c0 + g'x_b + x_b'Qc x_b/2 equal a direct filter whose last observation of the
    block is the exact observation (L = I, Sigma = eps I, v = x_b)."""
    P = pkg.problems
    spec = P.fhn_spec(n_obs=6, C=1, dt=1e-2, pin_eps=1e-3)
    o = orc.Oracle(spec)
    i1, i2, pin = P.chequerboard_layout(o.rec_nobs, 2, 0)
    o.set_layout(i1, i2, pin)
    xb = np.array([0.3, -0.7])
    b = 0  # first block [0,3], pinned
    assert pin[b] == 1
    F_tab = o.F0 + np.einsum("pij,j->pi", o.Psi, xb)
    c_tab = o.c0[o.pt_off[i1[b]]] + o.blk_g[b] @ xb + 0.5 * xb @ o.blk_Qc[b] @ xb
    # direct: truncate the problem at the block end with an exact observation there
    sub = P.make_spec(P.MODEL_FHN, spec["theta"], spec["x0"][0, 0], spec["obs_times"][:i2[b] + 1],
                      spec["obs_v"][:i2[b] + 1], [[1.0, 0.0]], [[0.01]], 1e-2)
    sub["obs_L"] = np.concatenate([np.pad(sub["obs_L"][:-1], ((0, 0), (0, 1), (0, 0))),
                                   np.eye(2)[None]])
    Sg = np.zeros((i2[b] + 1, 2, 2))
    Sg[:-1, 0, 0] = 0.01
    Sg[:-1, 1, 1] = 1.0
    Sg[-1] = 1e-3 * np.eye(2)
    sub["obs_Sigma"] = Sg
    sub["obs_v"] = np.concatenate([np.pad(sub["obs_v"][:-1], ((0, 0), (0, 1))), xb[None]])
    o2 = orc.Oracle.__new__(orc.Oracle)
    # the padded second observation row has L = 0, so it only adds a constant to c
    o2.__init__(sub)
    # the aux law of the last interval must still be linearised at the DATA observation
    custom = [(o.auxB[j], o.auxbeta[j], o.auxsig[j]) for j in range(i2[b] + 1)]
    o2.set_theta(spec["theta"], custom_aux=custom)
    sl = slice(0, o.pt_off[i2[b] + 1])
    assert np.abs(o2.H[sl] - o.H[sl]).max() < 1e-12 * np.abs(o.H[sl]).max()
    assert np.abs(o2.F0[sl] - F_tab[sl]).max() < 1e-12 * np.abs(F_tab[sl]).max()
    const = sum(0.5 * np.log(2 * np.pi * 1.0) for _ in range(i2[b]))  # padded rows: v=0, Sigma=1
    assert abs((o2.c0[0] - const) - c_tab) < 1e-12 * max(1.0, abs(c_tab))


def test_backward_filter_schemes_agree(pkg, orc):
    """The two exact flows of the H, F, c ODEs (step by step; every grid point from the interval's right end,
    which is what the device computes for d <= 3) differ by rounding only, pinned layouts included."""
    P = pkg.problems
    for mk, h in ((lambda: P.fhn_spec(n_obs=9, C=1, dt=0.004), 1), (lambda: P.l63_spec(n_obs=12, C=1, dt=1e-3), 2),
                  (lambda: P.lv_spec(n_obs=10, C=1, dt=0.01), 3), (lambda: P.fhn_spec(n_obs=12, C=1, dt=1e-3), 0)):
        spec = mk()
        a, b = orc.Oracle(spec, bf_scheme=0), orc.Oracle(spec, bf_scheme=1)
        for pat in (0, 1):
            lay = P.chequerboard_layout(a.rec_nobs, h, pat)
            a.set_layout(*lay)
            b.set_layout(*lay)
            for k in ("H", "F0", "Psi", "c0", "blk_g", "blk_Qc"):
                x, y = getattr(a, k), getattr(b, k)
                assert np.abs(x - y).max() <= 1e-11 * max(1.0, np.abs(x).max()), (k, np.abs(x - y).max())


def test_relayout_roundtrip(pkg, orc):
    """Re-inverting W under the law that generated X recovers W (to rounding) and ll."""
    P = pkg.problems
    spec = P.fhn_spec(n_obs=6, C=3, dt=1e-2)
    o = orc.Oracle(spec)
    o.init_paths(seed=5)
    W0, ll0 = o.accepted_W().copy(), o.ll.copy()
    o.relayout(*P.chequerboard_layout(o.rec_nobs, 0, 0))
    assert np.abs(o.accepted_W() - W0).max() < 1e-12
    assert np.abs(o.ll - ll0).max() < 1e-10
    # blocked layout: sum of block ll's is finite, W reproduces X through the new law
    o.relayout(*P.chequerboard_layout(o.rec_nobs, 1, 1))
    o.rho[:] = 1.0
    X0 = o.accepted_X().copy()
    llp, ll, acc = o.impute(seed=5, it=9)
    im = interior_mask(o.N)
    Xn = np.where(acc.any(), o.accepted_X(), o.proposal_X())
    err = np.abs(o.accepted_X()[:, im] - X0[:, im]).max()
    print("relayout round trip: max |X - X0| =", err)
    assert err < 1e-12


def test_chequerboard_layout_is_an_exact_partition(pkg, orc):
    """Block partitioning must be bit-exact: host mirror == oracle, blocks tile the intervals."""
    P = pkg.problems
    for nobs in ([100], [7], [1], [5, 8, 3]):
        for h in (0, 1, 2, 3, 5):
            for pat in (0, 1):
                a = P.chequerboard_layout(nobs, h, pat)
                b = orc.chequerboard_layout(nobs, h, pat)
                for x, y in zip(a, b):
                    assert np.array_equal(x, y)
                i1, i2, pin = a
                assert i1[0] == 0 and i2[-1] == sum(nobs) - 1
                assert np.array_equal(i1[1:], i2[:-1] + 1)
                ends = np.cumsum(nobs) - 1
                assert np.array_equal(pin == 0, np.isin(i2, ends))


def test_glue_path_drops_duplicated_joins(pkg, orc):
    spec = pkg.problems.fhn_spec(n_obs=4, C=2, dt=1e-2)
    o = orc.Oracle(spec)
    o.init_paths(seed=2)
    g = o.glue_path(1)
    X = o.accepted_X()[1]
    assert g.shape == (int(o.N.sum()) + 1, 2)
    ref = pkg.glue_containers([X[o.pt_off[j]:o.pt_off[j + 1]] for j in range(o.J)])
    assert np.array_equal(g, ref)
    assert np.array_equal(g[0], spec["x0"][1, 0])
