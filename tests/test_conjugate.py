"""Conjugate Gaussian parameter update (/root/reference src/updates.jl:83-105,
:225-273): the path integrals mu, W of compute_mu_and_W.

CPU: the NumPy restatement against a closed form (a path generated WITHOUT noise from the law itself
gives W^-1 mu = theta exactly, because dy = (phi theta + phi_c) dt on every step).
GPU: the device reduction against the NumPy restatement, and an end-to-end run that recovers the
parameters of simulated data."""
import numpy as np
import pytest


def _noise_free_containers(orc, model, th, x0, N, dt):
    """Per-interval containers [1][P][d] of a deterministic Euler path of the law."""
    d = len(x0)
    X, t = [], []
    x, tt = np.array(x0, float), 0.0
    for n in N:
        seg, ts = [x.copy()], [tt]
        for _ in range(n):
            x = x + orc.drift(model, d, th, x) * dt
            tt += dt
            seg.append(x.copy())
            ts.append(tt)
        X += seg
        t += ts
    return np.array(X)[None], np.array(t)


@pytest.mark.parametrize("name", ["fhn", "lv", "l63"])
def test_oracle_conjugate_recovers_theta_on_noise_free_path(orc, name):
    model = {"fhn": orc.MODEL_FHN, "lv": orc.MODEL_LV, "l63": orc.MODEL_L63}[name]
    th = {"fhn": np.array([0.1, -0.8, 1.5, 0.3, 0.3]), "lv": np.array([1.5, 1.0, 3.0, 1.0, 0.1, 0.1]),
          "l63": np.array([10.0, 28.0, 8.0 / 3.0, 1.0])}[name]
    x0 = {"fhn": [-1.0, -1.0], "lv": [2.0, 1.5], "l63": [1.0, 1.5, 20.0]}[name]
    N = [7, 12, 5]
    X, t = _noise_free_containers(orc, model, th, x0, N, 1e-3)
    idx, mu, W = orc.conjugate_sums(model, th, X, N, t)
    est = np.linalg.solve(W[0], mu[0])
    assert np.allclose(est, th[idx], rtol=1e-6), (est, th[idx])  # W of a short path is ill-conditioned for LV


@pytest.mark.gpu
@pytest.mark.parametrize("name,kw", [("fhn", dict(n_obs=6, C=5, dt=0.01)), ("lv", dict(n_obs=5, C=3, dt=0.01)),
                                     ("l63", dict(n_obs=4, C=2, dt=2e-3)), ("lvm", dict(n_obs=4, C=2, dt=0.005))])
def test_conjugate_sums_parity(pkg, orc, name, kw):
    """Device reduction over the accepted containers == NumPy restatement (1e-10 relative), before and
    after blocked sweeps have scattered the accepted path over both containers."""
    from diffusionmcmc_jl_b200.engine import Engine
    P = pkg.problems
    spec = {"fhn": P.fhn_spec, "lv": P.lv_spec, "l63": P.l63_spec,
            "lvm": lambda **k: P.lv_spec(multiplicative=True, **k)}[name](**kw)
    e = Engine(spec, seed=17)
    e.init_paths()
    for rnd in range(2):
        Xg, _ = e.download_state(0)
        idx, mu_o, W_o = orc.conjugate_sums(spec["model"], spec["theta"], Xg, e.N, spec["t"])
        assert list(e.conjugate_coords()) == list(idx)
        mu_g, W_g = e.conjugate_sums()
        sc_mu, sc_W = max(1.0, np.abs(mu_o).max()), max(1.0, np.abs(W_o).max())
        assert np.abs(mu_g - mu_o).max() <= 1e-10 * sc_mu, np.abs(mu_g - mu_o).max() / sc_mu
        assert np.abs(W_g - W_o).max() <= 1e-10 * sc_W, np.abs(W_g - W_o).max() / sc_W
        e.run_sweeps(10 * rnd, 5, half_len=1)


@pytest.mark.gpu
def test_conjugate_update_recovers_parameters(pkg):
    """run_ with [PathImputation, DiffusionConjugGsnUpdate(gamma, beta)] on FitzHugh-Nagumo data
    simulated with (gamma, beta) = (1.5, 0.0): the posterior mean lands near the truth."""
    data = pkg.problems.fhn_spec(n_obs=100, C=1, dt=1e-3, rho=0.96)
    th0 = data["theta"].copy()
    th0[2], th0[3] = 1.0, 0.4
    cu = pkg.DiffusionConjugGsnUpdate([2, 3], [0.0, 0.0], 100.0 * np.eye(2))
    mcmc = pkg.MCMC([pkg.PathImputation(0.96, "FitzHughNagumoAux"), cu], backend=pkg.DiffusionMCMCBackend())
    gws, lws = pkg.run_(mcmc, 400, data, th0, seed=8)
    chain = np.array(gws.state_history)[150:]
    assert lws[1].acceptance_history.mean() > 0.99          # conjugate draws are accepted
    g, b = chain[:, 2].mean(), chain[:, 3].mean()
    assert abs(g - 1.5) < 0.35 and abs(b - 0.0) < 0.35, (g, b)
