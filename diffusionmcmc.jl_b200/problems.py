"""Synthetic observation sets for the five BASELINE.json configs (host side, NumPy only).

A "spec" is a plain dict of NumPy arrays describing one inference problem in the reference's
vocabulary: recordings, observations (L, Sigma, v at times t_j), imputation grids per
inter-observation interval (what `OBS.setup_time_grids` hands to
`DiffusionGlobalSubworkspace`, /root/reference src/workspaces.jl:164-170), known starting
points (`KnownStartingPt`, docs/src/tutorials/simple_inference.md:52-60) and the pCN memory
`rho` per interval (src/updates.jl:53-74).

The reference's RNG streams (Julia MersenneTwister) cannot be reproduced, so data are generated
with NumPy PCG64 from a fixed seed.
"""
import numpy as np

MODEL_FHN, MODEL_LV, MODEL_L63, MODEL_L96, MODEL_LIN2, MODEL_LVM = range(6)
MODEL_NAMES = {"fhn": MODEL_FHN, "lv": MODEL_LV, "l63": MODEL_L63, "l96": MODEL_L96,
               "lin2": MODEL_LIN2, "lvm": MODEL_LVM}


def model_dims(model, d_in=0):
    return {MODEL_FHN: (2, 1, 5), MODEL_LV: (2, 2, 6), MODEL_L63: (3, 3, 4),
            MODEL_L96: (d_in, d_in, 3), MODEL_LIN2: (2, 2, 8), MODEL_LVM: (2, 2, 6)}[model]


def _drift(model, th, x):
    if model == MODEL_FHN:
        return np.array([(x[0] - x[0] ** 3 - x[1] + th[1]) / th[0], th[2] * x[0] - x[1] + th[3]])
    if model in (MODEL_LV, MODEL_LVM):
        return np.array([th[0] * x[0] - th[1] * x[0] * x[1], th[3] * x[0] * x[1] - th[2] * x[1]])
    if model == MODEL_L63:
        return np.array([th[0] * (x[1] - x[0]), th[1] * x[0] - x[1] - x[0] * x[2],
                         x[0] * x[1] - th[2] * x[2]])
    if model == MODEL_L96:
        return (np.roll(x, -1) - np.roll(x, 2)) * np.roll(x, 1) - x + th[0]
    if model == MODEL_LIN2:
        return np.array([th[0] * x[0] + th[1] * x[1] + th[4], th[2] * x[0] + th[3] * x[1] + th[5]])
    raise ValueError(model)


def _sigma(model, th, x, d, m):
    s = np.zeros((d, m))
    if model == MODEL_FHN:
        s[1, 0] = th[4]
    elif model == MODEL_LV:
        s[0, 0], s[1, 1] = th[4], th[5]
    elif model == MODEL_LVM:
        s[0, 0], s[1, 1] = th[4] * x[0], th[5] * x[1]
    elif model == MODEL_L63:
        s[np.arange(3), np.arange(3)] = th[3]
    elif model == MODEL_L96:
        s[np.arange(d), np.arange(d)] = th[1]
    elif model == MODEL_LIN2:
        s[0, 0], s[1, 1] = th[6], th[7]
    return s


def simulate(model, th, x0, t_end, dt, rng):
    """Euler-Maruyama truth on a fine grid; returns (t, X)."""
    d = len(x0)
    m = model_dims(model, d)[1]
    n = int(round(t_end / dt))
    X = np.empty((n + 1, d))
    X[0] = x0
    sq = np.sqrt(dt)
    for k in range(n):
        x = X[k]
        X[k + 1] = x + _drift(model, th, x) * dt + _sigma(model, th, x, d, m) @ (sq * rng.standard_normal(m))
    return np.arange(n + 1) * dt, X


def setup_time_grids(t0, obs_times, dt):
    """One uniform grid per inter-observation interval: N_j = max(1, round(len/dt)) steps."""
    ts, Ns = [], []
    a = t0
    for b in obs_times:
        n = max(1, int(round((b - a) / dt)))
        ts.append(a + (b - a) * (np.arange(n + 1) / n))
        Ns.append(n)
        a = b
    return np.array(Ns, np.int32), np.concatenate(ts)


def make_spec(model, theta, x0, obs_times, obs_v, L, Sigma, dt, C=1, rho=0.96, pin_eps=1e-4,
              t0=0.0):
    """Single recording.  obs_v [n][p]; L [p][d]; Sigma [p][p]."""
    model = MODEL_NAMES.get(model, model)
    x0 = np.atleast_1d(np.asarray(x0, float))
    d = len(x0)
    _, m, _ = model_dims(model, d)
    obs_v = np.atleast_2d(np.asarray(obs_v, float))
    n = len(obs_times)
    N, t = setup_time_grids(t0, obs_times, dt)
    L, Sigma = np.atleast_2d(np.asarray(L, float)), np.atleast_2d(np.asarray(Sigma, float))
    rec_first = np.zeros(n, np.int32)
    rec_first[0] = 1
    return dict(model=model, d=d, m=m, theta=np.asarray(theta, float), N=N, t=t,
                rec_first=rec_first, obs_L=np.broadcast_to(L, (n,) + L.shape).copy(),
                obs_Sigma=np.broadcast_to(Sigma, (n,) + Sigma.shape).copy(), obs_v=obs_v,
                x0=np.broadcast_to(x0, (C, 1, d)).copy(), rho=np.full(n, float(rho)),
                pin_eps=float(pin_eps), obs_times=np.asarray(obs_times, float), t0=float(t0))


def concat_recordings(specs):
    """Several recordings of the same model -> one spec (recordings are conditionally
    independent blocks, src/run!_alterations.jl:15-16)."""
    s0 = dict(specs[0])
    for k in ("N", "t", "rec_first", "obs_L", "obs_Sigma", "obs_v", "rho"):
        s0[k] = np.concatenate([s[k] for s in specs])
    s0["x0"] = np.concatenate([s["x0"] for s in specs], axis=1)
    return s0


# ----------------------------------------------------------------------------------------------
# the five named configs (BASELINE.json "configs").  `scale` < 1 shrinks the number
# of observations for parity tests; bench.py uses the full sizes.
# ----------------------------------------------------------------------------------------------
FHN_THETA = np.array([0.1, -0.8, 1.5, 0.0, 0.3])  # docs/src/tutorials/simple_inference.md:30


def fhn_spec(n_obs=50, C=1, dt=1e-3, rho=0.96, seed=1, obs_every=0.1, pin_eps=1e-4):
    rng = np.random.default_rng(seed)
    x0 = np.array([-1.0, -1.0])
    tt, X = simulate(MODEL_FHN, FHN_THETA, x0, n_obs * obs_every, 1e-4, rng)
    stride = int(round(obs_every / 1e-4))
    idx = np.arange(1, n_obs + 1) * stride
    v = X[idx, 0] + 0.1 * rng.standard_normal(n_obs)
    return make_spec(MODEL_FHN, FHN_THETA, x0, tt[idx], v[:, None], [[1.0, 0.0]], [[0.01]], dt,
                     C=C, rho=rho, pin_eps=pin_eps)


LV_THETA = np.array([1.5, 1.0, 3.0, 1.0, 0.1, 0.1])


def lv_spec(n_obs=200, C=1, dt=0.01, rho=0.9, seed=2, obs_every=0.1, pin_eps=1e-4,
            multiplicative=False):
    rng = np.random.default_rng(seed)
    model = MODEL_LVM if multiplicative else MODEL_LV
    x0 = np.array([2.0, 1.5])
    tt, X = simulate(model, LV_THETA, x0, n_obs * obs_every, 1e-3, rng)
    stride = int(round(obs_every / 1e-3))
    idx = np.arange(1, n_obs + 1) * stride
    v = X[idx] + 0.05 * rng.standard_normal((n_obs, 2))
    return make_spec(model, LV_THETA, x0, tt[idx], v, np.eye(2), 0.05 ** 2 * np.eye(2), dt,
                     C=C, rho=rho, pin_eps=pin_eps)


L63_THETA = np.array([10.0, 28.0, 8.0 / 3.0, 1.0])


def l63_spec(n_obs=10000, C=1, dt=1e-3, rho=0.9, seed=4, obs_every=0.01, pin_eps=1e-4):
    rng = np.random.default_rng(seed)
    x0 = np.array([1.5, -1.5, 25.0])
    tt, X = simulate(MODEL_L63, L63_THETA, x0, n_obs * obs_every, 1e-3, rng)
    stride = int(round(obs_every / 1e-3))
    idx = np.arange(1, n_obs + 1) * stride
    v = X[idx] + 0.5 * rng.standard_normal((n_obs, 3))
    return make_spec(MODEL_L63, L63_THETA, x0, tt[idx], v, np.eye(3), 0.25 * np.eye(3), dt,
                     C=C, rho=rho, pin_eps=pin_eps)


def l96_spec(n_obs=4096, C=256, d=40, dt=1e-4, rho=0.9, seed=5, obs_every=0.01, pin_eps=1e-4):
    rng = np.random.default_rng(seed)
    theta = np.array([8.0, 0.5, 2.3])  # F, sigma, aux linearisation level
    x0 = 8.0 + 0.01 * rng.standard_normal(d)
    x0[0] += 1.0
    tt, X = simulate(MODEL_L96, theta, x0, n_obs * obs_every, 1e-3, rng)
    stride = int(round(obs_every / 1e-3))
    idx = np.arange(1, n_obs + 1) * stride
    p = d // 2
    L = np.zeros((p, d))
    L[np.arange(p), 2 * np.arange(p)] = 1.0  # every other coordinate
    v = X[idx] @ L.T + 0.5 * rng.standard_normal((n_obs, p))
    return make_spec(MODEL_L96, theta, x0, tt[idx], v, L, 0.25 * np.eye(p), dt, C=C, rho=rho,
                     pin_eps=pin_eps)


def lin2_spec(n_obs=6, C=1, dt=0.01, rho=0.7, seed=7, obs_every=0.2):
    """Linear target equal to its aux law: G == 0 so ll = log h~(t_a, x_a) (SURVEY section 4 iii)."""
    rng = np.random.default_rng(seed)
    theta = np.array([-0.8, 0.4, -0.3, -1.1, 0.2, -0.1, 0.5, 0.7])
    x0 = np.array([0.5, -0.25])
    tt, X = simulate(MODEL_LIN2, theta, x0, n_obs * obs_every, 1e-3, rng)
    stride = int(round(obs_every / 1e-3))
    idx = np.arange(1, n_obs + 1) * stride
    v = X[idx, :1] + 0.1 * rng.standard_normal((n_obs, 1))
    return make_spec(MODEL_LIN2, theta, x0, tt[idx], v, [[1.0, 0.0]], [[0.01]], dt, C=C, rho=rho)


def chequerboard_layout(rec_nobs, half_len, pattern):
    """Host mirror of the layout tuple ((rec_idx, obs_range), ...) of src/schedule.jl:97-100,
    extended to `BlockingChequerboardSwitch(block_half_len)` (src/blocking.jl:7-9).  Returns
    0-based inclusive global interval ranges and the pinned flag of each block."""
    i1, i2, pin = [], [], []
    base = 0
    for n in rec_nobs:
        n = int(n)
        if half_len <= 0:
            i1.append(base), i2.append(base + n - 1), pin.append(0)
        else:
            start, length = 0, (half_len if pattern else 2 * half_len)
            while start < n:
                end = min(start + length, n)
                i1.append(base + start), i2.append(base + end - 1), pin.append(int(end < n))
                start, length = end, 2 * half_len
        base += n
    return (np.array(i1, np.int32), np.array(i2, np.int32), np.array(pin, np.int32))
