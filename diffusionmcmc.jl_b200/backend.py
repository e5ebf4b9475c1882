"""Host-side mirror of the reference's operator interface for the imputation path.

Names follow /root/reference src/DiffusionMCMC.jl:31-33 and ExtensibleMCMC's `run!` surface:
`DiffusionMCMCBackend`, `PathImputation`, `AdaptationPathImputation`, `BlockingChequerboardSwitch`,
`MCMC`, `run_` (Julia's `run!`), returning `(global_ws, local_wss)` like
docs/src/get_started/overview.md:93.  The Julia shim a maintainer would add is in julia/ and
INTEGRATION.md; Julia is not installed in this image, so this Python mirror is what the tests and
bench.py drive.  All arithmetic happens behind the C ABI (include/dmcmc.h) on the GPU.

Differences from the reference that are deliberate:
  * chains are a batch axis (n_chains keyword); histories are [iter][chain][block] and collapse to
    the reference's [iter][block] when n_chains == 1;
  * save_path! fires once per iteration, not once per block (src/run!_alterations.jl:45 quirk);
  * blocking is implemented (the reference has only the decorator and TODOs).
"""
import numpy as np

from .adaptation import (AdaptationMultiPathImputation, AdaptationPathImputation, NoAdaptation)
from .blocking import BlockingChequerboardSwitch
from .engine import Engine
from .path_saving_buffer import PathSavingBuffer
from .problems import chequerboard_layout, setup_time_grids
from .schedule import MCMCSchedule


class DiffusionMCMCBackend:
    """Dispatch tag (src/types.jl:34)."""


class PathImputation:
    """src/updates.jl:28-49: rho_s[block][interval], aux_laws, adpt."""

    def __init__(self, rho, aux_laws=None, adpt=None):
        if np.isscalar(rho):
            self.rho_s = [[float(rho)]]
        elif len(rho) and np.isscalar(rho[0]):
            self.rho_s = [list(map(float, rho))]
        else:
            self.rho_s = [list(map(float, r)) for r in rho]
        self.aux_laws = aux_laws
        adpt = adpt or NoAdaptation()
        if isinstance(adpt, AdaptationPathImputation):      # init_path_imp_adpt, adaptation.jl:155-158
            adpt = AdaptationMultiPathImputation(adpt)
        self.adpt = adpt

    def init_update(self, block_nintervals):
        """init_update! / update_rho! (src/updates.jl:53-74): one rho vector per block, filled
        with the first value."""
        while len(self.rho_s) < len(block_nintervals):
            self.rho_s.append([self.rho_s[0][0]])
        for r, n in zip(self.rho_s, block_nintervals):
            while len(r) < n:
                r.append(r[0])
        if isinstance(self.adpt, AdaptationMultiPathImputation):
            self.adpt.resize(len(block_nintervals))

    def rho_per_interval(self, blk_i1, blk_i2, J):
        rho = np.zeros(J)
        for b, (a, e) in enumerate(zip(blk_i1, blk_i2)):
            r = self.rho_s[b]
            rho[a:e + 1] = r[:e - a + 1]
        return rho


class RandomWalkUpdate:
    """Stand-in for eMCMC.RandomWalkUpdate(UniformRandomWalk(eps, positive), idx; prior)
    [UPSTREAM-RECALL]: theta deg_i = theta_i + U(-eps, eps), or theta_i * exp(U(-eps, eps)) when
    positive.  log_prior is flat unless given."""

    def __init__(self, eps, idx, positive=False, log_prior=None, critical=True):
        self.eps, self.idx, self.positive = float(eps), list(np.atleast_1d(idx)), positive
        self.log_prior = log_prior or (lambda th: 0.0)
        self.critical = critical

    def proposal(self, theta, rng):
        th = np.array(theta, float)
        u = rng.uniform(-self.eps, self.eps, len(self.idx))
        if self.positive:
            th[self.idx] = th[self.idx] * np.exp(u)
        else:
            th[self.idx] = th[self.idx] + u
        return th

    def log_transition_ratio(self, theta, theta_prop):
        """log q(theta | theta deg) - log q(theta deg | theta): Jacobian of the log-scale walk."""
        if self.positive:
            return float(np.sum(np.log(theta_prop[self.idx]) - np.log(theta[self.idx])))
        return 0.0


class DiffusionConjugGsnUpdate:
    """src/updates.jl:83-105: conjugate Gaussian update of the parameters `coords` (theta indices)
    with prior N(prior_mu, prior_Sigma).  The path integrals mu, W (compute_mu_and_W, :241-262) are
    reduced on the device (dmcmc_conjugate_sums); the posterior
        Sigma = (W + prior_Sigma^-1)^-1,  mean = Sigma (mu + prior_Sigma^-1 prior_mu)   (:270-273)
    and the draw stay on the host.  Chains are conditionally independent copies of the recording, so
    their integrals add (like the reference's loop over recordings, :232-246)."""

    def __init__(self, coords, prior_mu, prior_Sigma):
        self.coords = list(np.atleast_1d(coords))
        self.prior_mu = np.atleast_1d(np.asarray(prior_mu, float))
        self.prior_Sigma = np.atleast_2d(np.asarray(prior_Sigma, float))
        self.adpt = NoAdaptation()
        self.critical = True

    def posterior(self, engine, theta):
        idx = list(engine.conjugate_coords())
        if not set(self.coords) <= set(idx):
            raise ValueError(f"parameters {self.coords} are not conjugate for this law (conjugate: {idx})")
        mu, W = engine.conjugate_sums()
        mu, W = mu.sum(axis=0), W.sum(axis=0)
        S = [idx.index(c) for c in self.coords]
        F = [k for k in range(len(idx)) if k not in S]
        th = np.asarray(theta, float)
        mu_S = mu[S] - W[np.ix_(S, F)] @ th[[idx[k] for k in F]]  # fixed members of the family enter phi_c
        P0 = np.linalg.inv(self.prior_Sigma)
        Sigma = np.linalg.inv(W[np.ix_(S, S)] + P0)
        Sigma = 0.5 * (Sigma + Sigma.T)
        return Sigma @ (mu_S + P0 @ self.prior_mu), Sigma

    def proposal(self, engine, theta, rng):
        m, Sigma = self.posterior(engine, theta)
        th = np.array(theta, float)
        th[self.coords] = rng.multivariate_normal(m, Sigma)
        return th


class MCMC:
    """eMCMC.MCMC(updates_and_decorators; backend) (docs/src/get_started/overview.md:72-82)."""

    def __init__(self, updates_and_decorators, backend=None):
        self.updates_and_decorators = list(updates_and_decorators)
        self.backend = backend or DiffusionMCMCBackend()
        self.updates = [u for u in self.updates_and_decorators
                        if not isinstance(u, BlockingChequerboardSwitch)]


class DiffusionLocalWorkspace:
    """Per-update histories (src/workspaces.jl:226-257, :339-349): `acceptance_history[it][block]`,
    `ll_history[it][block]`, `ll_prop_history[it][block]` (with a chain axis in between when
    n_chains > 1)."""

    def __init__(self, update, layout, num_mcmc_steps, n_chains, kind):
        self.update, self.layout, self.kind = update, layout, kind
        nblk = len(layout[0]) if kind == "imputation" else 1
        shape = (num_mcmc_steps, n_chains, nblk)
        self._acc = np.zeros(shape, bool)
        self._ll = np.zeros(shape)
        self._llp = np.zeros(shape)
        self.n_chains = n_chains

    def _view(self, a):
        return a[:, 0, :] if self.n_chains == 1 else a

    @property
    def acceptance_history(self):
        return self._view(self._acc)

    @property
    def ll_history(self):
        return self._view(self._ll)

    @property
    def ll_prop_history(self):
        return self._view(self._llp)


class DiffusionGlobalWorkspace:
    """src/workspaces.jl:119-127: device state (engine = sub_ws_diff + sub_ws_diff°), data,
    parameter chain (`state_history`), `XX_buffer`."""

    def __init__(self, engine, data, XX_buffer, pnames):
        self.engine, self.data, self.XX_buffer, self.pnames = engine, data, XX_buffer, pnames
        self.state_history = []


def _layouts_for_updates(rec_nobs, updates_and_decorators):
    """extra_schedule_params (src/schedule.jl:15-75) with update_layout / update_layout_type!
    defined: each BlockingChequerboardSwitch moves to the next chequerboard pattern."""
    layouts, types = [], []
    cur = chequerboard_layout(rec_nobs, 0, 0)
    cur_type, n_switch = 1, 0
    for u in updates_and_decorators:
        if isinstance(u, BlockingChequerboardSwitch):
            cur = chequerboard_layout(rec_nobs, u.block_half_len, n_switch % 2)
            n_switch += 1
            cur_type = 1 + (n_switch % 2) + 1
        else:
            layouts.append(cur)
            types.append(tuple(cur_type for _ in rec_nobs))
    return layouts, tuple(types)


def run_(mcmc, num_mcmc_steps, data, theta_init=None, callbacks=(), dt=None, path_buffer_size=1,
         n_chains=None, device=0, seed=0, chain_offset=0, save_every=100, rng=None,
         exclude_updates=()):
    """`run!(mcmc, num_mcmc_steps, data, θinit, callbacks; dt, path_buffer_size, ...)`.

    `data` is a spec dict from problems.make_spec (the `AllObservations` stand-in).  When `dt`
    is given the imputation grids are rebuilt from the observation times
    (OBS.setup_time_grids, src/workspaces.jl:164-170; default there is 0.01).
    Returns (global_ws, local_wss)."""
    spec = dict(data)
    if dt is not None:
        N, t = setup_time_grids(spec.get("t0", 0.0), spec["obs_times"], dt)
        spec["N"], spec["t"] = N, t
    if theta_init is not None:
        spec["theta"] = np.asarray(theta_init, float)
    if n_chains is not None and spec["x0"].shape[0] != n_chains:
        spec["x0"] = np.broadcast_to(spec["x0"][:1], (n_chains,) + spec["x0"].shape[1:]).copy()
    rng = rng or np.random.default_rng(seed)
    eng = Engine(spec, device=device, seed=seed, chain_offset=chain_offset)
    eng.init_paths()  # init_paths! + init_ll! (src/workspaces.jl:80-92, :425-431)
    C, J = eng.C, eng.J

    layouts, layout_types = _layouts_for_updates(eng.rec_nobs, mcmc.updates_and_decorators)
    updates = mcmc.updates
    for u, lay in zip(updates, layouts):
        if isinstance(u, PathImputation):
            u.init_update([int(e - a + 1) for a, e in zip(lay[0], lay[1])])

    # glued time axis per recording (PathSavingBuffer ctor, src/path_saving_buffer.jl:33-44)
    times, j0 = [], 0
    for n in eng.rec_nobs:
        segs = [spec["t"][eng.pt_off[j]:eng.pt_off[j + 1]] for j in range(j0, j0 + int(n))]
        times.append(np.concatenate([s[:-1] for s in segs] + [segs[-1][-1:]]))
        j0 += int(n)
    buf = PathSavingBuffer(times, eng.d, path_buffer_size)
    gws = DiffusionGlobalWorkspace(eng, spec, buf, list(range(len(spec["theta"]))))
    lws = [DiffusionLocalWorkspace(u, lay, num_mcmc_steps, C,
                                   "imputation" if isinstance(u, PathImputation) else "param")
           for u, lay in zip(updates, layouts)]

    schedule = MCMCSchedule(num_mcmc_steps, len(updates), exclude_updates,
                            extra_info=dict(layout_types=layout_types))
    cur_layout = chequerboard_layout(eng.rec_nobs, 0, 0)
    theta = np.array(spec["theta"], float)
    all_imputation = all(isinstance(u, PathImputation) for u in updates) and not exclude_updates \
        and not callbacks

    def same(l1, l2):
        return all(np.array_equal(a, b) for a, b in zip(l1, l2))

    def save_paths():
        buf.save([eng.download_path(0, r) for r in range(eng.R)])

    def adapt(u, lay, mcmciter):
        if not isinstance(u.adpt, AdaptationMultiPathImputation):
            return False
        acc, prop = eng.accept_counts()
        u.adpt.register(acc, prop)
        if not u.adpt.time_to_update():
            return False
        for b, a in enumerate(u.adpt.v[:len(u.rho_s)]):   # readjust!, src/adaptation.jl:143-149
            new = a.readjust(u.rho_s[b][0], mcmciter)
            u.rho_s[b] = [new] * len(u.rho_s[b])
        return True

    lst = mcmc.updates_and_decorators
    has_switch = any(isinstance(u, BlockingChequerboardSwitch) for u in lst)
    # the device loop alternates the pattern with every sweep: that is the schedule only when a switch decorator stands
    # directly before every update (or there is no blocking at all)
    switch_before_each = (len(lst) == 2 * len(updates) and
                          all(isinstance(lst[2 * i], BlockingChequerboardSwitch) and lst[2 * i + 1] is updates[i]
                              for i in range(len(updates))) and
                          len({u.block_half_len for u in lst if isinstance(u, BlockingChequerboardSwitch)}) <= 1)
    if all_imputation and len(updates) in (1, 2) and (not has_switch or switch_before_each):
        # Smoothing fast path: the whole schedule is PathImputation sweeps, so the device loops
        # (dmcmc_run_sweeps) between the points where the host must look: adaptation and saving.
        half_len = 0
        for u in mcmc.updates_and_decorators:
            if isinstance(u, BlockingChequerboardSwitch):
                half_len = u.block_half_len
        nu = len(updates)
        stride = max(len(l[0]) for l in layouts)
        n_switch = sum(isinstance(u, BlockingChequerboardSwitch) for u in mcmc.updates_and_decorators)
        if half_len > 0 and n_switch % 2 == 1:
            # every switch decorator moves to the other pattern; with an odd number of switches per pass of the update
            # list the pattern an update sees alternates from one MCMC iteration to the next, so its histories must hold
            # the larger of the two patterns' block counts (the shifted pattern has one block more)
            stride = max(stride, max(len(chequerboard_layout(eng.rec_nobs, half_len, p)[0]) for p in (0, 1)))
            for w in lws:
                if w._acc.shape[2] < stride:
                    shape = (num_mcmc_steps, C, stride)
                    w._acc, w._ll, w._llp = np.zeros(shape, bool), np.full(shape, np.nan), np.full(shape, np.nan)
        k_adapt = min([u.adpt.v[0].adapt_every_k_steps for u in updates
                       if isinstance(u.adpt, AdaptationMultiPathImputation)] or [save_every])
        max_sw = max(1, min(k_adapt, save_every, num_mcmc_steps)) * nu
        # ordinary host arrays: the library stages through its own pinned rings (dmcmc_run_sweeps)
        rho_buf = np.empty((max_sw, J), np.float64)
        hist_buf = (np.empty((max_sw, C, stride), np.float64), np.empty((max_sw, C, stride), np.float64),
                    np.empty((max_sw, C, stride), np.uint8))
        it = 0
        while it < num_mcmc_steps:
            chunk = min(num_mcmc_steps - it, max(1, min(k_adapt, save_every - (it % save_every))))
            nsw = chunk * nu
            rho_sched = rho_buf[:nsw]
            for s in range(nsw):
                u, lay = updates[s % nu], layouts[s % nu]
                rho_sched[s] = u.rho_per_interval(lay[0], lay[1], J)
            hist = tuple(h[:nsw] for h in hist_buf)
            if half_len == 0 and not same(cur_layout, layouts[0]):
                eng.relayout(*layouts[0])
            eng.run_sweeps(it * nu, nsw, half_len, 0, rho_sched, hist)
            cur_layout = layouts[(nsw - 1) % nu]
            eng.nblk = len(cur_layout[0])
            for s in range(nsw):
                w = lws[s % nu]
                nb = w._acc.shape[2]
                w._llp[it + s // nu] = hist[0][s][:, :nb]
                w._ll[it + s // nu] = hist[1][s][:, :nb]
                w._acc[it + s // nu] = hist[2][s][:, :nb].astype(bool)
            it += chunk
            for u, lay in zip(updates, layouts):
                adapt(u, lay, it)
            if it % save_every == 0:
                save_paths()
        gws.state_history = [theta.copy() for _ in range(num_mcmc_steps)]
        return gws, lws

    # General path: the schedule is walked update by update (src/run!_alterations.jl:5-26, :357-369)
    for step in schedule:
        pidx, mcmciter = step.pidx - 1, step.mcmciter
        u, lay, w = updates[pidx], layouts[pidx], lws[pidx]
        if not same(cur_layout, lay):   # update_workspaces!: layout changed (src/workspaces.jl:467-471)
            eng.relayout(*lay)
            cur_layout = lay
        if isinstance(u, PathImputation):
            eng.set_rho(u.rho_per_interval(lay[0], lay[1], J))
            llp, ll, acc = eng.impute(mcmciter * len(updates) + pidx)
            w._llp[mcmciter - 1], w._ll[mcmciter - 1], w._acc[mcmciter - 1] = llp, ll, acc
            adapt(u, lay, mcmciter)
            if mcmciter % save_every == 0 and step.pidx == 1:   # save_path!, src/run!_alterations.jl:51-54
                save_paths()
        elif isinstance(u, DiffusionConjugGsnUpdate):
            th_prop = u.proposal(eng, theta, rng)                              # proposal!: conjugate_draw, src/updates.jl:225-246
            llp, ok = eng.param_loglik(th_prop, critical=True)                 # set_proposal! + compute_ll! (fixed W)
            accepted = bool(ok.all())                                          # conjugate draws are always accepted
            eng.param_commit(accepted)
            if accepted:
                theta = th_prop
            w._llp[mcmciter - 1, :, 0] = llp.sum(axis=1)
            w._ll[mcmciter - 1, :, 0] = llp.sum(axis=1) if accepted else eng.get_ll().sum(axis=1)
            w._acc[mcmciter - 1, :, 0] = accepted
        else:
            th_prop = u.proposal(theta, rng)                                   # proposal!
            llp, ok = eng.param_loglik(th_prop, critical=u.critical)           # set_proposal!
            ll = eng.get_ll()
            # one decision for all blocks and chains (src/run!_alterations.jl:367): E > -(Delta)
            delta = (llp.sum() - ll.sum() + u.log_prior(th_prop) - u.log_prior(theta)
                     + u.log_transition_ratio(theta, th_prop))
            accepted = bool(ok.all()) and (rng.exponential() > -delta)
            eng.param_commit(accepted)                                         # swap_laws_and_paths!
            if accepted:
                theta = th_prop
            w._llp[mcmciter - 1, :, 0] = llp.sum(axis=1)
            w._ll[mcmciter - 1, :, 0] = (llp if accepted else ll).sum(axis=1)
            w._acc[mcmciter - 1, :, 0] = accepted
        if step.pidx == len(updates) or True:
            if len(gws.state_history) < mcmciter:
                gws.state_history.append(theta.copy())
            else:
                gws.state_history[mcmciter - 1] = theta.copy()
        for cb in callbacks:
            cb(gws, lws, step)
    return gws, lws


class ImputationRunner:
    """What bench.py times: the sweeps of a smoothing chain under chequerboard blocking, either
    device-resident (no host traffic) or through the host buffers `run_` uses."""

    def __init__(self, engine, half_len=1, rho=0.96):
        self.e, self.half_len = engine, half_len
        self.lays = [chequerboard_layout(engine.rec_nobs, half_len, p) for p in (0, 1)]
        self.stride = max(len(l[0]) for l in self.lays)
        self.rho = rho
        self._host = None
        self.sweeps = 0

    def run_device(self, it0, n):
        self.e.run_sweeps(it0, n, self.half_len, self.sweeps & 1)
        self.sweeps += n

    def reserve_host(self, n):
        """Host buffers (ordinary NumPy arrays, as a Julia caller would pass) for n sweeps of rho
        schedule and history rows, allocated and touched once."""
        e = self.e
        if self._host is None or self._host[0].shape[0] < n:
            self._host = (np.empty((n, e.J), np.float64), np.empty((n, e.C, self.stride), np.float64),
                          np.empty((n, e.C, self.stride), np.float64),
                          np.empty((n, e.C, self.stride), np.uint8))
            for a in self._host:
                a[...] = 0  # touch the pages

    def run_host(self, it0, n):
        e = self.e
        self.reserve_host(n)
        rho, llp, ll, acc = (a[:n] for a in self._host)
        rho[...] = self.rho
        e.run_sweeps(it0, n, self.half_len, self.sweeps & 1, rho, (llp, ll, acc))
        self.sweeps += n
        h2d = e.J * 8
        nb = np.mean([len(l[0]) for l in self.lays]) if self.half_len > 0 else e.nblk
        d2h = int(e.C * nb * (8 + 8 + 1))
        return h2d, d2h

    def kernel_breakdown(self):
        ms, n = self.e.kernel_times()
        out = {"launches": int(n.sum())}
        for i, name in enumerate(["impute", "param", "relayout", "loglik"]):
            if n[i]:
                out[name + "_ms"] = float(ms[i] / n[i])
                out[name + "_launches"] = int(n[i])
        return out
