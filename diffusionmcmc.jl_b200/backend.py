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
import copy

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

    def __init__(self, update, layout, num_mcmc_steps, n_chains, kind, nblk=None):
        self.update, self.layout, self.kind = update, layout, kind
        if nblk is None:
            nblk = len(layout[0]) if kind == "imputation" else 1
        shape = (num_mcmc_steps, n_chains, nblk)   # nblk: the most blocks any pattern of this update has
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


class _BlockingPlan:
    """extra_schedule_params (src/schedule.jl:15-75) with `update_layout` / `update_layout_type!` defined (the
    reference calls them but defines neither): every BlockingChequerboardSwitch in the update list moves to the other
    chequerboard pattern, the first one to pattern 0.  The reference asserts an even number of switches per pass
    (src/schedule.jl:89), so that an update always sees the same layout; here an odd number is allowed as well, and
    then the switch count -- hence the pattern an update runs under -- carries over from one MCMC iteration to the
    next (`[Switch, PathImputation]` alternates the two patterns, which is what makes the sampler move the block
    end points)."""

    def __init__(self, rec_nobs, updates_and_decorators):
        self.rec_nobs = rec_nobs
        self.half_len, self.k_before = [], []   # per update: block_half_len in force (0: none), switches passed so far
        h, k = 0, 0
        for u in updates_and_decorators:
            if isinstance(u, BlockingChequerboardSwitch):
                h, k = int(u.block_half_len), k + 1
            else:
                self.half_len.append(h)
                self.k_before.append(k)
        self.n_switch = k
        self._cache = {}

    def pattern(self, ui, mcmciter):
        """chequerboard pattern (0 / 1) of update `ui` (0-based) in MCMC iteration `mcmciter` (1-based); None = no blocking"""
        if self.half_len[ui] <= 0 or self.k_before[ui] == 0:
            return None
        return ((mcmciter - 1) * self.n_switch + self.k_before[ui] - 1) % 2

    def patterns_seen(self, ui):
        if self.pattern(ui, 1) is None:
            return [None]
        return sorted({self.pattern(ui, 1), self.pattern(ui, 2)})

    def layout(self, ui, pat):
        key = (self.half_len[ui], pat)
        if key not in self._cache:
            self._cache[key] = chequerboard_layout(self.rec_nobs, 0, 0) if pat is None else \
                chequerboard_layout(self.rec_nobs, self.half_len[ui], pat)
        return self._cache[key]

    def layout_types(self):
        """layout type per update and recording in the first pass (src/schedule.jl:62-64), for MCMCSchedule.same_layout"""
        return tuple(tuple(1 if self.pattern(ui, 1) is None else 2 + self.pattern(ui, 1) for _ in self.rec_nobs)
                     for ui in range(len(self.half_len)))


class _ImputationState:
    """pCN memory and its adaptation for ONE (update, chequerboard pattern): rho_s[block][interval] and one
    AdaptationPathImputation per block (src/updates.jl:53-74, src/adaptation.jl:122-167).  An update that alternates
    between the two patterns keeps two of these: block b of pattern 0 and block b of pattern 1 are different blocks."""

    def __init__(self, update, layout):
        self.layout = layout
        n_iv = [int(e - a + 1) for a, e in zip(layout[0], layout[1])]
        first = update.rho_s[0][0]
        self.rho_s = [list(r) for r in update.rho_s[:len(n_iv)]]
        while len(self.rho_s) < len(n_iv):
            self.rho_s.append([first])
        for r, n in zip(self.rho_s, n_iv):
            while len(r) < n:
                r.append(r[0])
            del r[n:]
        self.adpt = copy.deepcopy(update.adpt)
        if isinstance(self.adpt, AdaptationMultiPathImputation):
            self.adpt.resize(len(n_iv))

    def rho_per_interval(self, J):
        rho = np.zeros(J)
        for r, a, e in zip(self.rho_s, self.layout[0], self.layout[1]):
            rho[a:e + 1] = r[:e - a + 1]
        return rho

    def register_and_adapt(self, acc_counts, proposed, mcmciter):
        """register! + time_to_update + readjust! (src/adaptation.jl:82-90, :100-113, :143-149)."""
        if not isinstance(self.adpt, AdaptationMultiPathImputation):
            return False
        self.adpt.register(acc_counts, proposed)
        if not self.adpt.time_to_update():
            return False
        for b, a in enumerate(self.adpt.v[:len(self.rho_s)]):
            new = a.readjust(self.rho_s[b][0], mcmciter)
            self.rho_s[b] = [new] * len(self.rho_s[b])
        return True


def run_(mcmc, num_mcmc_steps, data, theta_init=None, callbacks=(), dt=None, path_buffer_size=1,
         n_chains=None, device=0, seed=0, chain_offset=0, save_every=100, rng=None,
         exclude_updates=(), save_chains=(0,), device_adaptation=True):
    """`run!(mcmc, num_mcmc_steps, data, θinit, callbacks; dt, path_buffer_size, ...)`.

    `data` is a spec dict from problems.make_spec (the `AllObservations` stand-in).  When `dt`
    is given the imputation grids are rebuilt from the observation times
    (OBS.setup_time_grids, src/workspaces.jl:164-170; default there is 0.01).
    `save_every` replaces the reference's hard-coded 100 (src/run!_alterations.jl:52); `save_chains` are the chains
    whose glued accepted paths go into `XX_buffer` (one: `x` is [points][d] like the reference's; several:
    [chain][points][d]).  `device_adaptation`: in a pure-imputation schedule the pCN memory is adapted on the device
    (dmcmc_set_adaptation), so the sweep loop only returns to the host to save paths.  Returns (global_ws, local_wss)."""
    spec = dict(data)
    if dt is not None:
        N, t = setup_time_grids(spec.get("t0", 0.0), spec["obs_times"], dt)
        spec["N"], spec["t"] = N, t
    if theta_init is not None:
        spec["theta"] = np.asarray(theta_init, float)
    if n_chains is not None and spec["x0"].shape[0] != n_chains:
        spec["x0"] = np.broadcast_to(spec["x0"][:1], (n_chains,) + spec["x0"].shape[1:]).copy()
    updates = mcmc.updates
    if spec["x0"].shape[0] > 1 and any(not isinstance(u, PathImputation) for u in updates):
        # every chain of a context shares ONE theta (the tables are shared): a joint Metropolis step over C chains
        # would target prior x likelihood^C.  Chains are for smoothing; for inference run one chain per context.
        raise ValueError("parameter updates need n_chains == 1: the chains of one context share the parameter, "
                         "so a joint accept/reject over them would count the recording n_chains times")
    rng = rng or np.random.default_rng(seed)
    eng = Engine(spec, device=device, seed=seed, chain_offset=chain_offset)
    eng.init_paths()  # init_paths! + init_ll! (src/workspaces.jl:80-92, :425-431)
    C, J = eng.C, eng.J
    nu = len(updates)

    plan = _BlockingPlan(eng.rec_nobs, mcmc.updates_and_decorators)
    states = {}   # (update index, pattern) -> _ImputationState
    for ui, u in enumerate(updates):
        if isinstance(u, PathImputation):
            for pat in plan.patterns_seen(ui):
                states[(ui, pat)] = _ImputationState(u, plan.layout(ui, pat))

    # glued time axis per recording (PathSavingBuffer ctor, src/path_saving_buffer.jl:33-44)
    times, j0 = [], 0
    for n in eng.rec_nobs:
        segs = [spec["t"][eng.pt_off[j]:eng.pt_off[j + 1]] for j in range(j0, j0 + int(n))]
        times.append(np.concatenate([s[:-1] for s in segs] + [segs[-1][-1:]]))
        j0 += int(n)
    save_chains = [int(c) for c in save_chains]
    buf = PathSavingBuffer(times, eng.d, path_buffer_size, n_chains=len(save_chains))
    gws = DiffusionGlobalWorkspace(eng, spec, buf, list(range(len(spec["theta"]))))
    lws = []
    for ui, u in enumerate(updates):
        if isinstance(u, PathImputation):
            nb = max(len(plan.layout(ui, pat)[0]) for pat in plan.patterns_seen(ui))
            lws.append(DiffusionLocalWorkspace(u, plan.layout(ui, plan.pattern(ui, 1)), num_mcmc_steps, C, "imputation", nb))
        else:
            lws.append(DiffusionLocalWorkspace(u, plan.layout(ui, plan.pattern(ui, 1)), num_mcmc_steps, C, "param", 1))

    schedule = MCMCSchedule(num_mcmc_steps, nu, exclude_updates, extra_info=dict(layout_types=plan.layout_types()))
    cur_layout = chequerboard_layout(eng.rec_nobs, 0, 0)
    theta = np.array(spec["theta"], float)
    all_imputation = all(isinstance(u, PathImputation) for u in updates) and not exclude_updates \
        and not callbacks

    def same(l1, l2):
        return all(np.array_equal(a, b) for a, b in zip(l1, l2))

    pending_save = []

    def save_paths():
        """save! (src/path_saving_buffer.jl:52): the snapshot is taken on the device now, the copy to the host overlaps
        the sweeps that follow and is collected at the next save (or at the end)."""
        collect_saves()
        pending_save.append([eng.save_path_async(save_chains, r) for r in range(eng.R)])

    def collect_saves():
        while pending_save:
            glued = [eng.save_path_wait(tk) for tk in pending_save.pop(0)]
            buf.save([g[0] if len(save_chains) == 1 else g for g in glued])

    def finish():
        collect_saves()
        for (ui, pat), st in states.items():      # report the adapted rho back on the update object
            if pat == plan.pattern(ui, max(1, num_mcmc_steps)):
                updates[ui].rho_s = [list(r) for r in st.rho_s]
                updates[ui].adpt = st.adpt
        return gws, lws

    lst = mcmc.updates_and_decorators
    has_switch = plan.n_switch > 0
    # the device loop alternates the pattern with every sweep: that is the schedule only when a switch decorator stands
    # directly before every update (or there is no blocking at all)
    switch_before_each = (len(lst) == 2 * nu and
                          all(isinstance(lst[2 * i], BlockingChequerboardSwitch) and lst[2 * i + 1] is updates[i]
                              for i in range(nu)) and
                          len({u.block_half_len for u in lst if isinstance(u, BlockingChequerboardSwitch)}) <= 1)
    if all_imputation and nu in (1, 2) and (not has_switch or switch_before_each):
        # Smoothing fast path: the whole schedule is PathImputation sweeps, so the device loops
        # (dmcmc_run_sweeps) between the points where the host must look: adaptation and saving.
        half_len = plan.half_len[0] if has_switch else 0
        stride = max(w._acc.shape[2] for w in lws)
        adaptive = [st for st in states.values() if isinstance(st.adpt, AdaptationMultiPathImputation)]
        k_adapt = min([st.adpt.v[0].adapt_every_k_steps for st in adaptive] or [save_every])
        # device-side adaptation (dmcmc_set_adaptation): one rho vector and one set of counters per chequerboard pattern,
        # so it applies when every (update, pattern) state has its own pattern slot and all share one configuration
        slot = {key: (key[1] if key[1] is not None else 0) for key in states}
        cfgs = {tuple(getattr(st.adpt.v[0], f) for f in ("target_accpt_rate", "adapt_every_k_steps", "scale", "min", "max", "offset"))
                for st in adaptive}
        dev_adapt = (device_adaptation and adaptive and len(adaptive) == len(states) and len(cfgs) == 1
                     and len(set(slot.values())) == len(slot))
        by_slot = {slot[key]: st for key, st in states.items()}
        if dev_adapt:
            eng.set_adaptation(adaptive[0].adpt.v[0], nu, by_slot[0].rho_per_interval(J),
                               by_slot.get(1, by_slot[0]).rho_per_interval(J))
            k_adapt = save_every            # no reason to come back to the host for the adaptation
        row_bytes = C * stride * 17 * nu    # history bytes per MCMC iteration
        max_it = max(1, min(k_adapt, save_every, num_mcmc_steps, (256 << 20) // max(1, row_bytes)))
        max_sw = max_it * nu
        # pinned host arrays (dmcmc_host_alloc): the history rows land in them by DMA, no staging copy
        rho_buf = eng.pinned((max_sw, J), np.float64)
        hist_buf = (eng.pinned((max_sw, C, stride), np.float64), eng.pinned((max_sw, C, stride), np.float64),
                    eng.pinned((max_sw, C, stride), np.uint8))
        it = 0
        while it < num_mcmc_steps:
            chunk = min(num_mcmc_steps - it, max_it, max(1, min(k_adapt, save_every - (it % save_every))))
            nsw = chunk * nu
            sweep_state = []
            for s in range(nsw):
                ui, mcmciter = s % nu, it + s // nu + 1
                sweep_state.append(states[(ui, plan.pattern(ui, mcmciter))])
                if not dev_adapt:
                    rho_buf[s] = sweep_state[-1].rho_per_interval(J)
            hist = tuple(h[:nsw] for h in hist_buf)
            if half_len == 0 and not same(cur_layout, sweep_state[0].layout):
                eng.relayout(*sweep_state[0].layout)
            first_pattern = plan.pattern(0, it + 1) or 0
            eng.run_sweeps(it * nu, nsw, half_len, first_pattern, None if dev_adapt else rho_buf[:nsw], hist)
            cur_layout = sweep_state[-1].layout
            eng.nblk = len(cur_layout[0])
            for s, st in enumerate(sweep_state):
                w, row = lws[s % nu], it + s // nu
                nb = len(st.layout[0])
                w._llp[row, :, :nb] = hist[0][s][:, :nb]
                w._ll[row, :, :nb] = hist[1][s][:, :nb]
                w._acc[row, :, :nb] = hist[2][s][:, :nb].astype(bool)
                w._llp[row, :, nb:] = np.nan     # blocks this sweep's pattern does not have
                w._ll[row, :, nb:] = np.nan
                if not dev_adapt:
                    st.register_and_adapt(hist[2][s][:, :nb].sum(axis=0, dtype=np.int64), np.full(nb, C), row + 1)
            it += chunk
            if it % save_every == 0:
                save_paths()
        if dev_adapt:   # bring the adapted rho back into the host-side states
            for sl, st in by_slot.items():
                rho = eng.get_rho(sl)
                st.rho_s = [[float(rho[a])] * int(e - a + 1) for a, e in zip(st.layout[0], st.layout[1])]
            eng.set_adaptation(None)
        gws.state_history = [theta.copy() for _ in range(num_mcmc_steps)]
        return finish()

    # General path: the schedule is walked update by update (src/run!_alterations.jl:5-26, :357-369)
    for step in schedule:
        pidx, mcmciter = step.pidx - 1, step.mcmciter
        u, w = updates[pidx], lws[pidx]
        pat = plan.pattern(pidx, mcmciter)
        lay = plan.layout(pidx, pat)
        if not same(cur_layout, lay):   # update_workspaces!: layout changed (src/workspaces.jl:467-471)
            eng.relayout(*lay)
            cur_layout = lay
        if isinstance(u, PathImputation):
            st = states[(pidx, pat)]
            nb = len(lay[0])
            eng.set_rho(st.rho_per_interval(J))
            llp, ll, acc = eng.impute(mcmciter * nu + pidx)
            w._llp[mcmciter - 1, :, :nb], w._ll[mcmciter - 1, :, :nb], w._acc[mcmciter - 1, :, :nb] = llp, ll, acc
            w._llp[mcmciter - 1, :, nb:] = np.nan
            w._ll[mcmciter - 1, :, nb:] = np.nan
            st.register_and_adapt(acc.sum(axis=0, dtype=np.int64), np.full(nb, C), mcmciter)
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
            # one decision for all blocks (src/run!_alterations.jl:367): E > -(Delta)
            delta = (llp.sum() - ll.sum() + u.log_prior(th_prop) - u.log_prior(theta)
                     + u.log_transition_ratio(theta, th_prop))
            accepted = bool(ok.all()) and (rng.exponential() > -delta)
            eng.param_commit(accepted)                                         # swap_laws_and_paths!
            if accepted:
                theta = th_prop
            w._llp[mcmciter - 1, :, 0] = llp.sum(axis=1)
            w._ll[mcmciter - 1, :, 0] = (llp if accepted else ll).sum(axis=1)
            w._acc[mcmciter - 1, :, 0] = accepted
        if len(gws.state_history) < mcmciter:
            gws.state_history.append(theta.copy())
        else:
            gws.state_history[mcmciter - 1] = theta.copy()
        for cb in callbacks:
            cb(gws, lws, step)
    return finish()


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
        """Pinned host buffers (dmcmc_host_alloc; a Julia caller would wrap them as Arrays, or register its own with
        dmcmc_host_register) for n sweeps of rho schedule and history rows, allocated and touched once."""
        e = self.e
        if self._host is None or self._host[0].shape[0] < n:
            self._host = (e.pinned((n, e.J), np.float64), e.pinned((n, e.C, self.stride), np.float64),
                          e.pinned((n, e.C, self.stride), np.float64),
                          e.pinned((n, e.C, self.stride), np.uint8))
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
