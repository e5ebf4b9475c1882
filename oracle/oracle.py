"""ctypes front-end of the CPU oracle (TEST INFRASTRUCTURE, NOT PRODUCT CODE).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs import
this module.  See oracle/dmcmc_oracle.h for the parity status ("parity unpinned") and the
reference file:line each C function restates.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "liboracle.so")
_SRC = os.path.join(_HERE, "dmcmc_oracle.c")

MODEL_FHN, MODEL_LV, MODEL_L63, MODEL_L96, MODEL_LIN2, MODEL_LVM = range(6)


def build(force=False):
    """gcc -O2 -ffp-contract=off: the source's evaluation order is the arithmetic."""
    hdr = os.path.join(_HERE, "dmcmc_oracle.h")
    if (not force and os.path.exists(_LIB_PATH)
            and os.path.getmtime(_LIB_PATH) >= max(os.path.getmtime(_SRC), os.path.getmtime(hdr))):
        return _LIB_PATH
    cmd = ["gcc", "-O2", "-ffp-contract=off", "-mfma", "-fopenmp", "-shared", "-fPIC", "-Wall",
           "-o", _LIB_PATH, _SRC, "-lm"]
    subprocess.check_call(cmd)
    return _LIB_PATH


_dp = C.POINTER(C.c_double)
_ip = C.POINTER(C.c_int32)
_bp = C.POINTER(C.c_uint8)


class _Problem(C.Structure):
    _fields_ = [("model", C.c_int32), ("d", C.c_int32), ("m", C.c_int32), ("nth", C.c_int32),
                ("th", _dp), ("J", C.c_int32), ("R", C.c_int32), ("N", _ip), ("pt_off", _ip),
                ("st_off", _ip), ("t", _dp), ("auxB", _dp), ("auxbeta", _dp), ("auxsig", _dp),
                ("nblk", C.c_int32), ("blk_i1", _ip), ("blk_i2", _ip), ("blk_pinned", _ip),
                ("H", _dp), ("F0", _dp), ("Psi", _dp), ("c0", _dp), ("blk_g", _dp),
                ("blk_Qc", _dp), ("rho", _dp), ("C", C.c_int32), ("chain_offset", C.c_int32),
                ("interval_offset", C.c_int32)]


class _State(C.Structure):
    _fields_ = [("X", _dp * 2), ("W", _dp * 2), ("selX", _bp), ("selW", _bp), ("ll", _dp)]


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        _lib = C.CDLL(_LIB_PATH)
        _lib.orc_philox_exponential.restype = C.c_double
        _lib.orc_loglik_obs.restype = C.c_double
        _lib.orc_loglikhd_block.restype = C.c_double
    return _lib


def _d(a):
    return a.ctypes.data_as(_dp)


def _i(a):
    return a.ctypes.data_as(_ip)


def _b(a):
    return a.ctypes.data_as(_bp)


def f64(a):
    return np.ascontiguousarray(a, dtype=np.float64)


def i32(a):
    return np.ascontiguousarray(a, dtype=np.int32)


# ---------------------------------------------------------------------------------------------
# thin function wrappers
# ---------------------------------------------------------------------------------------------
def model_dims(model, d_in=0):
    d, m, nth = C.c_int(), C.c_int(), C.c_int()
    rc = lib().orc_model_dims(model, d_in, C.byref(d), C.byref(m), C.byref(nth))
    if rc:
        raise ValueError("bad model/dimension")
    return d.value, m.value, nth.value


def drift(model, d, th, x, t=0.0):
    th, x = f64(th), f64(x)
    b = np.zeros(d)
    lib().orc_drift(model, d, _d(th), C.c_double(t), _d(x), _d(b))
    return b


def sigma(model, d, m, th, x):
    th, x = f64(th), f64(x)
    s = np.zeros((d, m))
    lib().orc_sigma(model, d, m, _d(th), _d(x), _d(s))
    return s


def aux_build(model, d, m, th, v):
    th, v = f64(th), f64(np.atleast_1d(v))
    B, beta, sig = np.zeros((d, d)), np.zeros(d), np.zeros((d, m))
    lib().orc_aux_build(model, d, m, _d(th), _d(v), len(v), _d(B), _d(beta), _d(sig))
    return B, beta, sig


def obs_terms(L, Sigma, v):
    L, Sigma, v = f64(np.atleast_2d(L)), f64(np.atleast_2d(Sigma)), f64(np.atleast_1d(v))
    p, d = L.shape
    H, F, c = np.zeros((d, d)), np.zeros(d), C.c_double()
    rc = lib().orc_obs_terms(d, p, _d(L), _d(Sigma), _d(v), _d(H), _d(F), C.byref(c))
    if rc:
        raise ValueError("singular observation covariance")
    return H, F, c.value


def transition(B, beta, atil, dt):
    B, beta, atil = f64(B), f64(beta), f64(atil)
    d = B.shape[0]
    Phi, Q, m = np.zeros((d, d)), np.zeros((d, d)), np.zeros(d)
    lib().orc_transition(d, _d(B), _d(beta), _d(atil), C.c_double(dt), _d(Phi), _d(Q), _d(m))
    return Phi, Q, m


def philox4x32_10(ctr, key):
    ctr = np.ascontiguousarray(ctr, dtype=np.uint32)
    key = np.ascontiguousarray(key, dtype=np.uint32)
    out = np.zeros(4, dtype=np.uint32)
    lib().orc_philox4x32_10(ctr.ctypes.data_as(C.c_void_p), key.ctypes.data_as(C.c_void_p),
                            out.ctypes.data_as(C.c_void_p))
    return out


def philox_normal4(seed, it, interval, chain, wcomp, quad):
    z = np.zeros(4)
    lib().orc_philox_normal4(C.c_uint64(seed), C.c_uint32(it), C.c_uint32(interval),
                             C.c_uint32(chain), C.c_uint32(wcomp), C.c_uint32(quad), _d(z))
    return z


def philox_exponential(seed, it, first_interval, chain):
    return lib().orc_philox_exponential(C.c_uint64(seed), C.c_uint32(it),
                                        C.c_uint32(first_interval), C.c_uint32(chain))


def conjugate_structure(model, th, x):
    """(theta indices, rough-coordinate rows, phi [NR][NC], phi_c [NR], Gamma^-1 diagonal [NR]) of a law
    whose rough coordinates are linear in the parameters, b_rough = phi theta + phi_c
    (phi / phi_c of /root/reference src/updates.jl:296-330 written out per law)."""
    if model == MODEL_FHN:      # dX2 = (gamma X1 - X2 + beta) dt + sigma dW
        return [2, 3], [1], np.array([[x[0], 1.0]]), np.array([-x[1]]), np.array([1.0 / th[4] ** 2])
    if model in (MODEL_LV, MODEL_LVM):
        P = np.array([[x[0], -x[0] * x[1], 0.0, 0.0], [0.0, 0.0, -x[1], x[0] * x[1]]])
        s = np.array([th[4] * x[0], th[5] * x[1]]) if model == MODEL_LVM else np.array([th[4], th[5]])
        return [0, 1, 2, 3], [0, 1], P, np.zeros(2), 1.0 / s ** 2
    if model == MODEL_L63:
        P = np.diag([x[1] - x[0], x[0], -x[2]])
        return [0, 1, 2], [0, 1, 2], P, np.array([0.0, -x[1] - x[0] * x[2], x[0] * x[1]]), np.full(3, 1.0 / th[3] ** 2)
    raise ValueError("no conjugate structure for this law")


def conjugate_sums(model, th, X, N, t):
    """compute_mu_and_W (/root/reference src/updates.jl:241-262) over per-interval containers
    X [C][P][d] (N_j + 1 points each, as Oracle.accepted_X returns them) with grid times t [P]:
        mu += phi' Gamma^-1 dy - phi' Gamma^-1 phi_c dt ;  W += phi' Gamma^-1 phi dt   per Euler step.
    Pure-Python loops: small cases only."""
    C_ = X.shape[0]
    idx = conjugate_structure(model, th, X[0, 0])[0]
    nc = len(idx)
    mu, W = np.zeros((C_, nc)), np.zeros((C_, nc, nc))
    off = 0
    for n in N:
        for c in range(C_):
            for i in range(int(n)):
                x, xn = X[c, off + i], X[c, off + i + 1]
                dt = t[off + i + 1] - t[off + i]
                _, rows, P, pc, g = conjugate_structure(model, th, x)
                dy = xn[rows] - x[rows]
                G = np.diag(g)
                mu[c] += P.T @ G @ dy - P.T @ G @ pc * dt
                W[c] += P.T @ G @ P * dt
        off += int(n) + 1
    return idx, mu, W


def chequerboard_layout(rec_nobs, half_len, pattern):
    rec_nobs = i32(rec_nobs)
    J = int(rec_nobs.sum())
    i1, i2, pin = np.zeros(J, np.int32), np.zeros(J, np.int32), np.zeros(J, np.int32)
    nb = lib().orc_chequerboard_layout(len(rec_nobs), _i(rec_nobs), half_len, pattern,
                                       _i(i1), _i(i2), _i(pin))
    return i1[:nb].copy(), i2[:nb].copy(), pin[:nb].copy()


# ---------------------------------------------------------------------------------------------
# the oracle sampler state: accepted / proposal containers + selectors, reference-like AoS
# ---------------------------------------------------------------------------------------------
class Oracle:
    """CPU restatement of the imputation path for one theta-instance and C chains.

    spec fields used (all numpy): model, d, m, theta, N[J], t[P], rec_first[J] (1 when the
    interval is the first of its recording), obs_L[J][p][d], obs_Sigma[J][p][p], obs_v[J][p],
    x0[C][R][d], rho[J], pin_eps.
    """

    def __init__(self, spec, nthreads=1, chain_offset=0, interval_offset=0, bf_scheme=None):
        self.nthreads = nthreads
        # backward-filter scheme (both exact flows of the H, F, c ODEs, see orc_backward_filter_scheme): by default
        # the one the device kernel of that state dimension uses -- 1 (every grid point from the interval's right
        # end) for d <= 3, 0 (step by step) for Lorenz-96
        self.bf_scheme = bf_scheme if bf_scheme is not None else (1 if int(spec["d"]) <= 3 else 0)
        self.chain_offset, self.interval_offset = chain_offset, interval_offset
        self.model, self.d, self.m = int(spec["model"]), int(spec["d"]), int(spec["m"])
        self.N = i32(spec["N"])
        self.J = len(self.N)
        self.pt_off = i32(np.concatenate([[0], np.cumsum(self.N + 1)]))
        self.st_off = i32(np.concatenate([[0], np.cumsum(self.N)]))
        self.P, self.S = int(self.pt_off[-1]), int(self.st_off[-1])
        self.t = f64(spec["t"])
        assert self.t.shape == (self.P,)
        self.rec_first = i32(spec["rec_first"])
        self.rec_nobs = i32(np.diff(np.concatenate([np.flatnonzero(self.rec_first), [self.J]])))
        self.R = len(self.rec_nobs)
        self.obs_L, self.obs_Sigma, self.obs_v = f64(spec["obs_L"]), f64(spec["obs_Sigma"]), f64(spec["obs_v"])
        self.pin_eps = float(spec.get("pin_eps", 1e-4))
        self.x0 = f64(spec["x0"])
        self.C = self.x0.shape[0]
        self.rho = f64(np.broadcast_to(spec["rho"], (self.J,)).copy())
        d, m, Cn = self.d, self.m, self.C
        self.X = [np.zeros((Cn, self.P, d)), np.zeros((Cn, self.P, d))]
        self.W = [np.zeros((Cn, self.P, m)), np.zeros((Cn, self.P, m))]
        self.selX = np.zeros((Cn, self.J), np.uint8)
        self.selW = np.zeros((Cn, self.J), np.uint8)
        # starting points of the recordings live in the accepted containers (src/workspaces.jl:86-90)
        firsts = np.flatnonzero(self.rec_first)
        for r, j in enumerate(firsts):
            self.X[0][:, self.pt_off[j], :] = self.x0[:, r, :]
            self.X[1][:, self.pt_off[j], :] = self.x0[:, r, :]
        self.set_theta(spec["theta"])
        self.set_layout(*chequerboard_layout(self.rec_nobs, 0, 0))

    # -- law ------------------------------------------------------------------------------
    def set_theta(self, theta, custom_aux=None):
        self.theta = f64(theta)
        self.__dict__.pop("_bf_cache", None)
        d, m, J = self.d, self.m, self.J
        self.auxB, self.auxbeta, self.auxsig = np.zeros((J, d, d)), np.zeros((J, d)), np.zeros((J, d, m))
        for j in range(J):
            if custom_aux is not None:
                self.auxB[j], self.auxbeta[j], self.auxsig[j] = custom_aux[j]
            else:
                self.auxB[j], self.auxbeta[j], self.auxsig[j] = aux_build(
                    self.model, d, m, self.theta, self.obs_v[j])
        self.auxatil = f64(np.einsum("jik,jlk->jil", self.auxsig, self.auxsig))
        self.Hobs, self.Fobs, self.cobs = np.zeros((J, d, d)), np.zeros((J, d)), np.zeros(J)
        for j in range(J):
            self.Hobs[j], self.Fobs[j], self.cobs[j] = obs_terms(
                self.obs_L[j], self.obs_Sigma[j], self.obs_v[j])
        if hasattr(self, "blk_i1"):
            self.backward_filter()

    def set_layout(self, blk_i1, blk_i2, blk_pinned):
        self.blk_i1, self.blk_i2, self.blk_pinned = i32(blk_i1), i32(blk_i2), i32(blk_pinned)
        self.nblk = len(self.blk_i1)
        self.ll = np.full((self.C, self.nblk), -np.inf)
        self.backward_filter()

    def backward_filter(self):
        # tables are a function of (law, layout): kept per layout until the law changes, like the device's cache
        # (a chequerboard sampler alternates between two layouts; recomputing them every sweep would only slow the
        # CPU baseline down)
        key = (self.blk_i1.tobytes(), self.blk_i2.tobytes(), self.blk_pinned.tobytes(), self.bf_scheme)
        cache = self.__dict__.setdefault("_bf_cache", {})
        if key in cache:
            self.H, self.F0, self.Psi, self.c0, self.blk_g, self.blk_Qc = cache[key]
            return
        d, P, nb = self.d, self.P, self.nblk
        self.H, self.F0, self.Psi = np.zeros((P, d, d)), np.zeros((P, d)), np.zeros((P, d, d))
        self.c0, self.blk_g, self.blk_Qc = np.zeros(P), np.zeros((nb, d)), np.zeros((nb, d, d))
        lib().orc_backward_filter_scheme(
            d, self.J, _i(self.N), _i(self.pt_off), _d(self.t), _d(self.auxB), _d(self.auxbeta),
            _d(self.auxatil), _d(self.Hobs), _d(self.Fobs), _d(self.cobs), nb, _i(self.blk_i1),
            _i(self.blk_i2), _i(self.blk_pinned), C.c_double(self.pin_eps), int(self.bf_scheme), _d(self.H),
            _d(self.F0), _d(self.Psi), _d(self.c0), _d(self.blk_g), _d(self.blk_Qc))
        if len(cache) >= 4:
            cache.clear()
        cache[key] = (self.H, self.F0, self.Psi, self.c0, self.blk_g, self.blk_Qc)

    # -- ctypes views -----------------------------------------------------------------------
    def _problem(self):
        p = _Problem()
        p.model, p.d, p.m, p.nth = self.model, self.d, self.m, len(self.theta)
        p.th, p.J, p.R = _d(self.theta), self.J, self.R
        p.N, p.pt_off, p.st_off, p.t = _i(self.N), _i(self.pt_off), _i(self.st_off), _d(self.t)
        p.auxB, p.auxbeta, p.auxsig = _d(self.auxB), _d(self.auxbeta), _d(self.auxsig)
        p.nblk, p.blk_i1, p.blk_i2, p.blk_pinned = self.nblk, _i(self.blk_i1), _i(self.blk_i2), _i(self.blk_pinned)
        p.H, p.F0, p.Psi, p.c0 = _d(self.H), _d(self.F0), _d(self.Psi), _d(self.c0)
        p.blk_g, p.blk_Qc, p.rho, p.C = _d(self.blk_g), _d(self.blk_Qc), _d(self.rho), self.C
        p.chain_offset, p.interval_offset = self.chain_offset, self.interval_offset
        return p

    def _state(self):
        s = _State()
        s.X[0], s.X[1], s.W[0], s.W[1] = _d(self.X[0]), _d(self.X[1]), _d(self.W[0]), _d(self.W[1])
        s.selX, s.selW, s.ll = _b(self.selX), _b(self.selW), _d(self.ll)
        return s

    # -- updates ----------------------------------------------------------------------------
    def impute(self, Z=None, E=None, seed=0, it=0, force_accept=False, unit_mask=None):
        """One PathImputation update (src/run!_alterations.jl:5-26) over every (chain, block).
        Z [C][S][m], E [C][nblk] explicit noise, or Philox(seed, it) when Z is None."""
        p, s = self._problem(), self._state()
        n = self.C * self.nblk
        llp, ll, acc = np.zeros(n), np.zeros(n), np.zeros(n, np.uint8)
        use_philox = Z is None
        if not use_philox:
            Z, E = f64(Z), f64(E)
            assert Z.shape == (self.C, self.S, self.m) and E.shape == (self.C, self.nblk)
        mask = None if unit_mask is None else np.ascontiguousarray(unit_mask, np.uint8)
        lib().orc_impute_sweep(
            C.byref(p), C.byref(s), None if use_philox else _d(Z), None if use_philox else _d(E),
            C.c_uint64(seed), C.c_uint32(it), int(use_philox), int(force_accept),
            None if mask is None else _b(mask), _d(llp), _d(ll), _b(acc), self.nthreads)
        sh = (self.C, self.nblk)
        return llp.reshape(sh), ll.reshape(sh), acc.reshape(sh).astype(bool)

    def init_paths(self, Z=None, seed=0, max_tries=100):
        """init_paths! (src/workspaces.jl:80-92): un-preconditioned (rho = 0) forward_guide!
        with retry until success, then init_ll! (:425-431)."""
        rho_keep = self.rho.copy()
        self.rho[:] = 0.0
        done = np.zeros((self.C, self.nblk), bool)
        for tr in range(max_tries):
            mask = (~done).astype(np.uint8)
            E = np.zeros((self.C, self.nblk))
            _, _, acc = self.impute(Z=Z, E=None if Z is None else E, seed=seed, it=0xFFFF0000 + tr,
                                    force_accept=True, unit_mask=mask)
            done |= acc & mask.astype(bool)
            if done.all():
                break
            if Z is not None:
                raise RuntimeError("init_paths with explicit Z failed the bound check")
        else:
            raise RuntimeError("init_paths: no admissible path")
        self.rho[:] = rho_keep
        self.init_ll()

    def init_ll(self):
        p, s = self._problem(), self._state()
        lib().orc_init_ll(C.byref(p), C.byref(s), self.nthreads)

    def relayout(self, blk_i1, blk_i2, blk_pinned):
        """Switch block layout: new tables, W re-inversion, ll recompute (OUR definition)."""
        self.set_layout(blk_i1, blk_i2, blk_pinned)
        p, s = self._problem(), self._state()
        lib().orc_relayout(C.byref(p), C.byref(s), self.nthreads)

    def param_solve(self, theta_prop, custom_aux=None):
        """set_proposal! (src/run!_alterations.jl:177-211): returns (ll_prop [C][nblk], success [C])
        and leaves X_prop in the proposal containers; call param_commit to accept."""
        self._saved_law = (self.theta, self.auxB, self.auxbeta, self.auxsig, self.auxatil,
                           self.H, self.F0, self.Psi, self.c0, self.blk_g, self.blk_Qc)
        self.set_theta(theta_prop, custom_aux)
        p, s = self._problem(), self._state()
        llp, ok = np.zeros((self.C, self.nblk)), np.zeros(self.C, np.uint8)
        lib().orc_param_solve(C.byref(p), C.byref(s), _d(llp), _b(ok), self.nthreads)
        self._llprop = llp
        return llp.copy(), ok.astype(bool)

    def param_commit(self, accepted):
        accepted = np.ascontiguousarray(np.broadcast_to(accepted, (self.C,)), np.uint8)
        if accepted.all():
            pass  # keep the proposal law
        elif not accepted.any():
            (self.theta, self.auxB, self.auxbeta, self.auxsig, self.auxatil, self.H, self.F0,
             self.Psi, self.c0, self.blk_g, self.blk_Qc) = self._saved_law
            self.__dict__.pop("_bf_cache", None)   # it holds the rejected law's tables
        else:
            raise ValueError("one theta-instance per Oracle: accept all chains or none")
        p, s = self._problem(), self._state()
        lib().orc_param_commit(C.byref(p), C.byref(s), _b(accepted), _d(self._llprop))

    # -- read-outs (reference-like shapes) ---------------------------------------------------
    def accepted_X(self):
        """[C][P][d] accepted containers (per-interval N_j+1 points, duplicates at the joins)."""
        out = np.empty_like(self.X[0])
        for j in range(self.J):
            sl = slice(self.pt_off[j], self.pt_off[j + 1])
            sel = self.selX[:, j].astype(bool)
            out[:, sl] = np.where(sel[:, None, None], self.X[1][:, sl], self.X[0][:, sl])
        return out

    def accepted_W(self):
        out = np.empty_like(self.W[0])
        for j in range(self.J):
            sl = slice(self.pt_off[j], self.pt_off[j + 1])
            sel = self.selW[:, j].astype(bool)
            out[:, sl] = np.where(sel[:, None, None], self.W[1][:, sl], self.W[0][:, sl])
        return out

    def proposal_X(self):
        out = np.empty_like(self.X[0])
        for j in range(self.J):
            sl = slice(self.pt_off[j], self.pt_off[j + 1])
            sel = self.selX[:, j].astype(bool)
            out[:, sl] = np.where(sel[:, None, None], self.X[0][:, sl], self.X[1][:, sl])
        return out

    def proposal_W(self):
        out = np.empty_like(self.W[0])
        for j in range(self.J):
            sl = slice(self.pt_off[j], self.pt_off[j + 1])
            sel = self.selW[:, j].astype(bool)
            out[:, sl] = np.where(sel[:, None, None], self.W[0][:, sl], self.W[1][:, sl])
        return out

    def glue_path(self, chain, rec=0):
        """_copy_path! (src/path_saving_buffer.jl:54-64): [(sum N_j)+1][d] for one recording."""
        firsts = np.flatnonzero(self.rec_first)
        i1 = int(firsts[rec])
        i2 = i1 + int(self.rec_nobs[rec]) - 1
        npts = int(self.N[i1:i2 + 1].sum()) + 1
        out = np.zeros((npts, self.d))
        p, s = self._problem(), self._state()
        lib().orc_glue_path(C.byref(p), C.byref(s), chain, i1, i2, _d(out))
        return out
