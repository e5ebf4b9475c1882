"""NumPy-facing wrapper of the C ABI (include/dmcmc.h): one `Engine` = one `dmcmc_ctx` = the
device-resident replacement of the reference's `DiffusionGlobalSubworkspace` pair
(`sub_ws_diff`, `sub_ws_diff°`; /root/reference src/workspaces.jl:49-72, :119-127): accepted and
proposal containers of X and W for every chain, the guided-proposal laws and their
backward-filter tables.

Everything numerical happens in libdmcmc_b200.so on the GPU; this class only marshals arrays.
"""
import ctypes as C

import os

import numpy as np

from . import _lib
from .problems import chequerboard_layout


def _f64(a):
    return np.ascontiguousarray(a, dtype=np.float64)


def _i32(a):
    return np.ascontiguousarray(a, dtype=np.int32)


def _dp(a):
    return None if a is None else a.ctypes.data_as(_lib._dp)


def _ip(a):
    return a.ctypes.data_as(_lib._ip)


def _bp(a):
    return None if a is None else a.ctypes.data_as(_lib._bp)


class DmcmcError(RuntimeError):
    pass


class Engine:
    def __init__(self, spec, device=0, seed=0, chain_offset=0, n_chains=None, interval_offset=0,
                 store_noise=True):
        self.lib = _lib.load()
        self.spec = spec
        self.model, self.d, self.m = int(spec["model"]), int(spec["d"]), int(spec["m"])
        x0 = _f64(spec["x0"])
        if n_chains is not None:
            x0 = x0[chain_offset:chain_offset + n_chains] if x0.shape[0] >= chain_offset + n_chains \
                else np.broadcast_to(x0[:1], (n_chains,) + x0.shape[1:]).copy()
        self.C = x0.shape[0]
        cfg = _lib.Config(_lib.ABI_VERSION, self.model, self.d, self.C, chain_offset, interval_offset, device,
                          seed, float(spec.get("pin_eps", 1e-4)), int(bool(store_noise)))
        h = C.c_void_p()
        if self.lib.dmcmc_create(C.byref(cfg), C.byref(h)) != 0:
            raise DmcmcError(self.lib.dmcmc_last_error(None).decode())
        self.h = h
        if os.environ.get("DMCMC_COOP"):  # test hook: pin the imputation kernel (0 thread-per-unit, 1 warp-per-unit)
            self.set_cooperative(int(os.environ["DMCMC_COOP"]))
        self.N = _i32(spec["N"])
        self.J = len(self.N)
        self.pt_off = np.concatenate([[0], np.cumsum(self.N + 1)]).astype(np.int64)
        self.st_off = np.concatenate([[0], np.cumsum(self.N)]).astype(np.int64)
        self.P, self.S = int(self.pt_off[-1]), int(self.st_off[-1])
        rec_first = _i32(spec["rec_first"])
        self.rec_nobs = _i32(np.diff(np.concatenate([np.flatnonzero(rec_first), [self.J]])))
        self.R = len(self.rec_nobs)
        t = _f64(spec["t"])
        self._ck(self.lib.dmcmc_set_grid(self.h, self.R, _ip(self.rec_nobs), _ip(self.N), _dp(t)))
        L, Sg, v = _f64(spec["obs_L"]), _f64(spec["obs_Sigma"]), _f64(spec["obs_v"])
        self.p = L.shape[1]
        self._ck(self.lib.dmcmc_set_obs(self.h, self.p, _dp(L), _dp(Sg), _dp(v)))
        self._ck(self.lib.dmcmc_set_start(self.h, _dp(x0)))
        self.set_theta(spec["theta"])
        self.set_rho(np.broadcast_to(spec["rho"], (self.J,)))
        self.set_layout(*chequerboard_layout(self.rec_nobs, 0, 0))
        self.backward_filter()

    # -- plumbing ---------------------------------------------------------------------------
    def _ck(self, rc):
        if rc != 0:
            raise DmcmcError(self.lib.dmcmc_last_error(self.h).decode())

    def close(self):
        if getattr(self, "h", None):
            self.lib.dmcmc_destroy(self.h)
            self.h = None
            for ptr in getattr(self, "_pinned", []):
                self.lib.dmcmc_host_free(ptr)
            self._pinned = []

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # -- problem ----------------------------------------------------------------------------
    def set_theta(self, theta):
        th = _f64(theta)
        self._ck(self.lib.dmcmc_set_theta(self.h, _dp(th), len(th)))
        self.theta = th

    def set_aux(self, B, beta, sig, which=0):
        B, beta, sig = _f64(B), _f64(beta), _f64(sig)
        self._ck(self.lib.dmcmc_set_aux(self.h, which, _dp(B), _dp(beta), _dp(sig)))

    def set_rho(self, rho):
        rho = _f64(rho)
        self._ck(self.lib.dmcmc_set_rho(self.h, _dp(rho)))
        self.rho = rho

    def set_layout(self, i1, i2, pinned):
        self.blk_i1, self.blk_i2, self.blk_pinned = _i32(i1), _i32(i2), _i32(pinned)
        self.nblk = len(self.blk_i1)
        self._ck(self.lib.dmcmc_set_layout(self.h, self.nblk, _ip(self.blk_i1), _ip(self.blk_i2),
                                           _ip(self.blk_pinned)))

    def relayout(self, i1, i2, pinned):
        self.blk_i1, self.blk_i2, self.blk_pinned = _i32(i1), _i32(i2), _i32(pinned)
        self.nblk = len(self.blk_i1)
        self._ck(self.lib.dmcmc_relayout(self.h, self.nblk, _ip(self.blk_i1), _ip(self.blk_i2),
                                         _ip(self.blk_pinned)))

    def set_block_mask(self, active):
        if active is None:
            self._ck(self.lib.dmcmc_set_block_mask(self.h, None))
        else:
            a = _i32(active)
            assert len(a) == self.nblk
            self._ck(self.lib.dmcmc_set_block_mask(self.h, _ip(a)))

    def set_start(self, x0):
        x0 = _f64(x0)
        assert x0.shape == (self.C, self.R, self.d)
        self._ck(self.lib.dmcmc_set_start(self.h, _dp(x0)))

    def download_segment(self, j0, j1):
        n = int(self.N[j0:j1].sum())
        X = np.empty((self.C, n, self.d))
        self._ck(self.lib.dmcmc_download_segment(self.h, j0, j1, _dp(X)))
        return X

    def upload_segment(self, j0, j1, X):
        X = _f64(X)
        assert X.shape == (self.C, int(self.N[j0:j1].sum()), self.d)
        self._ck(self.lib.dmcmc_upload_segment(self.h, j0, j1, _dp(X)))

    # -- device-side exchange (block sharding over GPUs): raw device pointers, stream-ordered, no host sync --
    def segment_device(self, j0, j1, dev_ptr, chain_stride, upload):
        self._ck(self.lib.dmcmc_segment_device(self.h, j0, j1, C.c_void_p(dev_ptr), int(chain_stride), int(upload)))

    def set_start_device(self, dev_ptr, chain_stride):
        self._ck(self.lib.dmcmc_set_start_device(self.h, C.c_void_p(dev_ptr), int(chain_stride)))

    def stream(self):
        return int(self.lib.dmcmc_stream(self.h) or 0)

    def impute_async(self, it):
        self._ck(self.lib.dmcmc_impute_async(self.h, it))

    def backward_filter(self, which=0):
        self._ck(self.lib.dmcmc_backward_filter(self.h, which))

    def upload_tables(self, H, F0, Psi, c0, blk_g=None, blk_Qc=None, which=0):
        H, F0, c0 = _f64(H), _f64(F0), _f64(c0)
        Psi = None if Psi is None else _f64(Psi)
        blk_g = None if blk_g is None else _f64(blk_g)
        blk_Qc = None if blk_Qc is None else _f64(blk_Qc)
        self._ck(self.lib.dmcmc_upload_tables(self.h, which, _dp(H), _dp(F0), _dp(Psi), _dp(c0),
                                              _dp(blk_g), _dp(blk_Qc)))

    def download_tables(self, which=0):
        d, P, nb = self.d, self.P, self.nblk
        H, F0, Psi = np.zeros((P, d, d)), np.zeros((P, d)), np.zeros((P, d, d))
        c0, g, Qc = np.zeros(P), np.zeros((nb, d)), np.zeros((nb, d, d))
        self._ck(self.lib.dmcmc_download_tables(self.h, which, _dp(H), _dp(F0), _dp(Psi), _dp(c0),
                                                _dp(g), _dp(Qc)))
        return dict(H=H, F0=F0, Psi=Psi, c0=c0, blk_g=g, blk_Qc=Qc)

    # -- updates ----------------------------------------------------------------------------
    def init_paths(self, Z=None):
        Z = None if Z is None else _f64(Z)
        self._ck(self.lib.dmcmc_init_paths(self.h, _dp(Z)))

    def init_ll(self):
        self._ck(self.lib.dmcmc_init_ll(self.h))

    def impute(self, it=0, Z=None, E=None, want=True):
        n = (self.C, self.nblk)
        llp = np.empty(n) if want else None
        ll = np.empty(n) if want else None
        acc = np.empty(n, np.uint8) if want else None
        if Z is not None:
            Z, E = _f64(Z), _f64(E)
            assert Z.shape == (self.C, self.S, self.m) and E.shape == n
        self._ck(self.lib.dmcmc_impute(self.h, it, _dp(Z), _dp(E), _dp(llp), _dp(ll), _bp(acc)))
        if want:
            return llp, ll, acc.astype(bool)

    def set_adaptation(self, adpt, sweeps_per_iteration=1, rho0=None, rho1=None):
        """dmcmc_set_adaptation: `adpt` is an AdaptationPathImputation (or None to switch the device-side adaptation off)."""
        if adpt is None:
            self._ck(self.lib.dmcmc_set_adaptation(self.h, None, 1, None, None))
            return
        cfg = _lib.Adapt(adpt.target_accpt_rate, adpt.scale, adpt.min, adpt.max, adpt.offset, adpt.adapt_every_k_steps)
        r0 = None if rho0 is None else _f64(rho0)
        r1 = None if rho1 is None else _f64(rho1)
        self._ck(self.lib.dmcmc_set_adaptation(self.h, C.byref(cfg), int(sweeps_per_iteration), _dp(r0), _dp(r1)))

    def get_rho(self, pattern=0):
        rho = np.empty(self.J)
        self._ck(self.lib.dmcmc_get_rho(self.h, int(pattern), _dp(rho)))
        return rho

    def pinned(self, shape, dtype):
        """NumPy array over pinned host memory (dmcmc_host_alloc); freed with the engine."""
        dtype = np.dtype(dtype)
        n = int(np.prod(shape)) * dtype.itemsize
        ptr = self.lib.dmcmc_host_alloc(max(n, 1))
        if not ptr:
            raise DmcmcError("dmcmc_host_alloc failed")
        self._pinned = getattr(self, "_pinned", [])
        self._pinned.append(ptr)
        buf = (C.c_char * max(n, 1)).from_address(ptr)
        return np.frombuffer(buf, dtype=dtype, count=int(np.prod(shape))).reshape(shape)

    def run_sweeps(self, it0, n_sweeps, half_len=0, first_pattern=0, rho_sched=None,
                   hist=None):
        """dmcmc_run_sweeps.  hist = (ll_prop, ll, accepted) arrays [n][C][stride] or None."""
        stride = 0
        llp = ll = acc = None
        if hist is not None:
            llp, ll, acc = hist
            stride = llp.shape[2]
        self._ck(self.lib.dmcmc_run_sweeps(self.h, it0, n_sweeps, half_len, first_pattern,
                                           _dp(rho_sched), stride, _dp(llp), _dp(ll), _bp(acc)))
        self._sync_nblk()

    def run_imputation(self, it0, n_iters, histories=True):
        hist = None
        if histories:
            n = (n_iters, self.C, self.nblk)
            hist = (np.empty(n), np.empty(n), np.empty(n, np.uint8))
        self.run_sweeps(it0, n_iters, hist=hist)
        return hist if histories else (None, None, None)

    def param_loglik(self, theta_prop, critical=True):
        th = _f64(theta_prop)
        llp = np.empty((self.C, self.nblk))
        ok = np.empty(self.C, np.uint8)
        self._ck(self.lib.dmcmc_param_loglik(self.h, _dp(th), len(th), int(critical), _dp(llp), _bp(ok)))
        return llp, ok.astype(bool)

    def param_commit(self, accepted):
        self._ck(self.lib.dmcmc_param_commit(self.h, int(bool(accepted))))

    # -- block-sharded recording: collectives inside the library (NCCL bound at run time) ----------------
    @staticmethod
    def comm_unique_id():
        """128-byte NCCL unique id (dmcmc_comm_unique_id): made by one rank, handed to the others by the host."""
        buf = C.create_string_buffer(128)
        if _lib.load().dmcmc_comm_unique_id(buf) != 0:
            raise DmcmcError(_lib.load().dmcmc_last_error(None).decode())
        return buf.raw

    def comm_init(self, unique_id, rank, world):
        self._ck(self.lib.dmcmc_comm_init(self.h, C.c_char_p(unique_id), int(rank), int(world)))

    def comm_enable_p2p(self, mailbox_bytes):
        """Halo exchange over peer memory (CUDA IPC mailboxes, NVLink stores + flags) instead of NCCL send / recv;
        collective.  Returns True when every rank could map its neighbours (otherwise the exchange stays on NCCL)."""
        self._ck(self.lib.dmcmc_comm_enable_p2p(self.h, int(mailbox_bytes)))
        return bool(self.lib.dmcmc_comm_p2p_enabled(self.h))

    def impute_exchange_async(self, it, pattern, n_first, n_halo, left_edge_steps):
        """impute_async + the halo_exchange that follows it under `pattern`, fused into the sweep kernels when peer memory
        is on and the launch goes to the warp-per-unit kernel (else the two calls one after the other)."""
        self._ck(self.lib.dmcmc_impute_exchange_async(self.h, int(it), int(pattern), int(n_first), int(n_halo), int(left_edge_steps)))

    def comm_destroy(self):
        self._ck(self.lib.dmcmc_comm_destroy(self.h))

    def halo_exchange(self, after_pattern, n_first, n_halo, left_edge_steps):
        self._ck(self.lib.dmcmc_halo_exchange(self.h, int(after_pattern), int(n_first), int(n_halo), int(left_edge_steps)))

    def param_loglik_sharded(self, theta_prop, nblk_global, blk_global, critical=True):
        """dmcmc_param_loglik_sharded: (ll deg, ll) of every GLOBAL block [C][nblk_global] and success [C], identical on
        every rank."""
        th, gi = _f64(theta_prop), _i32(blk_global)
        assert len(gi) == self._sync_nblk()
        llp, ll = np.empty((self.C, nblk_global)), np.empty((self.C, nblk_global))
        ok = np.empty(self.C, np.uint8)
        self._ck(self.lib.dmcmc_param_loglik_sharded(self.h, _dp(th), len(th), int(critical), int(nblk_global), _ip(gi),
                                                     _dp(llp), _dp(ll), _bp(ok)))
        return llp, ll, ok.astype(bool)

    def conjugate_coords(self):
        """theta indices of the law's conjugate parameters (dmcmc_conjugate_dim)."""
        idx = np.zeros(8, np.int32)
        n = int(self.lib.dmcmc_conjugate_dim(self.h, _ip(idx)))
        return idx[:max(n, 0)].copy()

    def conjugate_sums(self):
        """(mu [C][nc], W [C][nc][nc]) of compute_mu_and_W over the accepted path, reduced on the device."""
        nc = len(self.conjugate_coords())
        mu, W = np.empty((self.C, nc)), np.empty((self.C, nc, nc))
        self._ck(self.lib.dmcmc_conjugate_sums(self.h, _dp(mu), _dp(W)))
        return mu, W

    # -- read-outs --------------------------------------------------------------------------
    def download_state(self, which=0):
        X = np.empty((self.C, self.P, self.d))
        W = np.empty((self.C, self.P, self.m))
        self._ck(self.lib.dmcmc_download_state(self.h, which, _dp(X), _dp(W)))
        return X, W

    def download_path(self, chain, rec=0):
        j0 = int(self.rec_nobs[:rec].sum())
        npts = int(self.N[j0:j0 + self.rec_nobs[rec]].sum()) + 1
        out = np.empty((npts, self.d))
        self._ck(self.lib.dmcmc_download_path(self.h, chain, rec, _dp(out)))
        return out

    def save_path_async(self, chains, rec=0):
        """dmcmc_save_path_async: snapshot of the glued accepted paths of `chains`; returns a ticket for save_path_wait."""
        ch = _i32(chains)
        tk = C.c_int32()
        self._ck(self.lib.dmcmc_save_path_async(self.h, len(ch), _ip(ch), rec, C.byref(tk)))
        j0 = int(self.rec_nobs[:rec].sum())
        npts = int(self.N[j0:j0 + self.rec_nobs[rec]].sum()) + 1
        return (tk.value, len(ch), npts)

    def save_path_wait(self, ticket):
        tk, n, npts = ticket
        out = np.empty((n, npts, self.d))
        self._ck(self.lib.dmcmc_save_path_wait(self.h, tk, _dp(out)))
        return out

    def _sync_nblk(self):
        self.nblk = int(self.lib.dmcmc_n_blocks(self.h))
        return self.nblk

    def get_ll(self):
        self._sync_nblk()
        ll = np.empty((self.C, self.nblk))
        self._ck(self.lib.dmcmc_get_ll(self.h, _dp(ll)))
        return ll

    def set_ll(self, ll):
        self._sync_nblk()
        ll = _f64(ll)
        assert ll.shape == (self.C, self.nblk)
        self._ck(self.lib.dmcmc_set_ll(self.h, _dp(ll)))

    def accept_counts(self):
        self._sync_nblk()
        a, p = np.zeros(self.nblk, np.int64), np.zeros(self.nblk, np.int64)
        self._ck(self.lib.dmcmc_accept_counts(self.h, a.ctypes.data_as(_lib._lp), p.ctypes.data_as(_lib._lp)))
        return a, p

    def set_cooperative(self, mode):
        """-1: imputation kernel chosen by unit count (default); 0: thread per (chain, block) unit; 1: warp per unit."""
        self._ck(self.lib.dmcmc_set_cooperative(self.h, int(mode)))

    def kernel_times(self):
        """(ms[4], launches[4]) per kernel mode since the last call: impute, param, relayout, ll."""
        ms, n = np.zeros(4), np.zeros(4, np.int64)
        self._ck(self.lib.dmcmc_kernel_times(self.h, _dp(ms), n.ctypes.data_as(_lib._lp)))
        return ms, n

    def set_timing(self, every):
        """Per-kernel duration events around every n-th solve launch (default 8; 0 = never)."""
        self._ck(self.lib.dmcmc_set_timing(self.h, int(every)))

    def launch_count(self):
        return int(self.lib.dmcmc_launch_count(self.h))

    def timer_start(self):
        self._ck(self.lib.dmcmc_timer_start(self.h))

    def timer_stop(self):
        ms = C.c_double()
        self._ck(self.lib.dmcmc_timer_stop(self.h, C.byref(ms)))
        return ms.value

    def kernel_ms(self):
        ms, n = self.kernel_times()
        return float(ms[0]), int(n[0])

    def device_bytes(self):
        return int(self.lib.dmcmc_device_bytes(self.h))
