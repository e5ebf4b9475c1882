"""Multi-GPU decomposition of the imputation path.

Two ways the path shards, both one process/context per GPU:

* chains: independent, no collective at all -- every rank builds its `Engine` with
  `chain_offset = rank * chains_per_rank` (Philox streams are keyed by the GLOBAL chain index, so
  the union of the shards is bit-identical to one big context).  See `shard_chains`.

* blocks of one long recording (BASELINE config 4): rank r owns the contiguous interval range
  [a_r, a_{r+1}) whose ends are chequerboard delimiters of pattern A.  Blocks of pattern A never
  cross a rank edge.  Blocks of pattern B straddle it, so every rank but the last also holds a
  halo copy of the next h intervals and updates the straddling block itself; after each sweep the
  freshly imputed half is exchanged with the neighbour:
      after sweep A : rank r+1 -> r   its first h intervals        (h*N*d doubles per chain)
      after sweep B : rank r   -> r+1 the halo + the edge point x(t_{a_{r+1}})
  These are the "block-boundary endpoint exchange" messages of the north star: a few KB, latency
  bound, sent with NCCL send/recv (gloo in the CPU tests).  Philox counters use the GLOBAL interval
  index (`interval_offset`), and each block's backward-filter table depends only on the block's own
  observations, so a sharded run reproduces the single-GPU run bit for bit.
"""
import os

import numpy as np


def shard_chains(n_chains, world):
    """[(offset, count)] per rank: contiguous, sizes differing by at most one."""
    base, rem = divmod(n_chains, world)
    out, off = [], 0
    for r in range(world):
        c = base + (1 if r < rem else 0)
        out.append((off, c))
        off += c
    return out


def shard_intervals(n_intervals, world, half_len):
    """Rank edges a_0=0 < a_1 < ... < a_world = n, every interior edge a multiple of 2*half_len."""
    unit = 2 * half_len
    nunits = n_intervals // unit
    if nunits < world:
        raise ValueError("too few chequerboard blocks for this many ranks")
    edges = [0]
    for r in range(1, world):
        edges.append((nunits * r // world) * unit)
    edges.append(n_intervals)
    return edges


def local_spec(spec, a, b_ext, x_start):
    """Sub-problem on intervals [a, b_ext) of a single-recording spec, starting from x_start [C][d]."""
    N = np.asarray(spec["N"])
    pt_off = np.concatenate([[0], np.cumsum(N + 1)])
    out = dict(spec)
    out["N"] = N[a:b_ext].copy()
    out["t"] = np.asarray(spec["t"])[pt_off[a]:pt_off[b_ext]].copy()
    rf = np.zeros(b_ext - a, np.int32)
    rf[0] = 1
    out["rec_first"] = rf
    for k in ("obs_L", "obs_Sigma", "obs_v", "rho"):
        out[k] = np.asarray(spec[k])[a:b_ext].copy()
    out["x0"] = np.asarray(x_start, float)[:, None, :].copy()
    return out


def local_layouts(a, b, b_ext, n, half_len, rank):
    """Local (0-based from a) block ranges, pinned flags and activity masks of the two patterns."""
    h = half_len
    lays = []
    for pattern in (0, 1):
        i1, i2, pin, act = [], [], [], []
        start = a
        if pattern == 1:  # first (half) block [a, a+h): the true first block on rank 0, else owned by the left rank
            i1.append(a), i2.append(min(a + h, n) - 1), act.append(1 if rank == 0 else 0)
            start = min(a + h, n)
        limit = b if pattern == 0 else b_ext
        while start < limit:
            end = min(start + 2 * h, n)
            if pattern == 0:
                end = min(end, b)
            i1.append(start), i2.append(end - 1), act.append(1)
            start = end
        if pattern == 0 and b_ext > b:  # halo filler, updated by the right neighbour in pattern A
            i1.append(b), i2.append(b_ext - 1), act.append(0)
        i1, i2 = np.array(i1, np.int32), np.array(i2, np.int32)
        pin = (i2 < n - 1).astype(np.int32)
        lays.append((i1 - a, i2 - a, pin, np.array(act, np.int32)))
    return lays


class BlockShardedSampler:
    """One rank of a block-sharded chequerboard sampler over a single long recording.

    `make_engine(local_spec, interval_offset)` returns an object with the Engine interface
    (set_layout, set_block_mask, backward_filter, impute, download_segment, upload_segment,
    set_start, set_rho); `full_path` is the initial accepted path [C][S+1][d] of the WHOLE record
    (every rank computes it identically, see `initial_full_path`)."""

    def __init__(self, spec, rank, world, half_len, make_engine, full_path):
        self.rank, self.world, self.h = rank, world, half_len
        N = np.asarray(spec["N"])
        self.n = len(N)
        self.N_max = int(N.max())   # longest interval of the whole recording (the same number on every rank)
        edges = shard_intervals(self.n, world, half_len)
        self.a, self.b = edges[rank], edges[rank + 1]
        self.b_ext = min(self.b + half_len, self.n) if rank < world - 1 else self.b
        st_off = np.concatenate([[0], np.cumsum(N)])
        self.Nloc = N[self.a:self.b_ext]
        # full_path has one point per step plus the start: point index = 1 + global step index
        x_start = full_path[:, st_off[self.a], :]
        self.engine = make_engine(local_spec(spec, self.a, self.b_ext, x_start), self.a)
        seg = full_path[:, st_off[self.a] + 1:st_off[self.b_ext] + 1, :]
        self.engine.upload_segment(0, self.b_ext - self.a, seg)
        self.lays = local_layouts(self.a, self.b, self.b_ext, self.n, half_len, rank)
        self.N_left_edge = int(N[self.a - 1]) if self.a > 0 else 0
        self.sweeps = 0

    # -- one PathImputation update under pattern `pat` (layout switch fused into the kernel) ---
    def sweep(self, pat, it):
        i1, i2, pin, act = self.lays[pat]
        e = self.engine
        e.set_layout(i1, i2, pin)
        e.set_block_mask(act)
        out = e.impute(it)
        self.sweeps += 1
        return out, act

    def sweep_async(self, pat, it):
        """The same update, enqueued on the engine's stream: nothing comes back to the host."""
        i1, i2, pin, act = self.lays[pat]
        e = self.engine
        e.set_layout(i1, i2, pin)
        e.set_block_mask(act)
        e.impute_async(it)
        self.sweeps += 1

    # -- halo messages ------------------------------------------------------------------------
    def pack_for_left(self):
        """after sweep A: my first h intervals, for the left neighbour's halo."""
        return self.engine.download_segment(0, min(self.h, self.b - self.a))

    def unpack_from_right(self, X):
        self.engine.upload_segment(self.b - self.a, self.b_ext - self.a, X)

    def pack_for_right(self):
        """after sweep B: the halo I just updated plus the edge point x(t_b) (my interval b-1's end)."""
        halo = self.engine.download_segment(self.b - self.a, self.b_ext - self.a)
        edge = self.engine.download_segment(self.b - self.a - 1, self.b - self.a)[:, -1:, :]
        return np.concatenate([edge, halo], axis=1)

    def unpack_from_left(self, msg):
        e = self.engine
        e.set_start(msg[:, :1, :].copy())           # [C][1][d]: my new starting point
        e.upload_segment(0, min(self.h, self.b - self.a), msg[:, 1:, :])

    def owned_path(self):
        """accepted X of the intervals this rank owns, [C][steps][d]."""
        return self.engine.download_segment(0, self.b - self.a)

    def msg_shape_right(self, C, d):
        steps = int(self.Nloc[self.b - self.a:self.b_ext - self.a].sum())
        return (C, 1 + steps, d)

    def msg_shape_left(self, C, d):
        return (C, int(self.Nloc[:min(self.h, self.b - self.a)].sum()), d)


def iterate(sampler, comm, it):
    """One MCMC iteration = sweep A, exchange, sweep B, exchange (all ranks call this together).
    `comm` has send(array, dst) [non-blocking], recv(shape, src) and flush()."""
    s = sampler
    C, d = s.engine.C, s.engine.d
    out_a = s.sweep(0, 2 * it)
    post_a(s, comm)
    take_a(s, comm, C, d)
    out_b = s.sweep(1, 2 * it + 1)
    post_b(s, comm)
    take_b(s, comm, C, d)
    return out_a, out_b


def post_a(s, comm):
    if s.world > 1 and s.rank > 0:
        comm.send(s.pack_for_left(), s.rank - 1)      # rank r -> r-1 : my first h intervals


def take_a(s, comm, C, d):
    if s.world > 1 and s.rank < s.world - 1:
        steps = int(s.Nloc[s.b - s.a:s.b_ext - s.a].sum())
        s.unpack_from_right(comm.recv((C, steps, d), s.rank + 1))
    comm.flush()


def post_b(s, comm):
    if s.world > 1 and s.rank < s.world - 1:
        comm.send(s.pack_for_right(), s.rank + 1)     # rank r -> r+1 : halo + edge point


def take_b(s, comm, C, d):
    if s.world > 1 and s.rank > 0:
        steps = int(s.Nloc[:min(s.h, s.b - s.a)].sum())
        s.unpack_from_left(comm.recv((C, 1 + steps, d), s.rank - 1))
    comm.flush()


class DeviceHalo:
    """Device-resident halo exchange of a BlockShardedSampler (GPU only): the packed segments live in
    CUDA tensors, pack/unpack kernels and the imputation sweeps are enqueued on the engine's stream,
    and the NCCL send/recv (`batch_isend_irecv`) is ordered on that same stream -- an iteration never
    synchronises with the host.  Messages are those of `iterate` (module docstring):
      after sweep A : rank r -> r-1   my first h intervals
      after sweep B : rank r -> r+1   my last owned interval (its end point is the neighbour's new
                                       starting point) followed by the halo intervals."""

    def __init__(self, sampler):
        import torch
        import torch.distributed as dist
        self.torch, self.dist, self.s = torch, dist, sampler
        s, e = sampler, sampler.engine
        self.stream = torch.cuda.ExternalStream(e.stream())
        C, d = e.C, e.d
        mk = lambda steps: torch.empty((C, int(steps), d), dtype=torch.float64, device="cuda")
        n_own = s.b - s.a
        self.h_first = min(s.h, n_own)
        self.n_first = int(s.Nloc[:self.h_first].sum())            # my first h intervals
        self.n_halo = int(s.Nloc[n_own:s.b_ext - s.a].sum())       # my halo (the right neighbour's first h)
        self.n_edge = int(s.Nloc[n_own - 1])                       # my last owned interval
        self.n_edge_left = int(s.N_left_edge) if s.rank > 0 else 0  # the left neighbour's last owned interval
        self.to_left = mk(self.n_first) if s.rank > 0 else None
        self.from_right = mk(self.n_halo) if s.rank < s.world - 1 else None
        self.to_right = mk(self.n_edge + self.n_halo) if s.rank < s.world - 1 else None
        self.from_left = mk(self.n_edge_left + self.n_first) if s.rank > 0 else None

    def _exchange(self, send, dst, recv, src):
        ops = []
        if send is not None:
            ops.append(self.dist.P2POp(self.dist.isend, send, dst))
        if recv is not None:
            ops.append(self.dist.P2POp(self.dist.irecv, recv, src))
        if ops:
            for w in self.dist.batch_isend_irecv(ops):
                w.wait()  # orders the current (= the engine's) stream after the transfer; the host does not block

    # the four phases of an iteration (tests drive them rank by rank on one device)
    def sweep_a(self, it):
        s, e, d = self.s, self.s.engine, self.s.engine.d
        s.sweep_async(0, 2 * it)
        if s.rank > 0:
            e.segment_device(0, self.h_first, self.to_left.data_ptr(), self.n_first * d, 0)

    def take_a(self):
        s, e, d = self.s, self.s.engine, self.s.engine.d
        if s.rank < s.world - 1:
            e.segment_device(s.b - s.a, s.b_ext - s.a, self.from_right.data_ptr(), self.n_halo * d, 1)

    def sweep_b(self, it):
        s, e, d = self.s, self.s.engine, self.s.engine.d
        s.sweep_async(1, 2 * it + 1)
        if s.rank < s.world - 1:
            e.segment_device(s.b - s.a - 1, s.b_ext - s.a, self.to_right.data_ptr(), (self.n_edge + self.n_halo) * d, 0)

    def take_b(self):
        s, e, d = self.s, self.s.engine, self.s.engine.d
        if s.rank > 0:
            stride = (self.n_edge_left + self.n_first) * d
            base = self.from_left.data_ptr()
            e.set_start_device(base + (self.n_edge_left - 1) * d * 8, stride)
            e.segment_device(0, self.h_first, base + self.n_edge_left * d * 8, stride, 1)

    def iterate(self, it):
        s = self.s
        with self.torch.cuda.stream(self.stream):
            self.sweep_a(it)
            if s.world > 1:
                self._exchange(self.to_left, s.rank - 1, self.from_right, s.rank + 1)
            self.take_a()
            self.sweep_b(it)
            if s.world > 1:
                self._exchange(self.to_right, s.rank + 1, self.from_left, s.rank - 1)
            self.take_b()


class LibHalo:
    """The same exchange with the collectives INSIDE the library (dmcmc_comm_init / dmcmc_halo_exchange /
    dmcmc_param_loglik_sharded: NCCL send/recv and all-reduce enqueued on the context's stream by the C ABI itself) --
    what a host without torch.distributed, i.e. the Julia plugin, calls.  `unique_id` comes from
    Engine.comm_unique_id() on one rank and is handed to the others by the host."""

    def __init__(self, sampler, unique_id, p2p=None, fused=None):
        s = self.s = sampler
        self.fused = (os.environ.get("DMCMC_P2P_FUSE", "1") != "0") if fused is None else bool(fused)
        s.engine.comm_init(unique_id, s.rank, s.world)
        n_own = s.b - s.a
        self.n_first = min(s.h, n_own)
        self.n_halo = s.b_ext - s.b
        self.left_edge_steps = int(s.N_left_edge) if s.rank > 0 else 0
        # peer-memory mailboxes for the exchange (DMCMC_P2P=0: stay on ncclSend / ncclRecv).  Size: the largest message of
        # any rank, h + 1 intervals of the longest interval of the whole recording (the same number on every rank)
        self.p2p = False
        if p2p is None:
            p2p = os.environ.get("DMCMC_P2P", "1") != "0"
        if p2p and s.world > 1:
            e = s.engine
            self.p2p = e.comm_enable_p2p((s.h + 1) * s.N_max * e.d * 8 * e.C)

    def iterate(self, it):
        s, e = self.s, self.s.engine
        if s.world > 1 and self.p2p and self.fused:
            # sweep + exchange in one call: the edge block's warps push / pull inside the sweep kernels
            for pat in (0, 1):
                i1, i2, pin, act = s.lays[pat]
                e.set_layout(i1, i2, pin)
                e.set_block_mask(act)
                e.impute_exchange_async(2 * it + pat, pat, self.n_first, self.n_halo, self.left_edge_steps)
                s.sweeps += 1
            return
        s.sweep_async(0, 2 * it)
        if s.world > 1:
            e.halo_exchange(0, self.n_first, self.n_halo, self.left_edge_steps)
        s.sweep_async(1, 2 * it + 1)
        if s.world > 1:
            e.halo_exchange(1, self.n_first, self.n_halo, self.left_edge_steps)

    def global_blocks(self):
        """(nblk_global, global index of every local block of pattern 0, -1 for the halo filler)."""
        s = self.s
        i1, i2, pin, act = s.lays[0]
        unit = 2 * s.h
        gidx = np.where(act > 0, (s.a + i1) // unit, -1).astype(np.int32)
        return (s.n + unit - 1) // unit, gidx

    def param_step(self, theta_prop, decide, critical=True):
        """Parameter update on the sharded recording under pattern 0 (its blocks never cross a rank edge): every rank
        solves its blocks, ll deg is all-reduced per global block, `decide(llp_global, ll_global, ok)` -> bool must be
        the same function of the same numbers on every rank (shared seed), then every rank commits."""
        s, e = self.s, self.s.engine
        i1, i2, pin, act = s.lays[0]
        e.relayout(i1, i2, pin)
        e.set_block_mask(act)
        nb_g, gidx = self.global_blocks()
        llp, ll, ok = e.param_loglik_sharded(theta_prop, nb_g, gidx, critical)
        accepted = bool(decide(llp, ll, ok))
        e.param_commit(accepted)
        return llp, ll, ok, accepted


class TorchComm:
    """send/recv of float64 NumPy arrays over torch.distributed (NCCL: staged through a CUDA
    tensor; gloo: CPU tensor)."""

    def __init__(self, device=None):
        import torch
        import torch.distributed as dist
        self.torch, self.dist = torch, dist
        self.device = device if device is not None else (
            torch.device("cuda", torch.cuda.current_device()) if dist.get_backend() == "nccl" else torch.device("cpu"))

        self.pending = []

    def send(self, arr, dst):
        t = self.torch.from_numpy(np.ascontiguousarray(arr)).to(self.device)
        self.pending.append((self.dist.isend(t, dst), t))

    def flush(self):
        for h, _ in self.pending:
            h.wait()
        self.pending = []

    def recv(self, shape, src):
        t = self.torch.empty(shape, dtype=self.torch.float64, device=self.device)
        self.dist.recv(t, src)
        return t.cpu().numpy()

    def allreduce_sum(self, arr):
        t = self.torch.from_numpy(np.ascontiguousarray(arr, dtype=np.float64)).to(self.device)
        self.dist.all_reduce(t)
        return t.cpu().numpy()


class LocalComm:
    """In-process mailbox for emulating several ranks on one device (tests): ranks are stepped in
    an order that makes every recv find its message."""

    def __init__(self):
        self.box = {}

    def view(self, rank):
        outer = self

        class _V:
            def send(self, arr, dst):
                outer.box[(rank, dst)] = np.array(arr, copy=True)

            def recv(self, shape, src):
                return outer.box.pop((src, rank))

            def flush(self):
                pass
        return _V()


def initial_full_path(spec, make_engine):
    """Every rank initialises the WHOLE record identically (same seed => same Philox streams) and
    keeps its slice: [C][S+1][d] with point 0 = the starting point."""
    e = make_engine(spec, 0)
    e.init_paths()
    J = len(np.asarray(spec["N"]))
    seg = e.download_segment(0, J)
    x0 = np.asarray(spec["x0"], float)[:, 0:1, :]
    e.close()
    return np.concatenate([x0, seg], axis=1)
