"""Block sharding over REAL GPUs with the collectives inside the library: one process and one context per
GPU, NCCL communicator from dmcmc_comm_init, end-point / halo exchange by dmcmc_halo_exchange (ncclSend / ncclRecv on
the context's stream), parameter step by dmcmc_param_loglik_sharded (all-reduce of the per-block ll deg, AND-reduce of
success).  No torch.distributed anywhere: this is the path a Julia host takes.

The sharded run must reproduce the single-context run BIT FOR BIT: accepted paths after imputation sweeps, the ll deg /
ll of every global block in a critical parameter step (accepted and rejected), and the paths after it.

Needs >= 2 GPUs (gpurun --gpus 2); skipped on a single-GPU box.  ref: src/run!_alterations.jl:5-26, :177-211."""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
H, ITERS, SEED = 2, 3, 123


def _ngpu():
    try:
        import torch
        return torch.cuda.device_count()
    except Exception:
        return 0


pytestmark = [pytest.mark.gpu, pytest.mark.skipif(_ngpu() < 2, reason="needs two GPUs (gpurun --gpus 2)")]


def _spec(pkg):
    return pkg.problems.l63_spec(n_obs=48, C=3, dt=2e-3, rho=0.9)


def _thetas(spec):
    th1, th2 = np.array(spec["theta"], float), np.array(spec["theta"], float)
    th1[1] += 0.3     # accepted
    th2[0] -= 0.2     # rejected
    return th1, th2


def _worker(rank, world, uid, q, p2p, fused):
    sys.path.insert(0, ROOT)
    import dmcmc_pkg
    pkg = dmcmc_pkg.load()
    from diffusionmcmc_jl_b200 import parallel
    from diffusionmcmc_jl_b200.engine import Engine
    spec = _spec(pkg)
    mk = lambda sp, off: Engine(sp, device=rank, seed=SEED, interval_offset=off)
    full0 = parallel.initial_full_path(spec, mk)
    s = parallel.BlockShardedSampler(spec, rank, world, H, mk, full0)
    halo = parallel.LibHalo(s, uid, p2p=p2p, fused=fused)
    th1, th2 = _thetas(spec)
    for it in range(ITERS):
        halo.iterate(it)
    p1 = halo.param_step(th1, lambda llp, ll, ok: True)
    for it in range(ITERS, 2 * ITERS):
        halo.iterate(it)
    p2 = halo.param_step(th2, lambda llp, ll, ok: False)
    halo.iterate(2 * ITERS)
    q.put((rank, s.owned_path(), p1[:3], p2[:3], halo.p2p))
    s.engine.comm_destroy()
    s.engine.close()


def _unsharded(pkg):
    from diffusionmcmc_jl_b200 import parallel
    from diffusionmcmc_jl_b200.engine import Engine
    spec = _spec(pkg)
    mk = lambda sp, off: Engine(sp, device=0, seed=SEED, interval_offset=off)
    full0 = parallel.initial_full_path(spec, mk)
    e = mk(spec, 0)
    J = len(spec["N"])
    e.upload_segment(0, J, full0[:, 1:, :])
    lay = [pkg.problems.chequerboard_layout([J], H, p) for p in (0, 1)]
    th1, th2 = _thetas(spec)

    def sweeps(it):
        for pat in (0, 1):
            e.set_layout(*lay[pat])
            e.impute(2 * it + pat)

    def pstep(th, accept):
        e.relayout(*lay[0])
        ll = e.get_ll().copy()
        llp, ok = e.param_loglik(th, critical=True)
        e.param_commit(accept)
        return llp, ll, ok
    for it in range(ITERS):
        sweeps(it)
    p1 = pstep(th1, True)
    for it in range(ITERS, 2 * ITERS):
        sweeps(it)
    p2 = pstep(th2, False)
    sweeps(2 * ITERS)
    X = e.download_segment(0, J)
    e.close()
    return X, p1, p2


@pytest.mark.parametrize("halo", ["nccl", "peer_memory", "peer_memory_fused"])
@pytest.mark.parametrize("world", [2, 4, 8])
def test_block_sharded_sweeps_and_parameter_step_in_library_nccl(pkg, world, halo):
    """halo = nccl: end-point exchange by ncclSend / ncclRecv; peer_memory: by stores into the neighbour's CUDA-IPC mailbox
    and sequence flags (dmcmc_comm_enable_p2p), in a kernel of its own; peer_memory_fused: the same stores and flags from
    inside the sweep kernels (dmcmc_impute_exchange_async: the edge block's warps push after their accept decision and
    pull before they read).  All three must reproduce one GPU bit for bit."""
    if _ngpu() < world:
        pytest.skip(f"needs {world} GPUs")
    import torch.multiprocessing as mp
    from diffusionmcmc_jl_b200.engine import Engine
    Xref, r1, r2 = _unsharded(pkg)
    uid = Engine.comm_unique_id()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, uid, q, halo != "nccl", halo == "peer_memory_fused")) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted([q.get(timeout=600) for _ in procs], key=lambda t: t[0])
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    assert all(r[4] == (halo != "nccl") for r in res), "peer-memory mailboxes could not be set up (CUDA IPC)"
    X = np.concatenate([r[1] for r in res], axis=1)
    assert np.array_equal(X, Xref), f"sharded over {world} GPUs != unsharded: max diff {np.abs(X - Xref).max()}"
    for rank, _, p1, p2, _ in res:
        for got, ref, what in ((p1, r1, "accepted"), (p2, r2, "rejected")):
            assert np.array_equal(got[0], ref[0]), f"rank {rank}: ll deg of the {what} parameter step differs from the unsharded run"
            assert np.array_equal(got[1], ref[1]), f"rank {rank}: ll of the {what} parameter step differs"
            assert np.array_equal(got[2], ref[2])
    print(f"block sharding over {world} GPUs (in-library collectives, halo by {halo}): paths, per-block ll deg / ll, success bitwise equal to one GPU")
