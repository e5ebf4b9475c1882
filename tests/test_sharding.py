"""Multi-GPU host logic: chain sharding and block sharding with halo exchange.
CPU: world_size-2 gloo processes drive the CPU oracle through the same BlockShardedSampler the
GPU path uses.  GPU: several ranks emulated on one device must equal the single-context run."""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
H, ITERS, SEED = 2, 3, 77


def _spec(pkg, C=2):
    return pkg.problems.fhn_spec(n_obs=16, C=C, dt=0.02, rho=0.8)


def _reference_run(make_engine, spec, pkg):
    """Unsharded run: global chequerboard layouts, one context."""
    from diffusionmcmc_jl_b200 import parallel
    full0 = parallel.initial_full_path(spec, make_engine)
    e = make_engine(spec, 0)
    J = len(spec["N"])
    e.upload_segment(0, J, full0[:, 1:, :])
    hist = []
    for it in range(ITERS):
        for pat in (0, 1):
            e.set_layout(*pkg.problems.chequerboard_layout([J], H, pat))
            e.set_block_mask(None)
            hist.append(e.impute(2 * it + pat))
    return e.download_segment(0, J), hist


def _sharded_local(make_engine, spec, world):
    from diffusionmcmc_jl_b200 import parallel
    full0 = parallel.initial_full_path(spec, make_engine)
    ss = [parallel.BlockShardedSampler(spec, r, world, H, make_engine, full0) for r in range(world)]
    box = parallel.LocalComm()
    comms = [box.view(r) for r in range(world)]
    C, d = ss[0].engine.C, ss[0].engine.d
    hists = [[] for _ in ss]
    for it in range(ITERS):
        for s, h in zip(ss, hists):
            h.append(s.sweep(0, 2 * it))
        for s, c in zip(ss, comms):
            parallel.post_a(s, c)
        for s, c in zip(ss, comms):
            parallel.take_a(s, c, C, d)
        for s, h in zip(ss, hists):
            h.append(s.sweep(1, 2 * it + 1))
        for s, c in zip(ss, comms):
            parallel.post_b(s, c)
        for s, c in zip(ss, comms):
            parallel.take_b(s, c, C, d)
    return np.concatenate([s.owned_path() for s in ss], axis=1), hists, ss


def test_shard_helpers(pkg):
    from diffusionmcmc_jl_b200 import parallel
    assert parallel.shard_chains(10, 4) == [(0, 3), (3, 3), (6, 2), (8, 2)]
    assert parallel.shard_intervals(100, 4, 1) == [0, 24, 50, 74, 100]
    e = parallel.shard_intervals(10000, 8, 5)
    assert e[0] == 0 and e[-1] == 10000 and all(x % 10 == 0 for x in e[1:-1])
    with pytest.raises(ValueError):
        parallel.shard_intervals(6, 4, 1)
    # the local layouts of all ranks tile the global chequerboard exactly (bit-exact partition)
    n, h, world = 40, 2, 3
    edges = parallel.shard_intervals(n, world, h)
    for pat in (0, 1):
        g1, g2, gp = pkg.problems.chequerboard_layout([n], h, pat)
        got = []
        for r in range(world):
            a, b = edges[r], edges[r + 1]
            b_ext = min(b + h, n) if r < world - 1 else b
            i1, i2, pin, act = parallel.local_layouts(a, b, b_ext, n, h, r)[pat]
            for x, y, p_, on in zip(i1, i2, pin, act):
                if on:
                    got.append((x + a, y + a, p_))
        assert sorted(got) == sorted(zip(g1, g2, gp))


def test_block_sharding_oracle_in_process(pkg, orc):
    """Sharded (2 and 3 emulated ranks) == unsharded, bit for bit, on the CPU oracle."""
    from helpers import OracleEngine
    spec = _spec(pkg)
    mk = lambda sp, off: OracleEngine(orc, sp, interval_offset=off, seed=SEED)
    Xref, href = _reference_run(mk, spec, pkg)
    for world in (2, 3):
        X, hists, ss = _sharded_local(mk, spec, world)
        assert np.array_equal(X, Xref), np.abs(X - Xref).max()


def _gloo_worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import torch.distributed as dist
    import dmcmc_pkg
    pkg = dmcmc_pkg.load()
    from diffusionmcmc_jl_b200 import parallel
    from helpers import OracleEngine
    from oracle import oracle as orc
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    spec = _spec(pkg)
    mk = lambda sp, off: OracleEngine(orc, sp, interval_offset=off, seed=SEED)
    full0 = parallel.initial_full_path(spec, mk)
    s = parallel.BlockShardedSampler(spec, rank, world, H, mk, full0)
    comm = parallel.TorchComm()
    lls = []
    for it in range(ITERS):
        (oa, _), (ob, _) = parallel.iterate(s, comm, it)
        # the parameter step's global log-likelihood: all-reduce of the local block sums
        lls.append(comm.allreduce_sum(np.array([np.nansum(ob[1])]))[0])
    q.put((rank, s.owned_path(), lls))
    dist.barrier()
    dist.destroy_process_group()


def test_block_sharding_gloo_world2(pkg, orc):
    """Two real processes over gloo: halo exchange + ll all-reduce; union equals the unsharded run."""
    import torch.multiprocessing as mp
    from helpers import OracleEngine
    spec = _spec(pkg)
    mk = lambda sp, off: OracleEngine(orc, sp, interval_offset=off, seed=SEED)
    Xref, href = _reference_run(mk, spec, pkg)
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_gloo_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted([q.get(timeout=120) for _ in procs], key=lambda t: t[0])
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    X = np.concatenate([r[1] for r in res], axis=1)
    assert np.array_equal(X, Xref)
    ll_ref = [np.nansum(href[2 * it + 1][1]) for it in range(ITERS)]
    assert np.allclose(res[0][2], ll_ref, rtol=1e-12) and np.allclose(res[1][2], ll_ref, rtol=1e-12)


@pytest.mark.gpu
def test_block_sharding_gpu_emulated_ranks(pkg):
    """Same statement on the CUDA path: 2 and 4 contexts on one device vs one context."""
    from diffusionmcmc_jl_b200.engine import Engine
    spec = pkg.problems.l63_spec(n_obs=48, C=3, dt=2e-3, rho=0.9)
    mk = lambda sp, off: Engine(sp, seed=SEED, interval_offset=off)
    global H
    Xref, href = _reference_run(mk, spec, pkg)
    for world in (2, 4):
        X, hists, ss = _sharded_local(mk, spec, world)
        assert np.array_equal(X, Xref), np.abs(X - Xref).max()
        for s in ss:
            s.engine.close()


@pytest.mark.gpu
def test_device_halo_exchange_emulated_ranks(pkg):
    """The device-resident exchange (packed segments in CUDA tensors, dmcmc_segment_device /
    dmcmc_set_start_device / dmcmc_impute_async, no host staging): 3 contexts on one device, the
    NCCL transfer replaced by a device-to-device copy, against the unsharded run, bit for bit."""
    import torch
    from diffusionmcmc_jl_b200 import parallel
    from diffusionmcmc_jl_b200.engine import Engine
    spec = pkg.problems.l63_spec(n_obs=48, C=3, dt=2e-3, rho=0.9)
    mk = lambda sp, off: Engine(sp, seed=SEED, interval_offset=off)
    Xref, _ = _reference_run(mk, spec, pkg)
    world = 3
    full0 = parallel.initial_full_path(spec, mk)
    ss = [parallel.BlockShardedSampler(spec, r, world, H, mk, full0) for r in range(world)]
    hs = [parallel.DeviceHalo(s) for s in ss]
    for it in range(ITERS):
        for h in hs:
            h.sweep_a(it)
        torch.cuda.synchronize()
        for r in range(world - 1):
            hs[r].from_right.copy_(hs[r + 1].to_left)
        torch.cuda.synchronize()
        for h in hs:
            h.take_a()
            h.sweep_b(it)
        torch.cuda.synchronize()
        for r in range(1, world):
            hs[r].from_left.copy_(hs[r - 1].to_right)
        torch.cuda.synchronize()
        for h in hs:
            h.take_b()
    X = np.concatenate([s.owned_path() for s in ss], axis=1)
    assert np.array_equal(X, Xref), np.abs(X - Xref).max()
    for s in ss:
        s.engine.close()


@pytest.mark.gpu
def test_chain_sharding_gpu(pkg):
    from diffusionmcmc_jl_b200 import parallel
    from diffusionmcmc_jl_b200.engine import Engine
    spec = pkg.problems.fhn_spec(n_obs=8, C=10, dt=0.01)
    whole = Engine(spec, seed=5)
    whole.init_paths()
    parts = [Engine(spec, seed=5, chain_offset=o, n_chains=c) for o, c in parallel.shard_chains(10, 3)]
    for p_ in parts:
        p_.init_paths()
    whole.run_sweeps(0, 6, 1)
    for p_ in parts:
        p_.run_sweeps(0, 6, 1)
    Xw, _ = whole.download_state(0)
    Xp = np.concatenate([p_.download_state(0)[0] for p_ in parts])
    assert np.array_equal(Xw, Xp)
