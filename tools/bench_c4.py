"""BASELINE config 4: Lorenz-63, one long recording, chequerboard blocks sharded over the GPUs of
one box with NCCL halo/endpoint exchange.  Launch with torchrun (one rank per GPU) or plain python.

  python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 tools/bench_c4.py --obs 10000
"""
import argparse, json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import torch.distributed as dist
import dmcmc_pkg

pkg = dmcmc_pkg.load()
from diffusionmcmc_jl_b200 import parallel
from diffusionmcmc_jl_b200.engine import Engine

ap = argparse.ArgumentParser()
ap.add_argument("--obs", type=int, default=10000)
ap.add_argument("--chains", type=int, default=1)
ap.add_argument("--half-len", type=int, default=5)
ap.add_argument("--iters", type=int, default=50)
ap.add_argument("--check", action="store_true", help="compare with the unsharded run (small --obs)")
ap.add_argument("--host-exchange", action="store_true", help="stage the halo messages through host memory (the gloo-testable path)")
args = ap.parse_args()

rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
local = int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
spec = pkg.problems.l63_spec(n_obs=args.obs, C=args.chains, dt=1e-3, rho=0.9)
mk = lambda sp, off: Engine(sp, device=local, seed=9, interval_offset=off)
full0 = parallel.initial_full_path(spec, mk)
s = parallel.BlockShardedSampler(spec, rank, world, args.half_len, mk, full0)
acc = []
if args.host_exchange:
    comm = parallel.TorchComm() if world > 1 else parallel.LocalComm().view(0)
    step = lambda it: acc.append(np.nanmean(parallel.iterate(s, comm, it)[0][0][2]))
else:
    halo = parallel.DeviceHalo(s)   # device-resident exchange, stream-ordered, no host sync per iteration
    step = halo.iterate
for it in range(3):
    step(it)
torch.cuda.synchronize()
if world > 1:
    dist.barrier()
s.engine.accept_counts()
acc = []
t0 = time.perf_counter()
for it in range(3, 3 + args.iters):
    step(it)
torch.cuda.synchronize()
if world > 1:
    dist.barrier()
dt = time.perf_counter() - t0
if not args.host_exchange:
    a_, p_ = s.engine.accept_counts()
    acc = [a_.sum() / max(1, p_.sum())]
own = s.owned_path()
steps = int(np.asarray(spec["N"]).sum()) * args.chains
if args.check:
    gathered = [None] * world
    if world > 1:
        dist.all_gather_object(gathered, own)
    else:
        gathered = [own]
    if rank == 0:
        from diffusionmcmc_jl_b200.problems import chequerboard_layout
        e = mk(spec, 0)
        J = len(spec["N"])
        e.upload_segment(0, J, full0[:, 1:, :])
        for it in range(3 + args.iters):
            for pat in (0, 1):
                e.set_layout(*chequerboard_layout([J], args.half_len, pat))
                e.impute(2 * it + pat)
        ref = e.download_segment(0, J)
        X = np.concatenate(gathered, axis=1)
        print("sharded == unsharded bitwise:", bool(np.array_equal(X, ref)), "max diff", float(np.abs(X - ref).max()))
if rank == 0:
    print(json.dumps({"config": "C4 Lorenz-63 blocks sharded", "n_gpus": world, "obs": args.obs, "chains": args.chains,
                      "half_len": args.half_len, "iters": args.iters, "exchange": "host" if args.host_exchange else "device", "ms_per_iter": dt / args.iters * 1e3,
                      "path_steps_per_sec": 2 * steps * args.iters / dt, "accept_rate": float(np.mean(acc))}))
if world > 1:
    dist.destroy_process_group()
