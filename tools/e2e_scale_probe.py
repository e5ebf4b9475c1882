"""Where does the end-to-end time go at N GPUs?  Run under torchrun: the C3 sweep loop through dmcmc_run_sweeps with
(a) all three history rows, (b) the accept flags only, (c) no histories (rho schedule only), max over ranks."""
import os, sys, time, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import torch.distributed as dist
import dmcmc_pkg
pkg = dmcmc_pkg.load()
from diffusionmcmc_jl_b200.engine import Engine, _dp, _bp
from diffusionmcmc_jl_b200.problems import chequerboard_layout

rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
spec = pkg.problems.fhn_spec(n_obs=100, C=1024, dt=1e-3, rho=0.96, seed=1)
e = Engine(spec, device=local, seed=1, chain_offset=1024 * rank)
e.init_paths()
n = 500
stride = max(len(chequerboard_layout(e.rec_nobs, 1, p)[0]) for p in (0, 1))
rho = np.full((n, e.J), 0.96)
llp, ll = np.zeros((n, e.C, stride)), np.zeros((n, e.C, stride))
acc = np.zeros((n, e.C, stride), np.uint8)
e.run_sweeps(0, 20, 1)
out = {}
for name, args in [("all", (_dp(rho), stride, _dp(llp), _dp(ll), _bp(acc))), ("acc_only", (_dp(rho), stride, None, None, _bp(acc))),
                   ("no_hist", (_dp(rho), 0, None, None, None)), ("device", (None, 0, None, None, None))]:
    for rep in range(2):
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        e._ck(e.lib.dmcmc_run_sweeps(e.h, 1000 * (1 + rep), n, 1, 0, *args))
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
    t = torch.tensor([dt], device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    out[name] = round(float(t.item()) / n * 1e6, 1)
if rank == 0:
    print(json.dumps({"n_gpus": world, "us_per_sweep": out, "cpus": os.cpu_count()}))
