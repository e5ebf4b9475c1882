#!/usr/bin/env python
"""bench.py -- bridge path-steps/sec of the imputation hot path (BASELINE.json metric).

Headline workload (N=1 and per GPU for N>1, weak scaling): BASELINE.json configs[2] = C3 -- FitzHugh-Nagumo,
1024 independent chains x 100 observations, dt = 1e-3 (100 Euler steps per inter-observation interval =>
1.024e7 path-steps per sweep per GPU), pCN rho = 0.96, chequerboard blocking with block_half_len = 1 (each sweep =
layout switch [accepted noise + ll recovered from the accepted path] + imputation of every (chain, block), fused in
one kernel), synthetic observations (NumPy PCG64 seed 1), Philox noise in-register.

A "step" is one PathImputation update (one sweep over every chain and block).

  value  : path-steps/s, state resident in HBM, K sweeps issued back to back; per rank the time is taken with CUDA
           events on the stream the kernels are launched on, MAX over ranks (no collective inside the timed region).
  e2e    : the same K sweeps through the reference-facing call (C ABI dmcmc_run_sweeps, what `run!` drives) with
           PINNED HOST buffers: the pCN rho schedule goes host->device and the ll deg / ll / accepted history rows of
           every sweep (src/run!_alterations.jl:57-75) come device->host inside the timed call; wall clock around the
           call (it returns when the last row has landed), MAX over ranks.
  configs: one entry per other named shape of BASELINE.json (C1, C2, C4, C5), same measurements at their own sizes
           (--configs c3 skips them).
  --impl reference : the CPU restatement of the reference path (oracle/, OpenMP over units) on the box's host cores
           on a bounded sample of the same workload.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

N_OBS, N_CHAINS, DT, RHO, HALF_LEN = 100, 1024, 1e-3, 0.96, 1
# FHN, fused layout switch (the timed mode): read the accepted X (8d) + write X deg (8d) = 32 B per path-step; the
# stored-noise mode (read W 8m + write W deg 8m + write X deg 8d) is 32 B as well for d = 2, m = 1
BYTES_PER_STEP = 32
METRIC, UNIT = "bridge path-steps/sec", "path-steps/s"
CPU_SAMPLE_CHAINS, CPU_SAMPLE_SWEEPS = 512, 24  # ~1.2e8 path-steps: 10-30 s of CPU work over the host cores


def profiled_traffic():
    """dram__bytes_read.sum + dram__bytes_write.sum of the dominant kernel, per launch, from the latest committed
    ncu --set full capture of this workload (profiles/r*_traffic.json); (None, None) if absent."""
    import glob
    for path in sorted(glob.glob(os.path.join(ROOT, "profiles", "r*_traffic.json")), reverse=True):
        try:
            with open(path) as f:
                return int(json.load(f)["traffic_bytes_per_launch"]), os.path.basename(path)
        except Exception:
            continue
    return None, None


def measured_peak():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured"
    except Exception:
        return 6650.0, "fallback"


class ClockSampler:
    """nvidia-smi clocks/throttle reasons DURING the timed region (B200_PROFILING.md recipe)."""
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}",
                 "--format=csv,noheader,nounits", "-lms", "100"],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.th = threading.Thread(target=self._read, daemon=True)
            self.th.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm = [float(r[0]) for r in self.rows if len(r) >= 6 and r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in self.rows if len(r) >= 6 and r[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({names[i] for r in self.rows if len(r) >= 6 for i in range(4)
                          if r[2 + i].lower().startswith("active")})
        return {"sm_mhz": float(np.median(sm)) if sm else None,
                "sm_max_mhz": max(mx) if mx else None, "reasons": reasons,
                "samples": len(sm)}


def make_spec(chains):
    import dmcmc_pkg
    pkg = dmcmc_pkg.load()
    return pkg, pkg.problems.fhn_spec(n_obs=N_OBS, C=chains, dt=DT, rho=RHO, seed=1)


def cpu_oracle_rate(spec_small, sweeps, nthreads, pkg, half_len=HALF_LEN, layouts=None):
    """path-steps/s of the CPU restatement on a bounded sample (same algorithm, same layouts)."""
    from oracle import oracle
    o = oracle.Oracle(spec_small, nthreads=nthreads)
    o.init_paths(seed=1)
    if layouts is None:
        layouts = [pkg.problems.chequerboard_layout(o.rec_nobs, half_len, p) for p in ((0, 1) if half_len > 0 else (0,))]
    o.relayout(*layouts[0])
    o.impute(seed=1, it=0)  # warm
    t0 = time.perf_counter()
    for it in range(sweeps):
        if len(layouts) > 1:
            o.relayout(*layouts[(it + 1) % len(layouts)])
        o.impute(seed=1, it=1 + it)
    dt = time.perf_counter() - t0
    return o.C * o.S * sweeps / dt, dt


def run_reference(args):
    """`--impl reference`: the reference's CPU path (oracle port; Julia cannot run here) with all
    host threads, on a bounded sample of the same workload."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    # bounded sample: probe the rate on 64 chains, then size the per-step sample so that the whole
    # --steps K run takes about a minute of wall time (between 16 and 512 of the 1024 chains)
    pkg, probe = make_spec(64)
    rate0, _ = cpu_oracle_rate(probe, max(1, min(args.warmup, 2)), cores, pkg)
    per_chain = N_OBS * 100
    sample_chains = int(min(512, max(16, (60.0 * rate0 / (args.steps * per_chain)) // 16 * 16)))
    pkg, spec = make_spec(sample_chains)
    rate, secs = cpu_oracle_rate(spec, args.steps, cores, pkg)
    ms = secs / args.steps * 1e3
    sample = (f"{sample_chains} of {N_CHAINS} chains x {N_OBS} obs x 100 steps, {args.steps} sweeps "
              f"(relayout + impute), OpenMP over (chain, block) units")
    line = {
        "impl": "reference", "metric": METRIC, "value": rate, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": workload_config(),
        "cpu_baseline": {"value": rate, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": rate, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "note": "CPU restatement (oracle/dmcmc_oracle.c) of the reference path; Julia run! cannot "
                "execute in this image; throughput is per-step-rate of the sample, not of 1024 chains",
    }
    print(json.dumps(line), flush=True)


def workload_config():
    return {"workload": "C3: FitzHugh-Nagumo d=2 m=1, 1024 chains x 100 obs per GPU, dt=1e-3 "
                        "(1.024e7 path-steps/sweep/GPU), pCN rho=0.96, chequerboard blocking "
                        "half_len=1 (layout switch fused into the imputation kernel), Philox noise",
            "chains_per_gpu": N_CHAINS, "n_obs": N_OBS, "steps_per_interval": 100,
            "block_half_len": HALF_LEN, "parallelism": "chains sharded, no collective",
            "l2": "state 0.5 GB per GPU >> 126 MB L2 (inputs larger than L2)"}


# ------------------------------------------------------------------------------------------------------------------
# the other named shapes (BASELINE.json configs[0], [1], [3], [4]): same measurements at their own sizes
# ------------------------------------------------------------------------------------------------------------------
def _time_device(eng, fn):
    eng.timer_start()
    fn()
    return eng.timer_stop()


def _roof(bytes_per_step, steps_per_sweep, ms_per_sweep):
    peak, which = measured_peak()
    ach = bytes_per_step * steps_per_sweep / (ms_per_sweep * 1e-3) / 1e9
    return {"bound": "hbm", "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak, "peak_source": which,
            "algorithmic_bytes_per_step": bytes_per_step}


def _sweep_config(pkg, name, spec, half_len, K, W, local_rank, bytes_per_step, store_noise=True, cpu=None, seed=1,
                  chain_offset=0):
    """A smoothing chain under chequerboard blocking (half_len = 0: none) through dmcmc_run_sweeps: device-resident and
    with pinned host histories.  Returns one config entry."""
    from diffusionmcmc_jl_b200.engine import Engine
    t0 = time.perf_counter()
    eng = Engine(spec, device=local_rank, seed=seed, store_noise=store_noise, chain_offset=chain_offset)
    eng.init_paths()
    setup_s = time.perf_counter() - t0
    steps = eng.C * eng.S
    first = [0]

    def run(n, hist=None, rho=None):
        eng.run_sweeps(1000 + first[0], n, half_len, first[0] & 1, rho, hist)
        first[0] += n
    t0 = time.perf_counter()
    run(max(W, 2))
    tables_s = time.perf_counter() - t0
    eng.launch_count()
    ms = _time_device(eng, lambda: run(K)) / K
    launches = eng.launch_count()
    nb = max(len(pkg.problems.chequerboard_layout(eng.rec_nobs, half_len, p)[0]) for p in (0, 1))
    hist = (eng.pinned((K, eng.C, nb), np.float64), eng.pinned((K, eng.C, nb), np.float64), eng.pinned((K, eng.C, nb), np.uint8))
    rho = eng.pinned((K, eng.J), np.float64)
    rho[...] = float(np.asarray(spec["rho"]).ravel()[0])
    run(max(W, 2), tuple(h[:max(W, 2)] for h in hist), rho[:max(W, 2)])
    t0 = time.perf_counter()
    run(K, hist, rho)
    e2e_ms = (time.perf_counter() - t0) / K * 1e3
    acc, prop = eng.accept_counts()
    out = {"workload": name, "n_chains": eng.C, "n_obs": eng.J, "path_steps_per_sweep": steps, "block_half_len": half_len,
           "units_per_sweep": eng.C * nb, "value": steps / ms * 1e3, "unit": UNIT, "ms_per_sweep": ms,
           "sweeps_per_sec": 1e3 / ms, "gpu_launches": launches,
           "e2e": {"value": steps / e2e_ms * 1e3, "unit": UNIT, "ms_per_sweep": e2e_ms, "h2d_bytes_per_step": eng.J * 8,
                   "d2h_bytes_per_step": eng.C * nb * 17},
           "roofline": _roof(bytes_per_step, steps, ms), "accept_rate": float(acc.sum() / max(1, prop.sum())),
           "device_GB": eng.device_bytes() / 1e9, "setup_s": {"engine+init_paths": setup_s, "tables+first sweeps": tables_s}}
    if cpu:
        out["cpu_baseline"] = cpu()
    return out, eng


def config_c1(pkg, K, W, local_rank):
    """C1: FitzHugh-Nagumo, 50 obs, dt = 1e-3, rho = 0.96, ONE chain, no blocking -- the reference's own tutorial run
    (docs/src/tutorials/simple_inference.md:77-105: 1e3 iterations in 'about 0.5 sec' = 2e3 it/s on its author's
    laptop).  One (chain, block) unit: a single warp's latency; not an HBM-bound shape."""
    spec = pkg.problems.fhn_spec(n_obs=50, C=1, dt=1e-3, rho=0.96, seed=1)

    def cpu():
        cores = 1  # one chain, no blocks: nothing to split over threads (the reference's loop is serial as well)
        rate, secs = cpu_oracle_rate(spec, 400, cores, pkg, half_len=0)
        return {"value": rate, "unit": UNIT, "cores": cores, "kind": "port", "mcmc_iters_per_sec": 400 / secs,
                "sample": "the whole config: 1 chain x 50 obs x 100 steps, 400 sweeps, one thread"}
    out, eng = _sweep_config(pkg, "C1: FitzHugh-Nagumo 1 chain x 50 obs, dt=1e-3, rho=0.96, unblocked (reference tutorial shape)",
                             spec, 0, max(K, 200), W, local_rank, 32, cpu=cpu)
    out["mcmc_iters_per_sec"] = 1e3 / out["e2e"]["ms_per_sweep"]   # one iteration = one imputation sweep, rows on the host
    out["note"] = "one unit of 5000 serial Euler steps: warp-per-unit kernel, latency-bound by construction (SURVEY 8d)"
    eng.close()
    return out


def config_c2(pkg, K, W, local_rank):
    """C2: Lotka-Volterra, 200 obs, ONE chain, 8 interval blocks of 25 (alternating with the half-shifted pattern),
    smoothing + random-walk update of one drift parameter.  One MCMC iteration = imputation sweep + parameter step
    (critical: aux law, backward filter, solve with the fixed accepted noise)."""
    from diffusionmcmc_jl_b200.engine import Engine
    P = pkg.problems
    spec = P.lv_spec(n_obs=200, C=1, dt=0.01, rho=0.9, seed=2)
    J = 200
    a1 = np.arange(0, J, 25, dtype=np.int32)
    lay_a = (a1, a1 + 24, (a1 + 24 < J - 1).astype(np.int32))
    b1 = np.concatenate([[0], np.arange(12, J, 25)]).astype(np.int32)
    b2 = np.concatenate([b1[1:] - 1, [J - 1]]).astype(np.int32)
    lays = [lay_a, (b1, b2, (b2 < J - 1).astype(np.int32))]
    eng = Engine(spec, device=local_rank, seed=2)
    eng.init_paths()
    steps, n = eng.C * eng.S, max(K, 200)

    def sweeps(it0, cnt):
        for it in range(it0, it0 + cnt):
            eng.set_layout(*lays[it & 1])
            eng.impute_async(it)
    sweeps(0, max(W, 4))
    eng.launch_count()
    ms = _time_device(eng, lambda: sweeps(100, n)) / n
    launches = eng.launch_count()
    # full MCMC iterations through the per-update calls run! makes: update!(::PathImputation) then the parameter update
    rng = np.random.default_rng(2)
    theta = np.array(spec["theta"], float)
    n_it, acc_p = max(50, min(n, 200)), 0
    t0 = time.perf_counter()
    for it in range(n_it):
        eng.set_layout(*lays[it & 1])
        eng.impute(5000 + it)
        th = theta.copy()
        th[0] *= np.exp(rng.uniform(-0.02, 0.02))
        llp, ok = eng.param_loglik(th, critical=True)
        ll = eng.get_ll()
        accept = bool(ok.all()) and rng.exponential() > -(llp.sum() - ll.sum() + np.log(th[0]) - np.log(theta[0]))
        eng.param_commit(accept)
        if accept:
            theta, acc_p = th, acc_p + 1
    it_ms = (time.perf_counter() - t0) / n_it * 1e3

    def cpu():
        rate, secs = cpu_oracle_rate(spec, 200, min(8, os.cpu_count() or 1), pkg, layouts=lays)
        return {"value": rate, "unit": UNIT, "cores": min(8, os.cpu_count() or 1), "kind": "port",
                "sample": "the whole config: 1 chain x 200 obs x 10 steps, 200 sweeps (relayout + impute), OpenMP over the 8-9 blocks"}
    out = {"workload": "C2: Lotka-Volterra 1 chain x 200 obs, dt=0.01, 8 blocks of 25 intervals alternating with the "
                       "half-shifted pattern, random-walk update of alpha",
           "n_chains": 1, "n_obs": J, "path_steps_per_sweep": steps, "units_per_sweep": 8, "value": steps / ms * 1e3,
           "unit": UNIT, "ms_per_sweep": ms, "sweeps_per_sec": 1e3 / ms, "gpu_launches": launches,
           "mcmc_iters_per_sec": 1e3 / it_ms,
           "e2e": {"value": steps / it_ms * 1e3, "unit": UNIT, "ms_per_iteration": it_ms,
                   "what": "update!(::PathImputation) with host history rows + critical parameter step (aux law, backward "
                           "filter, fixed-noise solve, accept, swap) per iteration",
                   "h2d_bytes_per_step": int(len(theta) * 8), "d2h_bytes_per_step": int(9 * 17 + 9 * 8 + 1)},
           "parameter_step_accept_rate": acc_p / n_it,
           "roofline": _roof(48, steps, ms), "cpu_baseline": cpu(),
           "note": "8-9 units of 250 steps: warp-per-unit kernel, latency-bound (SURVEY 8d)"}
    eng.close()
    return out


def config_c4(pkg, K, W, local_rank, rank, world):
    """C4: Lorenz-63, ONE chain, 10^4 obs (10 steps per interval), blocks of 2h = 10 intervals; N = 1: the whole record
    on one GPU; N > 1: contiguous interval ranges per GPU, halo / end-point exchange over NCCL after every sweep
     -- strong scaling of one chain."""
    P = pkg.problems
    spec = P.l63_spec(n_obs=10_000, C=1, dt=1e-3, rho=0.9, seed=4)
    steps = int(np.asarray(spec["N"]).sum())
    if world == 1:
        def cpu():
            cores = os.cpu_count() or 1
            rate, secs = cpu_oracle_rate(spec, 40, cores, pkg, half_len=5)
            return {"value": rate, "unit": UNIT, "cores": cores, "kind": "port",
                    "sample": "the whole config: 1 chain x 10^4 obs x 10 steps, 40 sweeps (relayout + impute), OpenMP over ~1000 blocks"}
        out, eng = _sweep_config(pkg, "C4: Lorenz-63 1 chain x 10^4 obs, dt=1e-3 (10 steps per interval), chequerboard half_len=5",
                                 spec, 5, max(K, 100), W, local_rank, 72, cpu=cpu, seed=4)
        # the parameter step of this shape (random-walk update of theta_2, critical)
        eng.relayout(*P.chequerboard_layout(eng.rec_nobs, 5, 0))
        th = np.array(spec["theta"], float)
        eng.param_loglik(th * (1 + 1e-4), critical=True)
        eng.param_commit(False)
        t0 = time.perf_counter()
        for i in range(20):
            eng.param_loglik(th * (1 + 1e-4 * (i + 2)), critical=True)
            eng.param_commit(False)
        out["parameter_step_ms"] = (time.perf_counter() - t0) / 20 * 1e3
        out["mcmc_iters_per_sec"] = 1e3 / (out["e2e"]["ms_per_sweep"] + out["parameter_step_ms"])
        eng.close()
        return out
    import torch
    import torch.distributed as dist
    from diffusionmcmc_jl_b200 import parallel
    from diffusionmcmc_jl_b200.engine import Engine
    mk = lambda sp, off: Engine(sp, device=local_rank, seed=4, interval_offset=off)
    full0 = parallel.initial_full_path(spec, mk)
    s = parallel.BlockShardedSampler(spec, rank, world, 5, mk, full0)
    # the collectives are the library's own (dmcmc_comm_init / dmcmc_halo_exchange / dmcmc_param_loglik_sharded);
    # torch.distributed only carries the 128-byte NCCL id to the other ranks
    uid = [Engine.comm_unique_id() if rank == 0 else None]
    dist.broadcast_object_list(uid, src=0)
    halo = parallel.LibHalo(s, uid[0], p2p=False)
    n = max(K, 1000)   # 100 iterations (7 ms) scattered by +-5 us from run to run

    def timed_iterations(first):
        for it in range(first, first + max(W, 3)):
            halo.iterate(it)
        torch.cuda.synchronize()
        dist.barrier()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for it in range(first + 10, first + 10 + n):
            halo.iterate(it)
        torch.cuda.synchronize()
        return time.perf_counter() - t0
    # A/B in the same run: the exchange by ncclSend / ncclRecv first, then by peer-memory mailboxes (NVLink stores +
    # sequence flags from the pack / unpack kernels, dmcmc_comm_enable_p2p); the line's value is the second when it is on
    dt_nccl = timed_iterations(0)
    p2p_on = os.environ.get("DMCMC_P2P", "1") != "0" and s.engine.comm_enable_p2p((s.h + 1) * s.N_max * s.engine.d * 8 * s.engine.C)
    halo.p2p = p2p_on
    halo.fused = False
    dt_p2p = timed_iterations(10_000) if p2p_on else dt_nccl
    # third: the same stores and flags from inside the sweep kernels (dmcmc_impute_exchange_async)
    fused_on = p2p_on and os.environ.get("DMCMC_P2P_FUSE", "1") != "0"
    halo.fused = fused_on
    dt = timed_iterations(20_000) if fused_on else dt_p2p
    halo.fused = False   # the part timings below call the exchange and the sweeps on their own
    # what the two parts cost on their own: the exchanges alone (same messages, nothing computed in between) and the
    # sweeps alone (no exchange; the halo goes stale, which does not change the work)
    e = s.engine
    torch.cuda.synchronize()
    dist.barrier()
    t0 = time.perf_counter()
    for it in range(n):
        e.halo_exchange(0, halo.n_first, halo.n_halo, halo.left_edge_steps)
        e.halo_exchange(1, halo.n_first, halo.n_halo, halo.left_edge_steps)
    torch.cuda.synchronize()
    dt_x = (time.perf_counter() - t0) / n
    dist.barrier()
    t0 = time.perf_counter()
    for it in range(n):
        s.sweep_async(0, 2 * (500 + it))
        s.sweep_async(1, 2 * (500 + it) + 1)
    torch.cuda.synchronize()
    dt_k = (time.perf_counter() - t0) / n
    # parameter step: every rank's blocks, all-reduce of the per-block ll deg, one decision, commit
    th = np.array(spec["theta"], float)
    halo.param_step(th * (1 + 1e-4), lambda llp, ll, ok: False)
    torch.cuda.synchronize()
    dist.barrier()
    t1 = time.perf_counter()
    for i in range(20):
        halo.param_step(th * (1 + 1e-4 * (i + 2)), lambda llp, ll, ok: False)
    dtp = (time.perf_counter() - t1) / 20
    tt = torch.tensor([dt, dtp, dt_x, dt_k, dt_nccl, dt_p2p], device="cuda", dtype=torch.float64)
    dist.all_reduce(tt, op=dist.ReduceOp.MAX)
    dt, dtp, dt_x, dt_k, dt_nccl, dt_p2p = (float(x) for x in tt)
    out = {"workload": "C4: Lorenz-63 1 chain x 10^4 obs, chequerboard half_len=5, interval blocks sharded over the GPUs, "
                       "end-point / halo exchange after every sweep (inside the library)",
           "n_chains": 1, "n_obs": 10_000, "n_gpus": world, "scaling": "strong", "path_steps_per_sweep": steps,
           "value": 2 * steps * n / dt, "unit": UNIT, "ms_per_iteration": dt / n * 1e3,
           "parameter_step_ms": dtp * 1e3, "mcmc_iters_per_sec": 1.0 / (dt / n / 2 + dtp),
           "exchanges_alone_ms_per_iteration": dt_x * 1e3, "sweeps_alone_ms_per_iteration": dt_k * 1e3,
           "halo": ("peer memory, fused into the sweep kernels (the edge block's warps push after their accept decision and pull "
                    "before they read)" if fused_on else
                    "peer memory (CUDA IPC mailboxes, NVLink stores + flags from the exchange kernel)" if p2p_on else "ncclSend / ncclRecv"),
           "ms_per_iteration_with_nccl_halo": dt_nccl / n * 1e3,
           "ms_per_iteration_with_peer_memory_halo_in_its_own_kernel": dt_p2p / n * 1e3,
           "what": "one imputation iteration = sweep A + exchange + sweep B + exchange; wall clock, MAX over ranks; parameter "
                   "step = relayout + backward filter + fixed-noise solve + all-reduce of per-block ll deg + commit; "
                   "mcmc_iters_per_sec counts one sweep + one parameter step per iteration like the N=1 entry"}
    s.engine.comm_destroy()
    s.engine.close()
    return out


def config_c5(pkg, K, W, local_rank, rank, world, n_obs=4096, chains=256):
    """C5: Lorenz-96 d = 40, every other coordinate observed, 4096 obs x 100 steps (dt = 1e-4), 256 chains (256 / N per
    GPU), chequerboard half_len = 16, store_noise = 0 (the accepted noise is never stored: 2 x X only)."""
    from diffusionmcmc_jl_b200 import _lib
    P = pkg.problems
    c_loc = chains // world
    spec = P.l96_spec(n_obs=n_obs, C=c_loc, d=40, dt=1e-4, rho=0.9, seed=5)
    n = max(3, min(K, 6))

    def cpu():
        cores = os.cpu_count() or 1
        small = P.l96_spec(n_obs=64, C=max(8, cores), d=40, dt=1e-4, rho=0.9, seed=5)
        rate, secs = cpu_oracle_rate(small, 3, cores, pkg, half_len=16)
        return {"value": rate, "unit": UNIT, "cores": cores, "kind": "port",
                "sample": f"{max(8, cores)} chains x 64 obs x 100 steps, 3 sweeps (relayout + impute), OpenMP over (chain, block) units, {secs:.1f} s"}
    out, eng = _sweep_config(pkg, f"C5: Lorenz-96 d=40 p=20, {c_loc} chains x {n_obs} obs x 100 steps (dt=1e-4), chequerboard "
                                  f"half_len=16, store_noise=0", spec, 16, n, 1, local_rank, 960, store_noise=False,
                             cpu=cpu if rank == 0 and world == 1 else None, seed=5, chain_offset=rank * c_loc)
    # FP64 view: per path-step three 40 x 40 mat-vecs (H x deg, H x for the recovered noise / accepted ll, Psi' x_b)
    # = 9600 flop, plus ~30 flop per component for drifts, residual, mix, update
    flops = 2 * 3 * 40 * 40 + 30 * 40
    tf = out["value"] * flops / 1e12
    peak = _lib.load().dmcmc_fp64_peak(local_rank)
    tpeak = _lib.load().dmcmc_fp64_tensor_peak(local_rank)
    out["fp64"] = {"bound": "fp64", "achieved": tf, "peak": peak, "unit": "TFLOP/s", "frac": tf / peak if peak > 0 else None,
                   "flops_per_step": flops, "peak_source": "measured here: dependent-free DFMA streams on every SM (dmcmc_fp64_peak)",
                   "tensor_peak": tpeak, "frac_of_tensor_peak": tf / tpeak if tpeak > 0 else None,
                   "tensor_peak_source": "measured here: independent mma.sync m8n8k4.f64 (DMMA) chains on every SM "
                                         "(dmcmc_fp64_tensor_peak); the kernel's three 40 x 40 mat-vecs per step run as DMMA, "
                                         "DMMA and DFMA share one FP64 datapath on B200"}
    if world > 1:
        out["n_gpus"], out["scaling"] = world, "strong: 256 chains / N per GPU; value = all ranks' path-steps / slowest rank's time"
    eng.close()
    return out


def other_configs(pkg, which, args, rank, local_rank, world):
    import torch
    out = {}
    for name in which:
        t0 = time.perf_counter()
        try:
            if name == "c1" and rank == 0:
                out[name] = config_c1(pkg, args.steps, args.warmup, local_rank)
            elif name == "c2" and rank == 0:
                out[name] = config_c2(pkg, args.steps, args.warmup, local_rank)
            elif name == "c4":
                out[name] = config_c4(pkg, args.steps, args.warmup, local_rank, rank, world)
            elif name == "c5":
                r = config_c5(pkg, args.steps, args.warmup, local_rank, rank, world)
                if world > 1:   # chains sharded: the job's rate is the sum over ranks at the slowest rank's time
                    import torch.distributed as dist
                    tt = torch.tensor([r["ms_per_sweep"], r["e2e"]["ms_per_sweep"]], device="cuda", dtype=torch.float64)
                    dist.all_reduce(tt, op=dist.ReduceOp.MAX)
                    r["ms_per_sweep"], r["e2e"]["ms_per_sweep"] = float(tt[0]), float(tt[1])
                    r["value"] = r["path_steps_per_sweep"] * world / r["ms_per_sweep"] * 1e3
                    r["e2e"]["value"] = r["path_steps_per_sweep"] * world / r["e2e"]["ms_per_sweep"] * 1e3
                    r["fp64"]["achieved_all_gpus"] = r["value"] * r["fp64"]["flops_per_step"] / 1e12
                out[name] = r
        except Exception as ex:   # a secondary shape must not take the headline line down with it
            out[name] = {"error": f"{type(ex).__name__}: {ex}"}
        if name in out:
            out[name]["wall_s"] = time.perf_counter() - t0
        if world > 1:
            torch.cuda.synchronize()
            torch.distributed.barrier()
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=2000)
    ap.add_argument("--warmup", type=int, default=20)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--configs", default=os.environ.get("DMCMC_BENCH_CONFIGS", "all"),
                    help="comma list of c1,c2,c3,c4,c5 or 'all': the headline line is always C3; the others are nested under 'configs'")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)

    if args.impl == "reference":
        return run_reference(args)

    import torch
    import torch.distributed as dist

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the imputation path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    pkg, spec = make_spec(N_CHAINS)
    from diffusionmcmc_jl_b200.engine import Engine
    from diffusionmcmc_jl_b200.backend import ImputationRunner

    # one context per GPU; Philox streams are keyed by the GLOBAL chain index
    eng = Engine(spec, device=local_rank, seed=1, chain_offset=rank * N_CHAINS)
    eng.init_paths()
    runner = ImputationRunner(eng, half_len=HALF_LEN)
    steps_per_sweep = eng.C * eng.S

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---------------- device-resident throughput (value) ----------------
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    runner.run_device(0, args.warmup)
    # keep the GPU under the bench's own load while nvidia-smi starts sampling
    runner.run_device(args.warmup, max(args.warmup, min(args.steps, 2000)))
    eng.kernel_ms()
    eng.launch_count()
    barrier()
    eng.timer_start()                      # CUDA events on the stream the kernels are launched on
    runner.run_device(100_000, args.steps)
    region_ms = eng.timer_stop()           # waits for this rank's K sweeps only: no collective inside the timed region
    kern = runner.kernel_breakdown()       # sampled single-launch durations (every 8th launch carries events)
    n_launches = eng.launch_count()
    # ---------------- end to end through the host API (e2e) ----------------
    runner.reserve_host(args.steps)  # pinned host buffers are allocated and touched outside the timed region
    runner.run_host(10_000, args.warmup)
    barrier()
    t1 = time.perf_counter()
    h2d, d2h = runner.run_host(20_000, args.steps)   # returns when the last history row is in host memory
    elapsed_e2e = time.perf_counter() - t1
    clocks = sampler.stop() if rank == 0 else None
    barrier()

    if world > 1:
        tt = torch.tensor([region_ms, elapsed_e2e], device="cuda", dtype=torch.float64)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        region_ms, elapsed_e2e = float(tt[0]), float(tt[1])
    elapsed = region_ms * 1e-3
    eng.close()

    which = ["c1", "c2", "c4", "c5"] if args.configs == "all" else [c for c in args.configs.split(",") if c in ("c1", "c2", "c4", "c5")]
    others = other_configs(pkg, which, args, rank, local_rank, world) if which else {}

    if rank == 0:
        total_steps = steps_per_sweep * world * args.steps
        value = total_steps / elapsed
        peak, which_peak = measured_peak()
        # average duration of the dominant kernel over the timed region: the region holds nothing but its
        # K back-to-back launches, bracketed by CUDA events on their stream
        ms_imp = region_ms / max(1, n_launches)
        achieved = BYTES_PER_STEP * steps_per_sweep / (ms_imp * 1e-3) / 1e9
        traffic, traffic_src = profiled_traffic()
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": elapsed / args.steps * 1e3,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
            "data": "synthetic", "config": workload_config(),
            "timing": "value: CUDA events on the launching stream around the K sweeps, per rank, MAX over ranks; "
                      "e2e: wall clock around the host call, per rank, MAX over ranks; barrier + synchronize before each region",
            "mcmc_chain_iters_per_sec": eng.C * world * args.steps / elapsed / 2.0,
            "sweeps_per_sec": args.steps / elapsed,
            "e2e": {"value": total_steps / elapsed_e2e, "unit": UNIT,
                    "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "ms_per_step": elapsed_e2e / args.steps * 1e3,
                    "host_buffers": "pinned (dmcmc_host_alloc): history rows land by DMA, one crossing of host memory"},
            "gpu_launches": n_launches,
            "kernels": dict(kern, region_ms_per_launch=ms_imp,
                            note="impute_ms: events around single launches (every 8th; the event pair adds launch "
                                 "latency); region_ms_per_launch: events around all K launches / K"),
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s",
                         "frac": achieved / peak, "traffic": traffic, "traffic_source": traffic_src, "peak_source": which_peak,
                         "kernel": "solve_kernel<ModelFHN, PSI, MODE_IMPUTE, PHILOX, REINV, FAST>",
                         "algorithmic_bytes_per_launch": BYTES_PER_STEP * steps_per_sweep,
                         "mode": "fused layout switch: reads the accepted X (16 B/step), writes X deg (16 B/step), W never touched"},
            "clocks": clocks,
        }
        if not args.no_cpu_baseline and world == 1:
            cores = os.cpu_count() or 1
            _, small = make_spec(CPU_SAMPLE_CHAINS)
            rate, secs = cpu_oracle_rate(small, CPU_SAMPLE_SWEEPS, cores, pkg)
            line["cpu_baseline"] = {
                "value": rate, "unit": UNIT, "cores": cores, "kind": "port",
                "sample": f"{CPU_SAMPLE_CHAINS} of {N_CHAINS} chains x {N_OBS} obs x 100 steps, "
                          f"{CPU_SAMPLE_SWEEPS} sweeps (relayout + impute), {secs:.1f} s wall, "
                          f"oracle/dmcmc_oracle.c with OpenMP over (chain, block) units"}
        if others:
            line["configs"] = others
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
