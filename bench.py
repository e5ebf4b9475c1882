#!/usr/bin/env python
"""bench.py -- bridge path-steps/sec of the imputation hot path (BASELINE.json metric).

Workload (N=1 and per GPU for N>1, weak scaling): BASELINE.json configs[2] -- FitzHugh-Nagumo,
1024 independent chains x 100 observations, dt = 1e-3 (100 Euler steps per inter-observation
interval => 1.024e7 path-steps per sweep per GPU), pCN rho = 0.96, chequerboard blocking with
block_half_len = 1 (each sweep = layout switch [W re-inversion + ll recompute] + imputation of
every (chain, block)), synthetic observations (NumPy PCG64 seed 1), Philox noise in-register.

A "step" is one PathImputation update (one sweep over every chain and block).

  value  : path-steps/s, state resident in HBM, K sweeps issued back to back on the device.
  e2e    : the same sweeps through the reference-facing call (`run!`-equivalent of the host
           mirror -> C ABI) with HOST buffers: per sweep the pCN rho vector goes host->device and
           the ll deg / ll / accepted history rows come device->host (src/run!_alterations.jl:57-75).
  --impl reference : the CPU restatement of the reference path (oracle/, OpenMP over units) on
           the box's host cores on a bounded sample of the same workload.
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
BYTES_PER_STEP = 32  # FHN: read W (8m) + write W deg (8m) + write X deg (8d), m=1 d=2
METRIC, UNIT = "bridge path-steps/sec", "path-steps/s"
CPU_SAMPLE_CHAINS, CPU_SAMPLE_SWEEPS = 512, 24  # ~1.2e8 path-steps: 10-30 s of CPU work over the host cores


def profiled_traffic():
    """dram__bytes_read.sum + dram__bytes_write.sum of the dominant kernel, per launch, from the
    committed ncu --set full capture of this workload (profiles/r1_traffic.json); None if absent."""
    try:
        with open(os.path.join(ROOT, "profiles", "r1_traffic.json")) as f:
            return int(json.load(f)["traffic_bytes_per_launch"])
    except Exception:
        return None


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


def cpu_oracle_rate(spec_small, sweeps, nthreads, pkg):
    """path-steps/s of the CPU restatement on a bounded sample (same algorithm, same layouts)."""
    from oracle import oracle
    o = oracle.Oracle(spec_small, nthreads=nthreads)
    o.init_paths(seed=1)
    lays = [pkg.problems.chequerboard_layout(o.rec_nobs, HALF_LEN, p) for p in (0, 1)]
    o.relayout(*lays[0])
    o.impute(seed=1, it=0)  # warm
    t0 = time.perf_counter()
    for it in range(sweeps):
        o.relayout(*lays[(it + 1) % 2])
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
                        "half_len=1 (relayout + imputation per sweep), Philox noise",
            "chains_per_gpu": N_CHAINS, "n_obs": N_OBS, "steps_per_interval": 100,
            "block_half_len": HALF_LEN, "parallelism": "chains sharded, no collective",
            "l2": "state 0.5 GB per GPU >> 126 MB L2 (inputs larger than L2)"}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=2000)
    ap.add_argument("--warmup", type=int, default=20)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
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
        os.environ["NCCL_DEBUG"] = "WARN"  # rank 0 prints exactly one line: no NCCL version banner on stdout
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
    t0 = time.perf_counter()
    eng.timer_start()                      # CUDA events on the stream the kernels are launched on
    runner.run_device(100_000, args.steps)
    region_ms = eng.timer_stop()
    barrier()
    elapsed = time.perf_counter() - t0
    kern = runner.kernel_breakdown()       # sampled single-launch durations (every 8th launch carries events)
    n_launches = eng.launch_count()
    # ---------------- end to end through the host API (e2e) ----------------
    runner.reserve_host(args.steps)  # host buffers are allocated and touched outside the timed region
    runner.run_host(10_000, args.warmup)
    barrier()
    t1 = time.perf_counter()
    h2d, d2h = runner.run_host(20_000, args.steps)
    barrier()
    elapsed_e2e = time.perf_counter() - t1
    clocks = sampler.stop() if rank == 0 else None

    if world > 1:
        tt = torch.tensor([elapsed, elapsed_e2e], device="cuda", dtype=torch.float64)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        elapsed, elapsed_e2e = float(tt[0]), float(tt[1])

    if rank == 0:
        total_steps = steps_per_sweep * world * args.steps
        value = total_steps / elapsed
        peak, which = measured_peak()
        # average duration of the dominant kernel over the timed region: the region holds nothing but its
        # K back-to-back launches, bracketed by CUDA events on their stream
        ms_imp = region_ms / max(1, n_launches)
        achieved = BYTES_PER_STEP * steps_per_sweep / (ms_imp * 1e-3) / 1e9
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": elapsed / args.steps * 1e3,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
            "data": "synthetic", "config": workload_config(),
            "mcmc_chain_iters_per_sec": eng.C * world * args.steps / elapsed / 2.0,
            "sweeps_per_sec": args.steps / elapsed,
            "e2e": {"value": total_steps / elapsed_e2e, "unit": UNIT,
                    "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "ms_per_step": elapsed_e2e / args.steps * 1e3},
            "gpu_launches": n_launches,
            "kernels": dict(kern, region_ms_per_launch=ms_imp,
                            note="impute_ms: events around single launches (every 8th; the event pair adds launch "
                                 "latency); region_ms_per_launch: events around all K launches / K"),
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s",
                         "frac": achieved / peak, "traffic": profiled_traffic(), "peak_source": which,
                         "kernel": "solve_kernel<ModelFHN, PSI, MODE_IMPUTE, PHILOX, REINV, FAST>",
                         "algorithmic_bytes_per_launch": BYTES_PER_STEP * steps_per_sweep},
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
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
