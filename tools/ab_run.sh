#!/bin/bash
# A/B harness: run the C3 probe with each library variant found in ab/ (last one restores ab/lib_base.so)
for f in ab/lib_*.so ab/lib_base.so; do
  cp $f diffusionmcmc.jl_b200/csrc/libdmcmc_b200.so
  echo "== $f"
  python tools/perf_probe.py 1024 100 1 64 2>&1 | tail -1
done
