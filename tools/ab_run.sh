#!/bin/bash
# A/B harness: run the mode probe with each library variant found in ab/
for f in ab/lib_*.so; do
  cp $f diffusionmcmc.jl_b200/csrc/libdmcmc_b200.so
  echo "== $f"
  python tools/probe_modes.py 2>&1 | tail -3
done
