#!/bin/bash
# A/B harness for the Lorenz-96 kernel: each library variant in ab/ x warps per CTA (DMCMC_L96_WARPS)
cp diffusionmcmc.jl_b200/csrc/libdmcmc_b200.so /tmp/lib_keep.so
for f in ab/lib_*.so; do
  cp $f diffusionmcmc.jl_b200/csrc/libdmcmc_b200.so
  for nc in ${NCS:-8}; do
    echo "== $f warps=$nc"
    DMCMC_L96_WARPS=$nc python tools/ncu_target_l96.py ${1:-1024} ${2:-256} 2>&1 | tail -n 1
  done
done
cp /tmp/lib_keep.so diffusionmcmc.jl_b200/csrc/libdmcmc_b200.so
