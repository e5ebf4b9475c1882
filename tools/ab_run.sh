#!/bin/bash
# A/B harness: run the C3 probe with each library variant found in ab/ x CTA sizes (the in-tree library is restored afterwards)
cp diffusionmcmc.jl_b200/csrc/libdmcmc_b200.so /tmp/lib_keep.so
for f in ab/lib_*.so; do
  cp $f diffusionmcmc.jl_b200/csrc/libdmcmc_b200.so
  for th in ${THREADS:-0}; do
    echo "== $f threads=$th"
    python tools/perf_probe2.py ${1:-400} $th 2>&1 | tail -1
  done
done
cp /tmp/lib_keep.so diffusionmcmc.jl_b200/csrc/libdmcmc_b200.so
