#!/bin/bash
# round 2, GPU call B: phase-cycle profile of the sweep kernels (profiling build)
mkdir -p gpurun_out
export LOOM_B200_LIB=$PWD/loom_b200/csrc/libloomb200_prof.so
for k in spec classic; do
  echo "== $k" >> gpurun_out/b_prof.log
  LB_PROFILE=1 LB_SWEEP_KERNEL=$k timeout 300 python scratch/prof_sweep.py 30000 >> gpurun_out/b_prof.log 2>&1
done
cat gpurun_out/b_prof.log
