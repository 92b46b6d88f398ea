#!/bin/bash
mkdir -p gpurun_out
export LB_SWEEP_KERNEL=spec
python benchmarks/ncu_sweep.py C2 6000 > gpurun_out/d_plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:k_cat_sweep_spec -s 2 -c 1 -f -o gpurun_out/d_spec8 python benchmarks/ncu_sweep.py C2 6000 > gpurun_out/d_ncu.log 2>&1
tail -n 3 gpurun_out/d_plain.log; tail -n 5 gpurun_out/d_ncu.log
ls -la gpurun_out/d_spec8.ncu-rep
