#!/bin/bash
mkdir -p gpurun_out
export LB_SWEEP_KERNEL=spec
python benchmarks/ncu_sweep.py C3 8000 > gpurun_out/e_plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:k_cat_sweep_spec -s 3 -c 1 -f -o gpurun_out/e_spec16 python benchmarks/ncu_sweep.py C3 8000 > gpurun_out/e_ncu.log 2>&1
tail -n 3 gpurun_out/e_plain.log; tail -n 5 gpurun_out/e_ncu.log
LB_SWEEP_KERNEL=classic timeout 300 python benchmarks/sweep_bench.py C3 1 200000 0 > gpurun_out/e_classic.json 2>gpurun_out/e_classic.err
cat gpurun_out/e_classic.json
