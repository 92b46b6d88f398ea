#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_cuda_cat.py tests/test_cuda_struct.py -x -q -m gpu 2>&1 | tail -2
LB_SWEEP_TIMING=1 timeout 1200 python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-aux > gpurun_out/y_bench_full.json 2> gpurun_out/y_bench_full.err
grep "LB_SWEEP_TIMING" gpurun_out/y_bench_full.err | tail -4
python -c "
import json; d=json.load(open('gpurun_out/y_bench_full.json')); print(d['value'], d['ms_per_step'], d['roofline_aux']['phases_ms_per_step'])"
timeout 600 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:k_cat_sweep -s 4 -c 4 --csv --log-file gpurun_out/y_sweep_dram.csv python bench.py --steps 1 --warmup 3 --no-aux --no-cpu-baseline > gpurun_out/y_sweep_dram.log 2>&1
grep "k_cat_sweep_spec<16" gpurun_out/y_sweep_dram.csv | cut -d, -f5,13,15 | cut -c1-160
