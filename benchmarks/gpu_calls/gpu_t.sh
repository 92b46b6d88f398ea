#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_cuda_cat.py -x -q -m gpu 2>&1 | tail -2
LB_SWEEP_TIMING=1 timeout 900 python bench.py --rows 200000 --steps 2 --warmup 3 --no-cpu-baseline --no-aux > gpurun_out/t_bench_small.json 2> gpurun_out/t_bench_small.err
grep "LB_SWEEP_TIMING" gpurun_out/t_bench_small.err | tail -5
python -c "
import json; d=json.load(open('gpurun_out/t_bench_small.json')); print(d['value'], d['ms_per_step'], d['roofline_aux']['phases_ms_per_step'], d['roofline']['launches'])"
LB_SWEEP_TIMING=1 timeout 1200 python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-aux > gpurun_out/t_bench_full.json 2> gpurun_out/t_bench_full.err
grep "LB_SWEEP_TIMING" gpurun_out/t_bench_full.err | tail -8
python -c "
import json; d=json.load(open('gpurun_out/t_bench_full.json')); print(d['value'], d['ms_per_step'], d['roofline_aux']['phases_ms_per_step'], d['roofline']['launches'])"
