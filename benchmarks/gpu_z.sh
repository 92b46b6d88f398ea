#!/bin/bash
mkdir -p gpurun_out
run() { LB_SWEEP_TIMING=1 timeout 900 python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-aux > gpurun_out/z_b.json 2> gpurun_out/z_b.err
  grep "round 0" gpurun_out/z_b.err | tail -1 | cut -c1-160
  python -c "
import json; d=json.load(open('gpurun_out/z_b.json')); print('$1', '%.4g'%d['value'], '%.0f'%d['ms_per_step'], d['roofline_aux']['phases_ms_per_step']['cat_kernel'])"; }
run "prefetch value rows"
LOOM_B200_LIB=$PWD/loom_b200/csrc/libloomb200_pf2.so run "prefetch value + count rows"
