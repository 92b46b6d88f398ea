#!/bin/bash
# round 2, GPU call S: gather unroll A/B; DRAM traffic of the sweep launches at the bench's own size and chain count
mkdir -p gpurun_out
export LB_SWEEP_KERNEL=spec
for v in "" _u5 _u9; do
  LOOM_B200_LIB=$PWD/loom_b200/csrc/libloomb200$v.so timeout 300 python benchmarks/sweep_bench.py C3 1 200000 32 | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('variant [$v] 1 chain', {k:round(v['us_per_action'],2) for k,v in d['phases'].items()})"
  LOOM_B200_LIB=$PWD/loom_b200/csrc/libloomb200$v.so timeout 300 python benchmarks/sweep_bench.py C3 10 200000 32 | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('variant [$v] 10 chains', {k:round(v['us_per_action'],2) for k,v in d['phases'].items()})"
done
unset LB_SWEEP_KERNEL
# one REMOVE+ADD pass of the default bench (10 chains x 1M rows): the four sweep kernels of the second sweep (the first is the
# greedy ADD pass), DRAM bytes only (two metrics: no multi-pass replay of a 10-second kernel beyond what ncu needs)
timeout 1500 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:k_cat_sweep -s 4 -c 4 --csv --log-file gpurun_out/s_sweep_dram.csv python bench.py --steps 1 --warmup 3 --no-aux --no-cpu-baseline > gpurun_out/s_sweep_dram.log 2>&1
echo "dram capture rc=$?"; grep -v "^==" gpurun_out/s_sweep_dram.csv | cut -c1-240 | tail -14
