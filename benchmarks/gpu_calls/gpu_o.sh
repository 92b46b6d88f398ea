#!/bin/bash
mkdir -p gpurun_out
timeout 300 python benchmarks/eph_bench.py 200000 10
timeout 600 python -m pytest tests/test_cuda_struct.py tests/test_cuda_cat.py -x -q -m gpu 2>&1 | tail -3
timeout 600 python bench.py --rows 200000 --steps 2 --warmup 3 --no-cpu-baseline --no-aux > gpurun_out/o_bench_small.json 2> gpurun_out/o_bench_small.err
echo "small rc=$?"; tail -n 3 gpurun_out/o_bench_small.err
python -c "
import json; d=json.load(open('gpurun_out/o_bench_small.json')); print(d['value'], d['ms_per_step'], d['roofline_aux']['phases_ms_per_step'])"
for spec in "bb 64 1000000" "bb 256 1000000" "dd16 128 1000000 2"; do
  timeout 300 python benchmarks/k1_bench.py $spec | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print(d['type'],d['features'],'gather %.3f ms | tc %.3f ms frac %.3f x%.2f'%(d['gather']['ms'],d['tcgen05']['ms'],d['tcgen05']['frac_of_hbm'],d['speedup_tc_over_gather']))"
done
