#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_cuda_cat.py -x -q -m gpu > gpurun_out/f_test_cat.log 2>&1
echo "test_cuda_cat rc=$?" >> gpurun_out/f_test_cat.log
tail -n 4 gpurun_out/f_test_cat.log
export LB_SWEEP_KERNEL=spec
for spec in "C3 1 200000 0" "C3 10 200000 0" "C2 1"; do
  timeout 300 python benchmarks/sweep_bench.py $spec >> gpurun_out/f_sweep.jsonl 2>> gpurun_out/f_sweep.err
done
LOOM_B200_LIB=$PWD/loom_b200/csrc/libloomb200_w12.so timeout 300 python benchmarks/sweep_bench.py C3 1 200000 0 >> gpurun_out/f_sweep.jsonl 2>> gpurun_out/f_sweep.err
LB_SPEC_WIDE16=0 timeout 300 python benchmarks/sweep_bench.py C3 1 200000 0 >> gpurun_out/f_sweep.jsonl 2>> gpurun_out/f_sweep.err
python - <<'PY'
import json
for l in open('gpurun_out/f_sweep.jsonl'):
    d=json.loads(l)
    print(d['workload'],d['chains'],d['rows'],d['ephemeral'],'ms/step %.0f'%d['ms_per_step'],'cells/s %.3g'%d['cells_per_s'],{k:round(v['us_per_action'],2) for k,v in d['phases'].items()})
PY
tail -n 3 gpurun_out/f_sweep.err
