#!/bin/bash
# round 2, GPU call C: K3s v2 (compiled rows): parity, phase profile, rates
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_cuda_cat.py -x -q -m gpu > gpurun_out/c_test_cat.log 2>&1
echo "test_cuda_cat rc=$?" >> gpurun_out/c_test_cat.log
tail -5 gpurun_out/c_test_cat.log
( export LOOM_B200_LIB=$PWD/loom_b200/csrc/libloomb200_prof.so
  LB_PROFILE=1 LB_SWEEP_KERNEL=spec timeout 300 python scratch/prof_sweep.py 30000 > gpurun_out/c_prof.log 2>&1 )
grep -A8 "kind 6\|kind 3" gpurun_out/c_prof.log
for spec in "C2 1" "C2 10" "C3 1 200000 0" "C3 10 200000 0"; do
  LB_SWEEP_KERNEL=spec timeout 300 python benchmarks/sweep_bench.py $spec >> gpurun_out/c_sweep.jsonl 2>> gpurun_out/c_sweep.err
done
LB_SPEC_WIDE16=0 LB_SWEEP_KERNEL=spec timeout 300 python benchmarks/sweep_bench.py C3 1 200000 0 >> gpurun_out/c_sweep.jsonl 2>> gpurun_out/c_sweep.err
LB_SWEEP_KERNEL=spec timeout 300 python benchmarks/sweep_bench.py C3 1 200000 32 >> gpurun_out/c_sweep.jsonl 2>> gpurun_out/c_sweep.err
python - <<'PY'
import json
for l in open('gpurun_out/c_sweep.jsonl'):
    d=json.loads(l)
    print(d['workload'],d['chains'],d['rows'],d['ephemeral'],'ms/step %.0f'%d['ms_per_step'],'cells/s %.3g'%d['cells_per_s'],{k:round(v['us_per_action'],2) for k,v in d['phases'].items()})
PY
tail -3 gpurun_out/c_sweep.err
