#!/bin/bash
mkdir -p gpurun_out
export LB_SWEEP_KERNEL=spec
rm -f gpurun_out/l_sweep.jsonl
for spec in "C3 1 200000 0" "C3 1 1000000 32" "C3 10 200000 32"; do
  timeout 300 python benchmarks/sweep_bench.py $spec >> gpurun_out/l_sweep.jsonl 2>> gpurun_out/l_sweep.err
done
LB_SPEC_CATONLY=0 timeout 300 python benchmarks/sweep_bench.py C3 1 200000 0 >> gpurun_out/l_sweep.jsonl 2>> gpurun_out/l_sweep.err
python - <<'PY'
import json
for l in open('gpurun_out/l_sweep.jsonl'):
    d=json.loads(l)
    print(d['workload'],d['chains'],d['rows'],d['ephemeral'],'ms/step %.0f'%d['ms_per_step'],'cells/s %.3g'%d['cells_per_s'],{k:round(v['us_per_action'],2) for k,v in d['phases'].items()})
PY
tail -n 3 gpurun_out/l_sweep.err
