#!/bin/bash
# round 2, GPU call U: the default bench run (N = 1) and the reference arm, as the driver runs them
mkdir -p gpurun_out
SECONDS=0
timeout 900 python bench.py --impl reference --gpus 1 --steps 3 --warmup 3 > gpurun_out/u_bench_ref.json 2> gpurun_out/u_bench_ref.err
echo "reference arm rc=$? wall ${SECONDS}s"; tail -n 2 gpurun_out/u_bench_ref.err
SECONDS=0
timeout 1500 python bench.py --gpus 1 --steps 3 --warmup 3 > gpurun_out/u_bench_n1.json 2> gpurun_out/u_bench_n1.err
echo "ours rc=$? wall ${SECONDS}s"; tail -n 3 gpurun_out/u_bench_n1.err
python - <<'PY'
import json
r=json.load(open('gpurun_out/u_bench_ref.json')); d=json.load(open('gpurun_out/u_bench_n1.json'))
print('reference %.4g cells/s (%s)'%(r['value'], r['cpu_baseline']['build']))
print('ours value %.4g e2e %.4g ms/step %.0f frac %.3g traffic %s'%(d['value'],d['e2e']['value'],d['ms_per_step'],d['roofline']['frac'],d['roofline']['traffic']))
print(d['roofline_aux']['phases_ms_per_step'])
print({k:(v if not isinstance(v,dict) else v.get('cells_per_s')) for k,v in d['rates'].items()})
PY
