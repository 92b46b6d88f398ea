#!/bin/bash
# round 2, final GPU call: the whole -m gpu suite, smoke(), the default bench run, the sweep's DRAM traffic at bench size
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/final_tests.log 2>&1
echo "pytest rc=$?"; tail -n 3 gpurun_out/final_tests.log
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
SECONDS=0
timeout 1500 python bench.py > gpurun_out/final_bench_n1.json 2> gpurun_out/final_bench_n1.err
echo "bench rc=$? wall ${SECONDS}s"; tail -n 2 gpurun_out/final_bench_n1.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/final_bench_n1.json'))
print('value %.4g e2e %.4g ms/step %.0f frac %.3g traffic %s launches %s'%(d['value'],d['e2e']['value'],d['ms_per_step'],d['roofline']['frac'],d['roofline']['traffic'],d['gpu_launches']))
print(d['roofline_aux']['phases_ms_per_step'])
print({k:(v if not isinstance(v,dict) else v.get('cells_per_s')) for k,v in d['rates'].items()})
print(d['cpu_baseline']['value'], d['cpu_baseline']['build'])
PY
timeout 1500 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:k_cat_sweep --csv --log-file gpurun_out/final_sweep_dram.csv python bench.py --steps 1 --warmup 3 --no-aux --no-cpu-baseline > gpurun_out/final_sweep_dram.log 2>&1
echo "traffic rc=$?"
