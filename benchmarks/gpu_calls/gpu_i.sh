#!/bin/bash
# round 2, GPU call I: tcgen05 K1 parity + A/B; K3s with the vectorised gather and the register patch: parity + rates
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_cuda_cat.py -x -q -m gpu > gpurun_out/i_cat_test.log 2>&1
echo "cat tests rc=$?"; tail -n 15 gpurun_out/i_cat_test.log
timeout 600 python -m pytest tests/test_zz_query_gpu.py tests/test_cuda_struct.py -x -q -m gpu > gpurun_out/i_tests.log 2>&1
echo "tests rc=$?"; tail -n 4 gpurun_out/i_tests.log
export LB_SWEEP_KERNEL=spec
for spec in "C3 1 200000 0" "C3 1 1000000 32" "C3 10 200000 32" "C2 1" "C2 10"; do
  timeout 300 python benchmarks/sweep_bench.py $spec >> gpurun_out/i_sweep.jsonl 2>> gpurun_out/i_sweep.err
done
python - <<'PY'
import json
for l in open('gpurun_out/i_sweep.jsonl'):
    d=json.loads(l)
    print(d['workload'],d['chains'],d['rows'],d['ephemeral'],'ms/step %.0f'%d['ms_per_step'],'cells/s %.3g'%d['cells_per_s'],{k:round(v['us_per_action'],2) for k,v in d['phases'].items()})
PY
tail -n 3 gpurun_out/i_sweep.err
unset LB_SWEEP_KERNEL
for spec in "bb 64 1000000" "bb 256 1000000" "dd16 128 1000000" "dd16 128 1000000 2" "dd256 16 200000" "dpd16 64 500000"; do
  timeout 300 python benchmarks/k1_bench.py $spec >> gpurun_out/i_k1.jsonl 2>> gpurun_out/i_k1.err
done
python - <<'PY'
import json
for l in open('gpurun_out/i_k1.jsonl'):
    d=json.loads(l)
    print(d['type'],d['features'],d['rows'],'G',d['groups'],'planes',d['planes'],'gather %.3f ms %.3g cells/s frac %.3f | tc %.3f ms %.3g cells/s frac %.3f | x%.2f diff %.2e'%(d['gather']['ms'],d['gather']['cells_per_s'],d['gather']['frac_of_hbm'],d['tcgen05']['ms'],d['tcgen05']['cells_per_s'],d['tcgen05']['frac_of_hbm'],d['speedup_tc_over_gather'],d.get('max_abs_diff_vs_gather',-1)))
PY
tail -n 5 gpurun_out/i_k1.err
timeout 600 python bench.py --rows 200000 --steps 2 --warmup 3 --no-cpu-baseline --no-aux > gpurun_out/i_bench_small.json 2> gpurun_out/i_bench_small.err
echo "small rc=$?"; tail -n 3 gpurun_out/i_bench_small.err
python -c "
import json; d=json.load(open('gpurun_out/i_bench_small.json')); print(d['value'], d['ms_per_step'], d['roofline_aux']['phases_ms_per_step'])"
