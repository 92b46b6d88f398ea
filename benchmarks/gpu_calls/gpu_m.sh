#!/bin/bash
# round 2, GPU call M: the whole -m gpu suite, the K1 A/B with ncu of the tcgen05 kernel, small bench
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/m_tests.log 2>&1
echo "pytest rc=$?"; tail -n 6 gpurun_out/m_tests.log
rm -f gpurun_out/m_k1.jsonl
for spec in "bb 64 1000000" "bb 256 1000000" "dd16 128 1000000" "dd16 128 1000000 2" "dd256 16 200000" "dpd16 64 500000"; do
  timeout 300 python benchmarks/k1_bench.py $spec >> gpurun_out/m_k1.jsonl 2>> gpurun_out/m_k1.err
done
python - <<'PY'
import json
for l in open('gpurun_out/m_k1.jsonl'):
    d=json.loads(l)
    print(d['type'],d['features'],d['rows'],'G',d['groups'],'planes',d['planes'],'gather %.3f ms %.3g cells/s frac %.3f | tc %.3f ms %.3g cells/s frac %.3f | x%.2f diff %.2e'%(d['gather']['ms'],d['gather']['cells_per_s'],d['gather']['frac_of_hbm'],d['tcgen05']['ms'],d['tcgen05']['cells_per_s'],d['tcgen05']['frac_of_hbm'],d['speedup_tc_over_gather'],d.get('max_abs_diff_vs_gather',-1)))
PY
tail -n 5 gpurun_out/m_k1.err
timeout 600 python bench.py --rows 200000 --steps 2 --warmup 3 --no-cpu-baseline --no-aux > gpurun_out/m_bench_small.json 2> gpurun_out/m_bench_small.err
echo "small rc=$?"; tail -n 3 gpurun_out/m_bench_small.err
python -c "
import json; d=json.load(open('gpurun_out/m_bench_small.json')); print(d['value'], d['ms_per_step'], d['roofline_aux']['phases_ms_per_step'])"
ncu --set full --clock-control none --import-source on -k regex:k_score_rows_tc -s 2 -c 1 -f -o gpurun_out/m_k1tc_ncu python benchmarks/k1_bench.py bb 64 1000000 > gpurun_out/m_k1_ncu.log 2>&1
tail -n 2 gpurun_out/m_k1_ncu.log
