#!/bin/bash
# round 2, GPU call J: K1t v2 (cell-loader warps): parity + A/B; K3s phase profile
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_cuda_cat.py -x -q -m gpu -k "tensor_core or score_rows" > gpurun_out/j_tc_test.log 2>&1
echo "tc tests rc=$?"; tail -n 6 gpurun_out/j_tc_test.log
for spec in "bb 64 1000000" "bb 256 1000000" "dd16 128 1000000" "dd16 128 1000000 2" "dd256 16 200000" "dpd16 64 500000"; do
  timeout 300 python benchmarks/k1_bench.py $spec >> gpurun_out/j_k1.jsonl 2>> gpurun_out/j_k1.err
done
python - <<'PY'
import json
for l in open('gpurun_out/j_k1.jsonl'):
    d=json.loads(l)
    print(d['type'],d['features'],d['rows'],'G',d['groups'],'planes',d['planes'],'gather %.3f ms %.3g cells/s frac %.3f | tc %.3f ms %.3g cells/s frac %.3f | x%.2f diff %.2e'%(d['gather']['ms'],d['gather']['cells_per_s'],d['gather']['frac_of_hbm'],d['tcgen05']['ms'],d['tcgen05']['cells_per_s'],d['tcgen05']['frac_of_hbm'],d['speedup_tc_over_gather'],d.get('max_abs_diff_vs_gather',-1)))
PY
tail -n 5 gpurun_out/j_k1.err
( export LOOM_B200_LIB=$PWD/loom_b200/csrc/libloomb200_prof.so
  LB_PROFILE=1 LB_SWEEP_KERNEL=spec timeout 300 python benchmarks/prof_sweep.py 30000 > gpurun_out/j_prof.log 2>&1 )
grep -A8 "kind 6" gpurun_out/j_prof.log
