#!/bin/bash
# round 2, GPU call Q: K3f (featureless-kind kernel) parity + rates; hybrid sparse score_diff
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_cuda_cat.py tests/test_cuda_struct.py tests/test_tare.py tests/test_z_loom_infer.py -x -q -m gpu > gpurun_out/q_tests.log 2>&1
echo "tests rc=$?"; tail -n 8 gpurun_out/q_tests.log
timeout 300 python benchmarks/k1d_bench.py 200000 | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('K1d dense %.2f ms sparse %.2f ms x%.2f diff %.1e'%(d['dense_score_value']['ms_all_kinds'],d['sparse_score_diff']['ms_all_kinds'],d['speedup_sparse_over_dense'],d['max_rel_diff_dense_vs_sparse']))"
timeout 900 python bench.py --rows 200000 --steps 2 --warmup 3 --no-cpu-baseline --no-aux > gpurun_out/q_bench_small.json 2> gpurun_out/q_bench_small.err
echo "small rc=$?"; tail -n 3 gpurun_out/q_bench_small.err
python -c "
import json; d=json.load(open('gpurun_out/q_bench_small.json')); print(d['value'], d['ms_per_step'], d['roofline_aux']['phases_ms_per_step'])"
LB_SWEEP_FLAT=0 timeout 900 python bench.py --rows 200000 --steps 2 --warmup 3 --no-cpu-baseline --no-aux 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('flat off:', d['value'], d['ms_per_step'])"
