#!/bin/bash
# round 2, GPU call P: sparse score_diff parity + timing, then the full default bench run (N = 1)
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_tare.py -x -q -m gpu > gpurun_out/p_tests.log 2>&1
echo "tare tests rc=$?"; tail -n 12 gpurun_out/p_tests.log
timeout 300 python benchmarks/k1d_bench.py 200000 > gpurun_out/p_k1d.json 2> gpurun_out/p_k1d.err; tail -n 3 gpurun_out/p_k1d.err; cat gpurun_out/p_k1d.json
/usr/bin/time -v timeout 1500 python bench.py > gpurun_out/p_bench.json 2> gpurun_out/p_bench.err
echo "full rc=$?"; grep -E "Elapsed|rank 0" gpurun_out/p_bench.err
