#!/bin/bash
# round 2, GPU call G: whole -m gpu suite, then the new C3 bench (small, then full)
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/g_tests.log 2>&1
echo "pytest rc=$?" >> gpurun_out/g_tests.log
tail -n 6 gpurun_out/g_tests.log
timeout 600 python bench.py --rows 100000 --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/g_bench_small.json 2> gpurun_out/g_bench_small.err
echo "small rc=$?"; tail -n 5 gpurun_out/g_bench_small.err; head -c 3000 gpurun_out/g_bench_small.json
timeout 900 python bench.py > gpurun_out/g_bench.json 2> gpurun_out/g_bench.err
echo "full rc=$?"; tail -n 5 gpurun_out/g_bench.err
