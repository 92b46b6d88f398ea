#!/bin/bash
# round 2, GPU call H: the tail of the -m gpu suite, then the new C3 bench (small, then full)
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_zz_query_gpu.py -x -q -m gpu > gpurun_out/h_tests.log 2>&1
echo "pytest rc=$?" >> gpurun_out/h_tests.log
tail -n 4 gpurun_out/h_tests.log
timeout 600 python bench.py --rows 100000 --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/h_bench_small.json 2> gpurun_out/h_bench_small.err
echo "small rc=$?"; tail -n 5 gpurun_out/h_bench_small.err; head -c 6000 gpurun_out/h_bench_small.json
timeout 1200 python bench.py > gpurun_out/h_bench.json 2> gpurun_out/h_bench.err
echo "full rc=$?"; tail -n 5 gpurun_out/h_bench.err
