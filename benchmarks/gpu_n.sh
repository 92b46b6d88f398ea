#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_dist.py -x -q -m gpu > gpurun_out/n_tests.log 2>&1
echo "tests rc=$?"; tail -n 6 gpurun_out/n_tests.log
grep "rank" gpurun_out/native_nccl_test.log | head -4
