#!/bin/bash
mkdir -p gpurun_out
SECONDS=0
timeout 1500 python bench.py > gpurun_out/p_bench.json 2> gpurun_out/p_bench.err
echo "full rc=$? wall ${SECONDS}s"; grep -E "rank 0" gpurun_out/p_bench.err; tail -n 3 gpurun_out/p_bench.err
