#!/bin/bash
# round 2, GPU call A: parity of the new sweep kernel, then single-chain / 10-chain rates, both kernels
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/a_smi.txt 2>&1
timeout 900 python -m pytest tests/test_cuda_cat.py -x -q -m gpu > gpurun_out/a_test_cat.log 2>&1
echo "test_cuda_cat rc=$?" >> gpurun_out/a_test_cat.log
tail -5 gpurun_out/a_test_cat.log
for spec in "C2 1" "C2 10" "C2 100"; do
  for k in spec classic; do
    LB_SWEEP_KERNEL=$k timeout 300 python benchmarks/sweep_bench.py $spec >> gpurun_out/a_sweep.jsonl 2>> gpurun_out/a_sweep.err
  done
done
LB_SWEEP_KERNEL=spec timeout 600 python benchmarks/sweep_bench.py C3 1 200000 32 >> gpurun_out/a_sweep.jsonl 2>> gpurun_out/a_sweep.err
LB_SWEEP_KERNEL=spec timeout 600 python benchmarks/sweep_bench.py C3 10 200000 0 >> gpurun_out/a_sweep.jsonl 2>> gpurun_out/a_sweep.err
cat gpurun_out/a_sweep.jsonl
tail -5 gpurun_out/a_sweep.err
timeout 900 python -m pytest tests/test_cuda_struct.py -x -q -m gpu > gpurun_out/a_test_struct.log 2>&1
echo "test_cuda_struct rc=$?" >> gpurun_out/a_test_struct.log
tail -5 gpurun_out/a_test_struct.log
