timeout 600 python -m pytest tests/test_cuda_cat.py tests/test_golden.py -m gpu -x -q 2>&1 | tail -1
LB_SWEEP_WARPS=1 python scratch/multi_bench.py 672 2>&1 | grep chains
python scratch/multi_bench.py 1 2>&1 | grep chains
