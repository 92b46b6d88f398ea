LB_SWEEP_WARPS=1 python -m pytest tests/test_cuda_cat.py tests/test_tare.py -m gpu -x -q 2>&1 | tail -1
LB_SWEEP_WARPS=1 python scratch/multi_bench.py 336
LB_SWEEP_WARPS=1 python scratch/multi_bench.py 672
