LB_SWEEP_WARPS=8 python scratch/multi_bench.py 42
LB_SWEEP_WARPS=4 python scratch/multi_bench.py 84
LB_SWEEP_WARPS=2 python scratch/multi_bench.py 169
LB_SWEEP_WARPS=1 python scratch/multi_bench.py 338
LB_SWEEP_WARPS=1 python scratch/multi_bench.py 169
LB_SWEEP_WARPS=2 python scratch/multi_bench.py 84
