for w in 8 4 2 1; do
  echo "== LB_SWEEP_WARPS=$w"
  LB_SWEEP_WARPS=$w python -m pytest tests/test_cuda_cat.py tests/test_tare.py -m gpu -x -q 2>&1 | tail -1
  LB_SWEEP_WARPS=$w bash scratch/quick_bench.sh 32
done
