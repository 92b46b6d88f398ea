set -e
for cfg in "1 8" "2 8" "1 16" "2 16"; do
  set -- $cfg
  touch loom_b200/csrc/lb_cat.cu
  make -C loom_b200/csrc EXTRA="-DLB_LOCKSTEP=$1 -DLB_CPB=$2" > /dev/null 2>&1
  echo "LOCKSTEP=$1 CPB=$2: $(grep -A3 k_cat_sweepILi1 loom_b200/csrc/lb_cat.o.ptxas.log | grep -E 'registers' | head -1)"
  LB_SWEEP_WARPS=1 timeout 300 python -m pytest tests/test_cuda_cat.py -m gpu -x -q -k "every_cta_width" 2>&1 | tail -1
  LB_SWEEP_WARPS=1 timeout 300 python scratch/multi_bench.py 672 2>&1 | grep chains
done
touch loom_b200/csrc/lb_cat.cu; make -C loom_b200/csrc > /dev/null 2>&1
