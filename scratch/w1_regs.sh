set -e
for c in 24 32; do
  touch loom_b200/csrc/lb_cat.cu
  make -C loom_b200/csrc EXTRA=-DLB_W1_CTAS=$c > /dev/null 2>&1
  grep -A3 "k_cat_sweepILi1" loom_b200/csrc/lb_cat.o.ptxas.log | grep -E "registers|spill" | tr '\n' ' '; echo
  n=$((2 * c * 148 / 7 / 8 * 8))
  LB_SWEEP_WARPS=1 python scratch/multi_bench.py $n 2>&1 | grep chains
done
touch loom_b200/csrc/lb_cat.cu; make -C loom_b200/csrc > /dev/null 2>&1
