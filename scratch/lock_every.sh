set -e
run() {  # args: extra-flags chains
  touch loom_b200/csrc/lb_cat.cu
  make -C loom_b200/csrc EXTRA="$1" > /dev/null 2>&1
  echo "== $1: $(grep -A3 k_cat_sweepILi1 loom_b200/csrc/lb_cat.o.ptxas.log | grep -E 'spill' | head -1)"
  LB_SWEEP_WARPS=1 timeout 300 python scratch/multi_bench.py $2 2>&1 | grep chains
}
run "-DLB_LOCKSTEP_EVERY=4" 672
run "-DLB_LOCKSTEP_EVERY=16" 672
run "-DLB_W1_CTAS=24" 1008
run "-DLB_W1_CTAS=24 -DLB_LOCKSTEP_EVERY=4" 1008
touch loom_b200/csrc/lb_cat.cu; make -C loom_b200/csrc > /dev/null 2>&1
