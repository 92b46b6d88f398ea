"""Aggregate cat-kernel throughput with many chains in one launch (lb_cat_sweep_multi).
usage: python scratch/multi_bench.py <chains> [rows]   (LB_SWEEP_WARPS picks the CTA width)"""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, bench
from loom_b200 import Engine, _abi, cat_sweep_multi
cuda = _abi.load_cuda()
n = int(sys.argv[1]); R = int(sys.argv[2]) if len(sys.argv) > 2 else 100000
model, rows = bench.load_workload("C2", R)
first, drain = bench.step_actions(R)
cells = 4 * rows.n_cells()
engines = []
for c in range(n):
    e = Engine(cuda); e.set_model(model); e.set_rows(rows); engines.append(e)
def step(it):
    seeds = np.arange(n, dtype=np.uint64) * 1000 + it
    cat_sweep_multi(engines, first, seeds); cat_sweep_multi(engines, drain, seeds)
step(0); step(1)
for e in engines: e.sync()
t0 = time.perf_counter(); engines[0].timer_start()
step(2)
ms = engines[0].timer_stop(); wall = time.perf_counter() - t0
print("chains %d W=%s: %.1f ms/step (wall %.1f) -> %.1f M cells/s aggregate, %.2f M per chain" % (
    n, os.environ.get("LB_SWEEP_WARPS", "auto"), ms, wall * 1e3, n * cells / ms / 1e3, cells / ms / 1e3))
