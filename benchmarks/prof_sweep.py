import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, bench
from loom_b200 import Engine, _abi
cuda = _abi.load_cuda()
R = int(sys.argv[1])
model, rows = bench.load_workload("C2", R)
first, drain = bench.step_actions(R)
e = Engine(cuda); e.set_model(model); e.set_rows(rows)
for it in range(3):
    e.cat_sweep(first, seed=it); e.cat_sweep(drain, seed=it)
print("actions per step", len(first) + len(drain), "ADD", 2 * R, "REMOVE", 2 * R, "steps 3")
e.close()
