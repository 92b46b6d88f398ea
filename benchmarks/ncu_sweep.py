"""One short single-chain sweep for ncu: C2 rows, ADD pass then REMOVE/ADD pass (2 launches per width class)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, bench
from loom_b200 import Engine, _abi, sweep_actions
cuda = _abi.load_cuda()
name = sys.argv[1]; R = int(sys.argv[2])
model, rows = bench.load_workload(name, R)
e = Engine(cuda); e.set_model(model); e.set_rows(rows)
e.cat_sweep(sweep_actions(R), seed=1)
e.cat_sweep(sweep_actions(R, add=True, remove=True), seed=2)
e.sync()
print("ok", [e.kind_info(k).n_groups for k in range(model.n_kinds)])
