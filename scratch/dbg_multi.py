import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import bench
from loom_b200 import Engine, _abi
cuda = _abi.load_cuda()
R = int(sys.argv[1]); C = int(sys.argv[2])
model, rows = bench.load_workload("C2", R)
first, drain = bench.step_actions(R)
for c in range(C):
    e = Engine(cuda)
    e.set_model(model); e.set_rows(rows)
    for it in range(4):
        seed = 1000 * c + it
        try:
            e.cat_sweep(first, seed=seed)
            g = [e.kind_info(k).n_groups for k in range(model.n_kinds)]
            e.cat_sweep(drain, seed=seed)
        except Exception as exc:
            print("FAIL chain", c, "it", it, "seed", seed, str(exc)[:100]); sys.exit(0)
        print("ok chain", c, "it", it, "groups", g, flush=True)
