import sys, os
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO); sys.path.insert(0, os.path.join(REPO, "tests"))
import numpy as np
import bench
from loom_b200 import Engine, _abi
import test_cuda_cat as t
cuda = _abi.load_cuda()
oracle = _abi.Abi(os.path.join(REPO, "oracle", "libloom_oracle.so"), "lo_")
R = 30000
model, rows = bench.load_workload("C2", R)
first, drain = bench.step_actions(R)
g = Engine(cuda); o = Engine(oracle)
for e in (g, o):
    e.set_model(model); e.set_rows(rows)
K = model.n_kinds
for it in range(4):
    seed = 5000 + it
    for name, acts in (("first", first), ("drain", drain)):
        picks, _ = g.cat_sweep(acts, seed=seed, debug=True)
        bad, err = t.replay(oracle, o, acts, seed, picks, None, 0)
        gi = [(g.kind_info(k).n_groups, g.kind_info(k).sample_size) for k in range(K)]
        oi = [(o.kind_info(k).n_groups, o.kind_info(k).sample_size) for k in range(K)]
        print(it, name, "bad", bad, "gpu", gi, "oracle", oi if oi != gi else "same", flush=True)
        if bad: sys.exit(0)
