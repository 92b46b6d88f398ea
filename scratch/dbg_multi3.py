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
hist = int(sys.argv[1]) if len(sys.argv) > 1 else 3
for it in range(hist):
    g.cat_sweep(first, seed=5000 + it); g.cat_sweep(drain, seed=5000 + it)
seed = 5003
CH = 2000
K = model.n_kinds
for off in range(0, 64000, CH):
    chunk = first[off:off + CH]
    picks, scores = g.cat_sweep(chunk, seed=seed, action_offset=off, debug=True, debug_stride=64)
    bad, err = t.replay(oracle, o, chunk, seed, picks, scores, 64, offset=off)
    finite = np.isfinite(scores[~np.isnan(scores)]).all()
    mx = [int(picks[:, k][picks[:, k] != 0xFFFFFFFF].max()) for k in range(K)]
    print(off, "bad", bad, "err %.2e" % err, "maxpick", mx, "G", [g.kind_info(k).n_groups for k in range(K)], flush=True)
    if bad or err > 1e-4:
        break
