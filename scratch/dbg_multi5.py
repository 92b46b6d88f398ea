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
for it in range(3):
    seed = 5000 + it
    for acts in (first, drain):
        picks, _ = g.cat_sweep(acts, seed=seed, debug=True)
        bad, err = t.replay(oracle, o, acts, seed, picks, None, 0)
        assert bad == 0
seed = 5003
off = 0
fine_from = int(sys.argv[1]); fine = int(sys.argv[2])
while off < len(first):
    ch = 2000 if off + 2000 <= fine_from else fine
    chunk = first[off:off + ch]
    print("chunk", off, ch, "G", [g.kind_info(k).n_groups for k in range(K)], flush=True)
    picks, scores = g.cat_sweep(chunk, seed=seed, action_offset=off, debug=True, debug_stride=64)
    for i in range(len(chunk)):
        if chunk[i] & 1: continue
        for k in range(K):
            nvalid = int(np.sum(~np.isnan(scores[i, k])))
            if picks[i, k] >= nvalid or not np.isfinite(scores[i, k][:nvalid]).all():
                print("ODD action", off + i, "row", chunk[i] >> 1, "kind", k, "pick", picks[i, k], "nvalid", nvalid)
                print("   scores", scores[i, k][:nvalid + 2])
                print("   prev picks", picks[max(0, i - 3):i + 1, k])
                sys.exit(0)
    bad, err = t.replay(oracle, o, chunk, seed, picks, scores, 64, offset=off)
    if bad or err > 1e-4:
        print("  bad", bad, "err", err); break
    off += ch
