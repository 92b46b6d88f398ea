import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np
from loom_b200 import Engine, _abi, synth, sweep_actions
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
cuda = _abi.load_cuda()
oracle = _abi.Abi(os.path.join(REPO, "oracle", "libloom_oracle.so"), "lo_")
R = int(sys.argv[1]); single = sys.argv[2] == "1"; mode = sys.argv[3]
types = [("bb", 5), ("dd4", 3), ("dd16", 3), ("dd256", 2), ("dpd7", 2), ("gp", 4), ("nich", 4)]
model, rows, _ = synth.generate(R, types, density=0.6, seed=11, single_kind=single)
o, g = Engine(oracle), Engine(cuda)
for e in (o, g):
    e.set_model(model); e.set_rows(rows)
acts = sweep_actions(R)
if mode == "ra":
    acts = np.concatenate([acts, sweep_actions(R, add=True, remove=True)])
S = 64
pg, sg = g.cat_sweep(acts, seed=42, debug=True, debug_stride=S)
po, so = o.cat_sweep(acts, seed=42, debug=True, debug_stride=S)
np.set_printoptions(precision=4, linewidth=220, suppress=True)
for k in range(model.n_kinds):
    bad = None
    for i in range(len(acts)):
        if pg[i, k] != po[i, k] or not np.allclose(sg[i, k], so[i, k], rtol=1e-4, atol=1e-4, equal_nan=True):
            bad = i; break
    print("kind", k, "nf", len(model.kind_features(k)), "first mismatch", bad)
    if bad is not None:
        i = bad
        print("  action", acts[i] >> 1, "remove" if acts[i] & 1 else "add", "picks", pg[i, k], po[i, k])
        n = int(np.sum(~np.isnan(so[i, k])))
        print("  gpu", sg[i, k][:n + 2]); print("  ora", so[i, k][:n + 2])
