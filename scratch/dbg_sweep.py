import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np
from loom_b200 import Engine, _abi, synth, sweep_actions
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
cuda = _abi.load_cuda()
oracle = _abi.Abi(os.path.join(REPO, "oracle", "libloom_oracle.so"), "lo_")
types = eval(sys.argv[1]) if len(sys.argv) > 1 else [("bb", 5), ("dd4", 3), ("gp", 4), ("nich", 4)]
model, rows, _ = synth.generate(200, types, density=0.6, seed=11, single_kind=True)
o, g = Engine(oracle), Engine(cuda)
for e in (o, g):
    e.set_model(model); e.set_rows(rows)
acts = sweep_actions(40)
pg, sg = g.cat_sweep(acts, seed=42, debug=True, debug_stride=16)
po, so = o.cat_sweep(acts, seed=42, debug=True, debug_stride=16)
np.set_printoptions(precision=4, linewidth=200, suppress=True)
for i in range(len(acts)):
    if pg[i, 0] != po[i, 0] or not np.allclose(sg[i, 0], so[i, 0], rtol=1e-4, atol=1e-4, equal_nan=True):
        print("first mismatch at action", i, "picks", pg[i, 0], po[i, 0])
        print(" gpu", sg[i, 0])
        print(" ora", so[i, 0])
        break
else:
    print("all match", pg[:, 0])
