import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import bench
from loom_b200 import Engine, _abi
cuda = _abi.load_cuda()
R = 30000
model, rows = bench.load_workload("C2", R)
first, drain = bench.step_actions(R)
e = Engine(cuda)
e.set_model(model); e.set_rows(rows)
# replay the history of chain 5: its 0..2 then it 3 in chunks
for it in range(3):
    e.cat_sweep(first, seed=5000 + it); e.cat_sweep(drain, seed=5000 + it)
print("history done; Gcap via groups:", [e.kind_info(k).n_groups for k in range(model.n_kinds)], flush=True)
seed = 5003
CH = 2000
for off in range(0, len(first), CH):
    chunk = first[off:off + CH]
    try:
        e.cat_sweep(chunk, seed=seed, action_offset=off)
    except Exception as exc:
        print("FAIL at offset", off, str(exc)[:80]); sys.exit(0)
    print(off, [e.kind_info(k).n_groups for k in range(model.n_kinds)], flush=True)
print("first ok")
