"""Cat-kernel sweep on one workload: N chains over ONE shared row block, device time from CUDA events.
A step = ADD all + REMOVE/ADD all + REMOVE all (4 actions per row), as in bench.py; the phases are also timed alone
(ADD-only, REMOVE+ADD pairs and REMOVE-only rates separately).
usage: sweep_bench.py WORKLOAD CHAINS [ROWS [EPHEMERAL_KINDS]]   -> one JSON line
env:   LB_SWEEP_KERNEL=spec|classic, LB_SWEEP_WARPS=1|2|4|8 (classic width)"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np

import bench
from loom_b200 import Engine, _abi, cat_sweep_multi

cuda = _abi.load_cuda()
name, n = sys.argv[1], int(sys.argv[2])
R = int(sys.argv[3]) if len(sys.argv) > 3 else 0
E = int(sys.argv[4]) if len(sys.argv) > 4 else 0
model, rows = bench.load_workload(name, R)
R = rows.n_rows
first, drain = bench.step_actions(R)
cells = 4 * rows.n_cells()
engines = []
for c in range(n):
    e = Engine(cuda)
    e.set_model(model)
    if c == 0:
        e.set_rows(rows)
    else:
        e.share_rows(engines[0])
    if E:
        e.init_featureless_kinds(E, seed=100 + c)
    engines.append(e)


def step(it):
    seeds = np.arange(n, dtype=np.uint64) * 1000 + it
    cat_sweep_multi(engines, first, seeds)
    cat_sweep_multi(engines, drain, seeds)


step(0)
for e in engines:
    e.sync()
engines[0].launch_count(reset=True)
engines[0].timer_start()
step(1)
ms = engines[0].timer_stop()
launches = engines[0].launch_count()
phases = {}
seeds = np.arange(n, dtype=np.uint64) * 1000 + 2
for pname, acts, per_row in (("add_only", first[:R], 1), ("remove_add_pairs", first[R:], 2), ("remove_only", drain, 1)):
    for e in engines:
        e.sync()
    engines[0].timer_start()
    cat_sweep_multi(engines, acts, seeds)
    t = engines[0].timer_stop()
    phases[pname] = {"ms": t, "cells_per_s": n * per_row * rows.n_cells() / (t * 1e-3),
                     "us_per_action": 1e3 * t / len(acts)}
groups = [int(engines[0].kind_info(k).n_groups) for k in range(engines[0].n_kinds())]
empty = all(int(e.get_groups(0)[0].sum()) == 0 for e in engines[:4])
print(json.dumps({"workload": name, "kernel": os.environ.get("LB_SWEEP_KERNEL", "auto"),
                  "classic_warps": os.environ.get("LB_SWEEP_WARPS"), "chains": n, "rows": R,
                  "kinds": engines[0].n_kinds(), "ephemeral": E, "ms_per_step": ms,
                  "cells_per_s": n * cells / (ms * 1e-3), "sweep_launches": int(launches[1]), "phases": phases,
                  "drained_to_empty": empty}))
