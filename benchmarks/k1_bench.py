"""K1 batch score_value A/B: the gather kernel (k_score_rows) against the tcgen05 contraction (k_score_rows_tc) on
one kind holding FEATURES columns of one type.  usage: k1_bench.py TYPE FEATURES ROWS [PLANES] -> one JSON line
TYPE: bb | dd16 | dd256 | dpd16 ...; the chain state is the generating partition (G ~ 30-60 groups)."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np

from loom_b200 import Engine, _abi, synth

t, F, R = sys.argv[1], int(sys.argv[2]), int(sys.argv[3])
planes = int(sys.argv[4]) if len(sys.argv) > 4 else 3
cuda = _abi.load_cuda()
model, rows, truth = synth.generate(R, [(t, F)], density=0.5, seed=7, single_kind=True)
e = Engine(cuda)
e.set_model(model)
e.set_rows(rows)
uniq, inv = np.unique(truth["assign"][0], return_inverse=True)
e.set_groups(0, len(uniq))
e.set_assignments(0, inv.astype(np.uint32))
e.accumulate_rows()
G = e.kind_info(0).n_groups
peak = json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json"))).get("hbm_gbs", 6650.0)
n_obs = rows.n_cells()
out = {"type": t, "features": F, "rows": R, "groups": G, "planes": planes, "observed_cells": n_obs}
os.environ["LB_K1_TC_PLANES"] = str(planes)
ref = None
for name, flag in (("gather", "0"), ("tcgen05", "1")):
    os.environ["LB_K1_TC"] = flag
    s = e.score_rows(0, 0, min(R, 20000))
    if ref is None:
        ref = s
    else:
        out["max_abs_diff_vs_gather"] = float(np.abs(s.astype(np.float64) - ref).max())
        out["max_abs_score"] = float(np.abs(ref).max())
    for _ in range(3):
        e.score_rows(0, fetch=False)
    e.sync()
    reps = 10
    e.timer_start()
    for _ in range(reps):
        e.score_rows(0, fetch=False)
    ms = e.timer_stop() / reps
    width = 1 if t.startswith(("bb", "dd")) else 4
    bytes_alg = n_obs * width + R * F / 8.0 + 4.0 * G * R          # cells + mask in, 4 G out per row
    out[name] = {"ms": ms, "cells_per_s": n_obs / (ms * 1e-3), "rows_per_s": R / (ms * 1e-3),
                 "achieved_GBs": bytes_alg / (ms * 1e-3) / 1e9, "frac_of_hbm": bytes_alg / (ms * 1e-3) / 1e9 / peak}
out["speedup_tc_over_gather"] = out["gather"]["ms"] / out["tcgen05"]["ms"]
print(json.dumps(out))
