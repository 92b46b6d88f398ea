"""K1 on a taxi-like (C4-shaped) block: ProductMixture::score_value of the decoded rows (dense gather, k_score_rows)
against ProductMixture::score_diff from the tare-compressed form (k_score_rows_diff, O(nnz x G)), every kind.
The device tares and sparsifies the block itself (lb_differ_*).  usage: k1d_bench.py ROWS -> one JSON line"""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np

from loom_b200 import Engine, _abi, synth

R = int(sys.argv[1])
cuda = _abi.load_cuda()
model, rows, truth = synth.generate_config("C4", rows=R)
e = Engine(cuda)
e.set_model(model)
e.set_rows(rows)
t0 = time.perf_counter()
to, tv = e.make_tare()
d = e.sparsify(to, tv)
t_differ = time.perf_counter() - t0
e.set_rows_diff(**d)
for k in range(model.n_kinds):
    uniq, inv = np.unique(truth["assign"][k], return_inverse=True)
    e.set_groups(k, len(uniq))
    e.set_assignments(k, inv.astype(np.uint32))
e.accumulate_rows()
os.environ["LB_K1_TC"] = "0"
K = model.n_kinds
worst = 0.0
for k in range(K):
    a, b = e.score_rows(k, 0, min(R, 4000)).astype(np.float64), e.score_rows_diff(k, 0, min(R, 4000)).astype(np.float64)
    worst = max(worst, float((np.abs(a - b) / np.maximum(1.0, np.abs(a))).max()))
out = {"rows": R, "features": model.n_features, "kinds": K, "tared_columns": int(to.sum()), "dense_cells": rows.n_cells(),
       "pos_entries": int(len(d["pos_feature"])), "neg_entries": int(len(d["neg_feature"])),
       "groups": [int(e.kind_info(k).n_groups) for k in range(K)], "max_rel_diff_dense_vs_sparse": worst,
       "tare_and_sparsify_wall_ms": 1e3 * t_differ}
for name, fn in (("dense_score_value", e.score_rows), ("sparse_score_diff", e.score_rows_diff)):
    for k in range(K):
        fn(k, fetch=False)
    e.sync()
    e.timer_start()
    reps = 3
    for _ in range(reps):
        for k in range(K):
            fn(k, fetch=False)
    ms = e.timer_stop() / reps
    out[name] = {"ms_all_kinds": ms, "rows_per_s": R / (ms * 1e-3), "cells_per_s": rows.n_cells() / (ms * 1e-3)}
out["speedup_sparse_over_dense"] = out["dense_score_value"]["ms_all_kinds"] / out["sparse_score_diff"]["ms_all_kinds"]
print(json.dumps(out))
