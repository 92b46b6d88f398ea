"""Tare / sparsify measurement on the C2 mix with the boolean columns made sparse (P(true) = 0.02, the
text-presence style columns of config C4): k_differ_hist (add_rows) and k_differ_rows + scan (compress_rows), device
time from CUDA events on the engine's stream, against the numpy-free oracle on host cores.
usage: differ_bench.py ROWS [oracle]   -> one JSON line"""
import ctypes as C, json, os, sys, time
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
import numpy as np
from loom_b200 import Engine, _abi, synth
R = int(sys.argv[1])
use_oracle = len(sys.argv) > 2
model, rows, _ = synth.generate(R, [("bb", 25), ("dd256", 25), ("gp", 25), ("nich", 25)], density=0.5, seed=2)
rng = np.random.default_rng(0)
rows.cat8[:25] = rng.random((25, R)) < 0.02
rows.mask[:25] = 0xFFFFFFFF                                      # sparse columns are always observed (C4)
abi = _abi.Abi(os.path.join(REPO, "oracle", "libloom_oracle.so"), "lo_") if use_oracle else _abi.load_cuda()
e = Engine(abi); e.set_model(model); e.set_rows(rows)
obs = rows.observed()
width = np.array([1 if f["type"] in ("bb", "dd") else 4 for f in model.features])
counted = np.array([f["type"] != "nich" for f in model.features])
mask_bytes = R / 8.0
tare_bytes = float((obs[counted] * width[counted, None]).sum()) + mask_bytes * counted.sum()
row_bytes = float((obs * width[:, None]).sum()) + mask_bytes * model.n_features
to, tv = e.make_tare()                                            # warm-up + the tare used below
d = e.sparsify(to, tv)
def timed(fn, reps=5):
    e.sync(); e.timer_start()
    for _ in range(reps): fn()
    return e.timer_stop() / reps
null_i32 = C.POINTER(C.c_int32)()
def add_rows():
    abi.check(abi.differ_add_rows(e.h, null_i32, None))
abi.check(abi.differ_reset(e.h))
ms_hist = timed(add_rows)
n_pos, n_neg = C.c_uint64(), C.c_uint64()
to_p, tv_p = to.ctypes.data_as(_abi._u8p), tv.ctypes.data_as(_abi._u32p)
def compress():
    abi.check(abi.differ_compress_rows(e.h, to_p, tv_p, C.byref(n_pos), C.byref(n_neg)))
ms_rows = timed(compress)
out_bytes = 16.0 * (R + 1) + 8.0 * n_pos.value + 4.0 * n_neg.value
print(json.dumps({"impl": "oracle" if use_oracle else "cuda", "cores": os.cpu_count(), "rows": R, "features": model.n_features,
                  "tared_columns": int(to.sum()), "cells": int(obs.sum()), "pos_entries": n_pos.value, "neg_entries": n_neg.value,
                  "tare": {"ms": ms_hist, "rows_per_s": R / (ms_hist * 1e-3), "algorithmic_bytes": tare_bytes,
                           "achieved_gb_s": tare_bytes / (ms_hist * 1e-3) / 1e9},
                  "sparsify": {"ms": ms_rows, "rows_per_s": R / (ms_rows * 1e-3), "algorithmic_bytes": 2 * row_bytes + out_bytes,
                               "achieved_gb_s": (2 * row_bytes + out_bytes) / (ms_rows * 1e-3) / 1e9,
                               "note": "count pass + scan + fill pass: the block is read twice; includes one 16-byte read-back "
                                       "of the totals between the passes"},
                  "timing": "engine timer (CUDA events on the launch stream; monotonic clock for the oracle), mean of 5"}))
