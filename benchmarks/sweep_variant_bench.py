"""A/B of a build variant of the library on the headline kernel: N chains over ONE shared C2 row block, one timed
step (ADD all + REMOVE/ADD all + REMOVE all) after one warm-up step, device time from CUDA events.  The library is the
product's unless LOOM_B200_LIB points at a variant (loom_b200/csrc/Makefile `variants`).
With a third argument `streams` every chain walks the rows in its own order (lb_cat_sweep_streams: what loom's
per-sample shuffles amount to) instead of the shared order.
usage: sweep_variant_bench.py CHAINS ROWS [streams]   -> one JSON line"""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, bench
from loom_b200 import Engine, _abi, cat_sweep_multi, cat_sweep_streams
cuda = _abi.load_cuda()
n, R = int(sys.argv[1]), int(sys.argv[2])
streams = len(sys.argv) > 3 and sys.argv[3] == "streams"
model, rows = bench.load_workload("C2", R)
first, drain = bench.step_actions(R)
cells = 4 * rows.n_cells()
engines = []
for c in range(n):
    e = Engine(cuda); e.set_model(model)
    if c == 0: e.set_rows(rows)
    else: e.share_rows(engines[0])
    engines.append(e)
if streams:      # the same ADD / REMOVE pattern, row positions permuted per chain
    rng = np.random.default_rng(0)
    perms = np.stack([rng.permutation(R).astype(np.uint32) for _ in range(n)])
    first_s = np.stack([(p[first >> 1] << 1) | (first & 1) for p in perms])
    drain_s = np.stack([(p[drain >> 1] << 1) | (drain & 1) for p in perms])
def step(it):
    seeds = np.arange(n, dtype=np.uint64) * 1000 + it
    if streams:
        cat_sweep_streams(engines, first_s, seeds); cat_sweep_streams(engines, drain_s, seeds)
    else:
        cat_sweep_multi(engines, first, seeds); cat_sweep_multi(engines, drain, seeds)
step(0)
for e in engines: e.sync()
engines[0].timer_start()
step(1)
ms = engines[0].timer_stop()
# the phases of a step on their own (ADD-only and REMOVE+ADD-pair rates reported separately)
phases = {}
if not streams:
    seeds = np.arange(n, dtype=np.uint64) * 1000 + 2
    for name, acts, per_row in (("add_only", first[:R], 1), ("remove_add_pairs", first[R:], 2), ("remove_only", drain, 1)):
        for e in engines: e.sync()
        engines[0].timer_start()
        cat_sweep_multi(engines, acts, seeds)
        t = engines[0].timer_stop()
        phases[name] = {"ms": t, "cells_per_s": n * per_row * rows.n_cells() / (t * 1e-3)}
# the state after the step is the empty one for every chain; a cheap invariant that the variant kept the books
empty = all(int(e.get_groups(0)[0].sum()) == 0 for e in engines[:4])
print(json.dumps({"lib": os.path.basename(cuda.path), "row_order": "per chain" if streams else "shared", "chains": n, "rows": R, "ms_per_step": ms,
                  "cells_per_s": n * cells / (ms * 1e-3), "phases": phases, "drained_to_empty": empty}))
