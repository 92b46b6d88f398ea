import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from loom_b200 import Engine, _abi, synth, sweep_actions
cuda = _abi.load_cuda()
R, F = int(sys.argv[1]), int(sys.argv[2])
model, rows, _ = synth.generate(R, [("dd256", F)], density=0.5, seed=3, single_kind=True)
e = Engine(cuda); e.set_model(model); e.set_rows(rows)
e.cat_sweep(sweep_actions(20000), seed=1)
G = e.kind_info(0).n_groups
for _ in range(3): e.score_rows(0, fetch=False)
e.sync(); e.timer_start()
reps = 10
for _ in range(reps): e.score_rows(0, fetch=False)
ms = e.timer_stop() / reps
ncell = rows.n_cells()
in_bytes = ncell + R * F / 8.0
out_bytes = 4.0 * G * R
print("R %d F %d G %d: %.3f ms  %.2f Gcells/s  in %.1f MB out %.1f MB  %.0f GB/s (%.1f%% of 6533.5)" % (
    R, F, G, ms, ncell / ms / 1e6, in_bytes / 1e6, out_bytes / 1e6, (in_bytes + out_bytes) / ms / 1e6,
    100 * (in_bytes + out_bytes) / ms / 1e6 / 6533.5))
