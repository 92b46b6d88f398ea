"""Does a width class of the sweep run slower beside the others?  The two widest kinds of C3 (128 + 64 columns) alone,
then the same two kinds with the small kinds and 32 ephemeral kinds beside them.  usage: iso_bench.py ROWS CHAINS"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from loom_b200 import Engine, _abi, synth, sweep_actions, cat_sweep_multi, hyperprior

cuda = _abi.load_cuda()
R, n = int(sys.argv[1]), int(sys.argv[2])
def run(types, f2k, eph, label):
    model, rows, _ = synth.generate(R, types, density=0.5, seed=3, f2k=f2k)
    engs = []
    for c in range(n):
        e = Engine(cuda); e.set_model(model)
        if c == 0: e.set_rows(rows)
        else: e.share_rows(engs[0])
        e.set_hyper_prior(hyperprior.defaults())
        if eph: e.init_featureless_kinds(eph, seed=100 + c)
        engs.append(e)
    seeds = np.arange(n, dtype=np.uint64) * 1000
    cat_sweep_multi(engs, sweep_actions(R), seeds)
    pairs = sweep_actions(R, add=True, remove=True)
    cat_sweep_multi(engs, pairs, seeds + 1)
    for e in engs: e.sync()
    engs[0].timer_start()
    cat_sweep_multi(engs, pairs, seeds + 2)
    ms = engs[0].timer_stop()
    print("%-40s %8.1f ms  %.2f us per row action" % (label, ms, 1e3 * ms / len(pairs)), flush=True)
    for e in engs: e.close()
run([("dd256", 192)], [0] * 128 + [1] * 64, 0, "128 + 64 columns alone")
run([("dd256", 192)], [0] * 128 + [1] * 64, 32, "... + 32 ephemeral kinds")
run([("dd256", 256)], [0] * 128 + [1] * 64 + [2] * 32 + [3] * 16 + [4] * 8 + [5] * 4 + [6] * 2 + [7, 8], 0, "all 9 kinds of C3")
run([("dd256", 256)], [0] * 128 + [1] * 64 + [2] * 32 + [3] * 16 + [4] * 8 + [5] * 4 + [6] * 2 + [7, 8], 32, "all 9 kinds + 32 ephemeral")
