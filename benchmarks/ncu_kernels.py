"""One pass over every batch kernel of the path, for ncu (each kernel launched a handful of times on a warm engine):
K1 gather (k_score_rows), K1t (k_score_rows_tc), K2 (k_accum_feature via accumulate_rows), K4 + K5 (kind proposer),
K7 (k_hyper_features), the differ passes, and a short many-chain sweep.
usage: ncu_kernels.py WORKLOAD ROWS [CHAINS]      (the state is the generating partition, as in bench.py)"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np

import bench
from loom_b200 import Engine, _abi, cat_sweep_multi, hyperprior, sweep_actions, synth

cuda = _abi.load_cuda()
name, R = sys.argv[1], int(sys.argv[2])
chains = int(sys.argv[3]) if len(sys.argv) > 3 else 0
model, rows, truth = bench.load_workload(name, R, with_truth=True)
e = Engine(cuda)
e.set_model(model)
e.set_rows(rows)
bench.install_truth(e, truth)
big = int(np.argmax([len(model.kind_features(k)) for k in range(model.n_kinds)]))
os.environ["LB_K1_TC"] = "0"
e.score_rows(big, fetch=False)                 # K1 gather
e.accumulate_rows(sign=-1)                     # K2
e.accumulate_rows(sign=+1)
e.init_featureless_kinds(32, seed=5)
e.kind_proposer_score(fetch=False)             # K4 + K5
e.set_hyper_prior(hyperprior.defaults())
e.hyper_run(seed=3)                            # K7
e.sync()
# K1t on a kind it is meant for: low-cardinality categorical columns
m2, r2, t2 = synth.generate(min(R, 400000), [("dd16", 128)], density=0.5, seed=7, single_kind=True)
e2 = Engine(cuda)
e2.set_model(m2)
e2.set_rows(r2)
bench.install_truth(e2, np.stack(t2["assign"]).astype(np.uint32))
os.environ["LB_K1_TC"] = "0"
e2.score_rows(0, fetch=False)
os.environ["LB_K1_TC"] = "1"
e2.score_rows(0, fetch=False)
e2.sync()
if chains:
    engs = []
    for c in range(chains):
        g = Engine(cuda)
        g.set_model(model)
        g.share_rows(e)
        g.init_featureless_kinds(32, seed=100 + c)
        engs.append(g)
    n = min(R, 20000)
    cat_sweep_multi(engs, sweep_actions(n), np.arange(chains, dtype=np.uint64))
    cat_sweep_multi(engs, sweep_actions(n, add=True, remove=True), np.arange(chains, dtype=np.uint64))
    for g in engs:
        g.sync()
print("ok")
