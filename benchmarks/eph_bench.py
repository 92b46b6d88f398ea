"""init_featureless_kinds timing: direct draw vs a draw prepared ahead of time (lb_prepare_featureless_kinds).
usage: eph_bench.py ROWS [CHAINS]"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import bench
from loom_b200 import Engine, _abi, hyperprior, sweep_actions
import concurrent.futures

cuda = _abi.load_cuda()
R = int(sys.argv[1]); n = int(sys.argv[2]) if len(sys.argv) > 2 else 1
model, rows, truth = bench.load_workload("C3", R, with_truth=True)
engs = []
for c in range(n):
    e = Engine(cuda); e.set_model(model)
    if c == 0: e.set_rows(rows)
    else: e.share_rows(engs[0])
    e.set_hyper_prior(hyperprior.defaults())
    bench.install_truth(e, truth)
    engs.append(e)
def run(prep):
    if prep:
        for c, e in enumerate(engs): e.prepare_featureless_kinds(32, seed=5 + c)
        time.sleep(3.0)
    t0 = time.perf_counter()
    with concurrent.futures.ThreadPoolExecutor(max_workers=n) as pool:
        list(pool.map(lambda c: engs[c].init_featureless_kinds(32, seed=5 + c), range(n)))
    return time.perf_counter() - t0
for prep in (False, True, False, True):
    print("prepared" if prep else "direct  ", "%.3f s for %d chain(s) x 32 kinds x %d rows" % (run(prep), n, R), [engs[0].kind_info(k).n_groups for k in range(9, 14)], flush=True)
