import sys, os, time, threading
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")
import numpy as np
import bench
from loom_b200 import Engine, _abi
cuda = _abi.load_cuda()
R = int(sys.argv[1]); C = int(sys.argv[2])
model, rows = bench.load_workload("C2", R)
first, drain = bench.step_actions(R)
engines = []
for c in range(C):
    e = Engine(cuda); e.set_model(model); e.set_rows(rows); engines.append(e)
def step(c, it, base):
    e = engines[c]; seed = base + 1000 * c + it
    t0 = time.perf_counter(); e.cat_sweep(first, seed=seed); t1 = time.perf_counter()
    g = [e.kind_info(k).n_groups for k in range(model.n_kinds)]
    e.cat_sweep(drain, seed=seed); t2 = time.perf_counter()
    return (t1 - t0) * 1e3, (t2 - t1) * 1e3, g
# sequential, different seed bases
for base in ():
    for it in range(5):
        a, b, g = step(0, it, base)
        print("seq base", base, "it", it, "first %.0f drain %.0f ms" % (a, b), "groups", g, flush=True)
# threaded
for base in (0, 500000):
    for it in range(3, 6):
        res = [None] * C
        def run(c): res[c] = step(c, it, base)
        ths = [threading.Thread(target=run, args=(c,)) for c in range(C)]
        t0 = time.perf_counter(); [t.start() for t in ths]; [t.join() for t in ths]
        print("thr base", base, "it", it, "wall %.0f" % ((time.perf_counter() - t0) * 1e3), [(round(r[0]), max(r[2])) for r in res], flush=True)
