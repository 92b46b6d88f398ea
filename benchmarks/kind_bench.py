"""Kind-kernel measurement on a C3-shaped workload (DD-256 features): K4 + K5 per proposal round,
GPU vs the oracle on host cores.  usage: kind_bench.py ROWS FEATURES [ORACLE_ROWS]"""
import sys, os, time
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
import numpy as np
from loom_b200 import Engine, _abi, synth, sweep_actions
from loom_b200.engine import RowBlock
R, F = int(sys.argv[1]), int(sys.argv[2])
RO = int(sys.argv[3]) if len(sys.argv) > 3 else 0
t0 = time.time()
model, rows, truth = synth.generate(R, [("dd256", F)], density=0.5, seed=3)
print("generated in %.1fs: kinds %d" % (time.time() - t0, model.n_kinds), flush=True)
def setup(abi, rows, truth_assign):
    e = Engine(abi); e.set_model(model); e.set_rows(rows)
    for k in range(model.n_kinds):        # install the generating partition (no Gibbs needed to time K4/K5)
        a = truth_assign[k][:rows.n_rows].astype(np.uint32)
        uniq, inv = np.unique(a, return_inverse=True)
        e.set_groups(k, len(uniq)); e.set_assignments(k, inv.astype(np.uint32))
    e.accumulate_rows()
    e.init_featureless_kinds(32, seed=5)
    return e
cuda = _abi.load_cuda()
g = setup(cuda, rows, truth["assign"])
K = g.n_kinds()
g.kind_proposer_score(fetch=False); g.sync()
reps = 3
t_acc = t_fin = 0.0
for _ in range(reps):
    g.timer_start(); g.proposer_accumulate(); t_acc += g.timer_stop()
    g.timer_start(); g.proposer_finish(fetch=False); t_fin += g.timer_stop()
t_acc /= reps; t_fin /= reps
ncell = rows.n_cells()
G = [g.kind_info(k).n_groups for k in range(K)]
stat_bytes = sum(4 * 257 * ((gg + 31) // 32 * 32) for gg in G) * F          # uint32 [dim+1][pst] per feature per kind
print("GPU: R %d F %d K %d groups %s" % (R, F, K, G[:10]))
print("  K4 accumulate %.2f ms  -> %.3g proposer cells/s" % (t_acc, ncell * K / t_acc * 1e3))
print("  K5 score      %.2f ms  -> %.3g feature-scores/s ; table bytes %.1f MB -> %.0f GB/s (%.1f%% of 6533.5)" % (
    t_fin, F * K / t_fin * 1e3, stat_bytes / 1e6, stat_bytes / t_fin / 1e6, 100 * stat_bytes / t_fin / 1e6 / 6533.5))
ms = t_acc + t_fin
if RO:
    oracle = _abi.Abi(os.path.join(REPO, "oracle", "libloom_oracle.so"), "lo_")
    W = (RO + 31) // 32
    sub = RowBlock(RO, rows.cat8[:, :RO], rows.wide[:, :RO], rows.mask[:, :W])
    o = setup(oracle, sub, truth["assign"])
    t0 = time.perf_counter(); Lo = o.kind_proposer_score(); dt = time.perf_counter() - t0
    print("oracle (%d cores): R %d: %.1f ms per round  feature-scores/s %.3g  proposer cells/s %.3g" % (
        os.cpu_count(), RO, dt * 1e3, F * o.n_kinds() / dt, sub.n_cells() * o.n_kinds() / dt))
