"""Query path measurement on the C2 mix: lb_query_score (batch Score of every row: K1 per kind +
row log-sum-exp) and lb_query_sample (joint draws of half the features given the other half), device vs the oracle
on host cores.  usage: query_bench.py ROWS DRAWS [oracle]   -> one JSON line"""
import json, os, sys, time
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
import numpy as np
from loom_b200 import Engine, _abi, synth
from loom_b200.engine import RowBlock
R, DRAWS = int(sys.argv[1]), int(sys.argv[2])
use_oracle = len(sys.argv) > 3
model, rows, truth = synth.generate(R, [("bb", 25), ("dd256", 25), ("gp", 25), ("nich", 25)], density=0.5, seed=2)
abi = _abi.Abi(os.path.join(REPO, "oracle", "libloom_oracle.so"), "lo_") if use_oracle else _abi.load_cuda()
e = Engine(abi); e.set_model(model); e.set_rows(rows)
for k in range(model.n_kinds):            # the generating partition as the latent state
    uniq, inv = np.unique(truth["assign"][k][:R].astype(np.uint32), return_inverse=True)
    e.set_groups(k, len(uniq)); e.set_assignments(k, inv.astype(np.uint32))
e.accumulate_rows()
to_sample = np.arange(1, model.n_features, 2, dtype=np.uint32)
e.query_score(0, min(R, 4096)); e.query_sample(0, to_sample, 256, seed=1)      # warm-up
t_score = 1e9
for _ in range(5):                                                              # best of 5: the call is ~ms
    t0 = time.perf_counter(); s = e.query_score(); t_score = min(t_score, time.perf_counter() - t0)
t0 = time.perf_counter(); v = e.query_sample(0, to_sample, DRAWS, seed=2); t_sample = time.perf_counter() - t0
assert np.isfinite(s).all()
print(json.dumps({"impl": "oracle" if use_oracle else "cuda", "fused": os.environ.get("LB_QUERY_FUSED") == "1", "cores": os.cpu_count(), "rows": R, "kinds": model.n_kinds,
                  "groups": [int(e.kind_info(k).n_groups) for k in range(model.n_kinds)],
                  "score_rows_per_s": R / t_score, "score_cells_per_s": rows.n_cells() / t_score, "score_ms": t_score * 1e3,
                  "draws": DRAWS, "features_per_draw": len(to_sample), "sample_values_per_s": DRAWS * len(to_sample) / t_sample,
                  "sample_ms": t_sample * 1e3, "timing": "host wall clock around the C ABI call, results copied to host"}))
