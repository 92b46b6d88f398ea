#!/bin/bash
timeout 600 python -m pytest tests/test_cuda_cat.py tests/test_tare.py tests/test_zz_query_gpu.py -x -q -m gpu -k "score or query or tare or diff" 2>&1 | tail -2
python - <<'PY'
import os, sys
sys.path.insert(0, '.')
import numpy as np, bench
from loom_b200 import Engine, _abi
cuda=_abi.load_cuda()
model, rows, truth = bench.load_workload("C3", 400000, with_truth=True)
e=Engine(cuda); e.set_model(model); e.set_rows(rows); bench.install_truth(e, truth)
big=int(np.argmax([len(model.kind_features(k)) for k in range(model.n_kinds)]))
for _ in range(2): e.score_rows(big, fetch=False)
e.sync(); e.timer_start()
for _ in range(5): e.score_rows(big, fetch=False)
ms=e.timer_stop()/5
G=e.kind_info(big).n_groups; kf=model.kind_features(big)
cells=float(rows.cells_per_feature()[kf].sum())
byt=cells+rows.n_rows*len(kf)/8+4.0*G*rows.n_rows
print("K1 gather, C3 big kind, 400k rows, G=%d: %.3f ms, %.3g cells/s, %.0f GB/s algorithmic"%(G,ms,cells/(ms*1e-3),byt/(ms*1e-3)/1e9))
PY
LB_SWEEP_KERNEL=spec timeout 300 python benchmarks/sweep_bench.py C3 30 200000 32 | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('30 chains x 200k rows (spec kernels):', {k:(round(v['us_per_action'],2), '%.3g'%v['cells_per_s']) for k,v in d['phases'].items()})"
