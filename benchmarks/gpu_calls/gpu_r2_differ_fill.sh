#!/bin/bash
# round 2: the coalesced fill pass of compress_rows (k_differ_fill) -- bit-exact tests, then A/B against the thread-per-row walk
mkdir -p gpurun_out
timeout 100 python -m pytest tests/test_tare.py tests/test_reference_pins.py -x -q -m gpu > gpurun_out/fill_tests.log 2>&1; echo "pytest rc=$?"; tail -n 2 gpurun_out/fill_tests.log
timeout 40 python benchmarks/differ_bench.py 1000000 > gpurun_out/r2_differ_bench_fill.json 2>/dev/null; echo "new rc=$?"
LB_DIFFER_FILL=rows timeout 40 python benchmarks/differ_bench.py 1000000 > gpurun_out/r2_differ_bench_rows.json 2>/dev/null; echo "old rc=$?"
python -c "
import json
for f in ('fill','rows'):
    d=json.load(open('gpurun_out/r2_differ_bench_%s.json'%f)); print(f, d['sparsify']['ms'], d['sparsify']['achieved_gb_s'], d['pos_entries'])
"
