#!/bin/bash
# round 2, GPU call X: the whole -m gpu suite (new K3f test included), smoke(), a C4-shaped and a C2 bench line
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/x_tests.log 2>&1
echo "pytest rc=$?"; tail -n 4 gpurun_out/x_tests.log
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 900 python bench.py --workload C4 --rows 100000 --steps 2 --warmup 3 --no-aux --no-cpu-baseline > gpurun_out/x_bench_c4.json 2> gpurun_out/x_bench_c4.err
echo "C4-shaped rc=$?"; tail -n 3 gpurun_out/x_bench_c4.err
python -c "
import json; d=json.load(open('gpurun_out/x_bench_c4.json')); print('C4 x100k:', d['value'], d['ms_per_step'], d['roofline_aux']['phases_ms_per_step'], d['config']['kinds'])"
timeout 900 python bench.py --workload C2 --steps 3 --warmup 3 --no-aux --no-cpu-baseline > gpurun_out/x_bench_c2.json 2> gpurun_out/x_bench_c2.err
echo "C2 rc=$?"; tail -n 3 gpurun_out/x_bench_c2.err
python -c "
import json; d=json.load(open('gpurun_out/x_bench_c2.json')); print('C2:', d['value'], d['ms_per_step'], d['roofline_aux']['phases_ms_per_step'])"
