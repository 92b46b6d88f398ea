#!/bin/bash
# round 2, last GPU call: the whole -m gpu suite (with the reference-pin differ test on the device) and smoke()
mkdir -p gpurun_out
timeout 400 python -m pytest tests -x -q -m gpu > gpurun_out/last_tests.log 2>&1
echo "pytest rc=$?"; tail -n 3 gpurun_out/last_tests.log
timeout 120 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -1
