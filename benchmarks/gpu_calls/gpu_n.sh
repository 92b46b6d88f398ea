#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_dist.py -x -q -m gpu 2>&1 | tail -2
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29523 bench.py --gpus 2 --rows 100000 --steps 1 --warmup 3 --no-aux > gpurun_out/n_b2.json 2> gpurun_out/n_b2.err
echo "bench N=2 rc=$?"; grep -v "^\*\|OMP_NUM" gpurun_out/n_b2.err | tail -n 3
python -c "
import json; d=json.loads([l for l in open('gpurun_out/n_b2.json') if l.startswith('{\"metric\"')][-1]); print(d['n_gpus'], '%.4g'%d['value'], d['kind_kernel']['sharded']['allreduce_busbw_GBs'])"
