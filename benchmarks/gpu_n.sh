#!/bin/bash
# round 2, GPU call N (2 GPUs): native NCCL sharding test, prepared seating test, bench at N=2 on a reduced block
mkdir -p gpurun_out
nvidia-smi -L
timeout 900 python -m pytest tests/test_dist.py tests/test_cuda_struct.py -x -q -m gpu > gpurun_out/n_tests.log 2>&1
echo "tests rc=$?"; tail -n 8 gpurun_out/n_tests.log
cat gpurun_out/native_nccl_test.log 2>/dev/null | grep "rank" | head -4
NCCL_DEBUG=INFO timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --rows 200000 --steps 2 --warmup 3 --no-aux > gpurun_out/n_bench2_small.json 2> gpurun_out/n_bench2_small.err
echo "bench N=2 small rc=$?"; grep -E "NCCL INFO (Channel|Connected|comm|NVLS|Using network|ncclCommInitRank)" gpurun_out/n_bench2_small.err | head -8; grep -v "NCCL INFO" gpurun_out/n_bench2_small.err | tail -n 6
python - <<'PY'
import json
d=json.loads(open('gpurun_out/n_bench2_small.json').read().strip().splitlines()[-1])
print('value %.4g ms/step %.0f n_gpus %d'%(d['value'],d['ms_per_step'],d['n_gpus']))
print(json.dumps(d['kind_kernel'].get('sharded'),indent=1)[:1500])
print(json.dumps({k:v for k,v in d['roofline_aux'].items() if 'shard' in k},indent=1)[:1200])
PY
