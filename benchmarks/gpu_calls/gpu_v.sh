#!/bin/bash
# round 2, GPU call V (2 GPUs): the default bench at N = 2, launched as the driver launches it
mkdir -p gpurun_out
SECONDS=0
NCCL_DEBUG=INFO NCCL_DEBUG_FILE=gpurun_out/v_nccl_%h_%p.log timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29519 bench.py --gpus 2 --steps 3 --warmup 3 > gpurun_out/v_bench_n2.json 2> gpurun_out/v_bench_n2.err
echo "N=2 rc=$? wall ${SECONDS}s"; grep -v "^\*\|OMP_NUM" gpurun_out/v_bench_n2.err | tail -n 4
cat gpurun_out/v_nccl_*.log | grep -E "Init COMPLETE|via P2P|NVLS|Channel 00/" | head -6 > gpurun_out/v_nccl_summary.txt; rm -f gpurun_out/v_nccl_*.log; cat gpurun_out/v_nccl_summary.txt | cut -c1-200
python - <<'PY'
import json
d=json.loads([l for l in open('gpurun_out/v_bench_n2.json') if l.startswith('{"metric"')][-1])
print('value %.4g e2e %.4g ms/step %.0f n_gpus %d'%(d['value'],d['e2e']['value'],d['ms_per_step'],d['n_gpus']))
sh=d['kind_kernel']['sharded']; print({k:sh[k] for k in ('round_ms','k4_accumulate_ms','k5_score_ms','allreduce_ms','allreduce_bytes','allreduce_busbw_GBs','allgather_ms','value')})
print({k:v for k,v in d['roofline_aux'].items() if 'shard' in k})
PY
