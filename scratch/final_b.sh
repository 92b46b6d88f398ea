set -x
timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 2 --warmup 3 > gpurun_out/r1_bench_n2.json 2> gpurun_out/r1_bench_n2.err
tail -2 gpurun_out/r1_bench_n2.err; head -c 400 gpurun_out/r1_bench_n2.json
