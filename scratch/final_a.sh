set -x
python bench.py > gpurun_out/r1_bench_n1.json 2> gpurun_out/r1_bench_n1.err
python bench.py --impl reference > gpurun_out/r1_bench_reference_arm.json 2> gpurun_out/r1_bench_ref.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 40000 --csv --log-file gpurun_out/r1_launches.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/r1_launches_bench.log 2>&1
LB_SWEEP_WARPS=1 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:k_cat_sweep -s 2 -c 2 --csv --log-file gpurun_out/r1_sweep_multi_dram.csv python scratch/multi_bench.py 672 > gpurun_out/r1_sweep_multi_dram.log 2>&1
LB_SWEEP_WARPS=1 ncu --set full --import-source on --clock-control none -k regex:k_cat_sweep -s 2 -c 1 -o gpurun_out/r1_k_cat_sweep_w1 python scratch/multi_bench.py 672 1500 > gpurun_out/r1_k_cat_sweep_w1.log 2>&1
tail -2 gpurun_out/r1_bench_n1.err
