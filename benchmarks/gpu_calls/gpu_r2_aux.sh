#!/bin/bash
# round 2, the "next" rows on the B200: tare / sparsify (n4) and the query path (n3) timed, the differ
# kernels' DRAM traffic, then the default bench line of the final tree
mkdir -p gpurun_out
timeout 120 python benchmarks/differ_bench.py 1000000 > gpurun_out/r2_differ_bench.json 2> gpurun_out/r2_differ_bench.err; echo "differ rc=$?"
timeout 60 python benchmarks/differ_bench.py 1000000 oracle > gpurun_out/r2_differ_bench_oracle.json 2>/dev/null; echo "differ oracle rc=$?"
timeout 120 python benchmarks/query_bench.py 100000 1000000 > gpurun_out/r2_query_bench.json 2> gpurun_out/r2_query_bench.err; echo "query rc=$?"
timeout 60 python benchmarks/query_bench.py 100000 100000 oracle > gpurun_out/r2_query_bench_oracle.json 2>/dev/null; echo "query oracle rc=$?"
timeout 150 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:k_differ --csv --log-file gpurun_out/r2_differ_dram.csv python benchmarks/differ_bench.py 1000000 > /dev/null 2>&1; echo "ncu rc=$?"
SECONDS=0
timeout 330 python bench.py > gpurun_out/r2_bench_n1_final.json 2> gpurun_out/r2_bench_n1_final.err; echo "bench rc=$? wall ${SECONDS}s"
cut -c1-300 gpurun_out/r2_differ_bench.json; cut -c1-300 gpurun_out/r2_query_bench.json
