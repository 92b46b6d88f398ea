python bench.py > gpurun_out/r1_bench_n1.json 2> gpurun_out/r1_bench_n1.err
tail -2 gpurun_out/r1_bench_n1.err
python bench.py --impl reference > gpurun_out/r1_bench_reference_arm.json 2> gpurun_out/r1_bench_ref.err
python -c "
import json
d=json.load(open('gpurun_out/r1_bench_n1.json')); print(d['value'], d['e2e']['value'], d['ms_per_step'], d['config']['samples'], d['rates']['single_chain_cells_per_s'], d['clocks'])
r=json.load(open('gpurun_out/r1_bench_reference_arm.json')); print('ref', r['value'])
"
