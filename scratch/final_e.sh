python bench.py --impl reference > gpurun_out/r1_bench_reference_arm.json 2> gpurun_out/r1_bench_ref.err
python bench.py > gpurun_out/r1_bench_n1.json 2> gpurun_out/r1_bench_n1.err
python -c "
import json
r=json.load(open('gpurun_out/r1_bench_reference_arm.json')); print('ref', r['value'], r['cpu_baseline']['sample'][-160:])
d=json.load(open('gpurun_out/r1_bench_n1.json')); print(d['value'], d['e2e']['value'], d['cpu_baseline']['value'], d['clocks']['reasons'])"
