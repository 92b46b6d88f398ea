timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -2
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
python bench.py > gpurun_out/r1_bench_n1.json 2> gpurun_out/r1_bench_n1.err
python -c "
import json
d=json.load(open('gpurun_out/r1_bench_n1.json')); print(d['value'], d['e2e']['value'], d['ms_per_step'], d['config']['samples'], d['rates']['single_chain_cells_per_s'], d['clocks']['reasons'], d['cpu_baseline']['value']); print({k:(v.get('ms') if isinstance(v,dict) else v) for k,v in d['roofline_aux'].items()}); print(d['roofline']['traffic'])"
