# usage: bash scratch/quick_bench.sh [samples...]   -> one line per setting
for s in "$@"; do
  python bench.py --steps 2 --warmup 3 --no-cpu-baseline --samples $s 2>/dev/null | python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print('samples', d['config']['samples'], 'value %.1f M  e2e %.1f M  single %.2f M'%(d['value']/1e6, d['e2e']['value']/1e6, d['rates']['single_chain_cells_per_s']/1e6))
"
done
