for s in 21 32 42; do
  python bench.py --steps 2 --warmup 3 --no-cpu-baseline --samples $s 2>/dev/null | python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print('samples',d['config'].get('chains', d['config']), 'value %.1f M e2e %.1f M ms %.1f'%(d['value']/1e6, d['e2e']['value']/1e6, d['ms_per_step']))
"
done
CUDA_DEVICE_MAX_CONNECTIONS=64 python bench.py --steps 2 --warmup 3 --no-cpu-baseline --samples 42 2>/dev/null | python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print('conn64 value %.1f M e2e %.1f M'%(d['value']/1e6, d['e2e']['value']/1e6))
"
