for loc in 0 1; do
  STB_LOCALITY=$loc python bench.py --no-cpu-baseline --no-e2e --steps 3 --warmup 2 > gpurun_out/exp.json 2>/dev/null
  python -c "
import json; d=json.load(open('gpurun_out/exp.json')); k=d['kernels']; g=lambda n: k.get(n,{}).get('ms_per_step',0)
print('locality=$loc', round(d['ms_per_step'],2), 'node_insert', g('node_insert'), 'count', g('count_first'), 'assign', g('assign_ids'), 'resolve', g('resolve_ids'), 'leaf', g('leaf_insert'), 'clear', g('table_clear'))"
done
