for probe in 0 1; do for chunk in 0x40000000 0x800000 0x200000 0x80000; do
  STB_PROBE_FIRST=$probe STB_NODE_CHUNK=$chunk python bench.py --no-cpu-baseline --no-e2e --steps 3 --warmup 2 > gpurun_out/exp.json 2>/dev/null
  python -c "
import json; d=json.load(open('gpurun_out/exp.json')); k=d['kernels']; print('probe=$probe chunk=$chunk', round(d['ms_per_step'],2), 'insert', k['node_insert']['ms_per_step'], 'count', k['count_first']['ms_per_step'], 'launches', d['gpu_launches'])"
done; done
