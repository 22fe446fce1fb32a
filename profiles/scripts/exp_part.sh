# partitioned node levels: bucket size x L2 budget sweep (3.1 Gbp, N=1)
for min in 0x7fffffff 0x200000; do for bucket in 0x20000 0x40000 0x80000; do for l2 in 32 64; do
  STB_PART_MIN=$min STB_PART_BUCKET=$bucket STB_PART_L2_MB=$l2 python bench.py --no-cpu-baseline --no-e2e --steps 3 --warmup 2 > gpurun_out/exp.json 2>/dev/null
  python -c "
import json; d=json.load(open('gpurun_out/exp.json')); k=d['kernels']; g=lambda n: k.get(n,{}).get('ms_per_step',0)
print('min=$min bucket=$bucket l2=$l2', round(d['ms_per_step'],2), 'hist', g('part_hist'), 'scat', g('part_scatter'), 'ins', g('bucket_insert'), 'ans', g('bucket_answer'), 'clear', g('table_clear'), 'node_insert', g('node_insert'), 'count', g('count_first'), 'launches', d['gpu_launches'])"
  [ $min = 0x7fffffff ] && break 2
done; done; done
