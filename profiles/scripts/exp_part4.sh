# partitioned node levels v3 (bucket-ordered records, per-bucket table regions prefetched into L2, epoch tags)
for cfg in "0x7fffffff 0x40000" "0x200000 0x40000" "0x200000 0x10000" "0x200000 0x100000" "0x800000 0x40000"; do
  set -- $cfg
  STB_PART_MIN=$1 STB_PART_BUCKET=$2 python bench.py --no-cpu-baseline --no-e2e --steps 3 --warmup 2 > gpurun_out/exp.json 2>/dev/null
  python -c "
import json; d=json.load(open('gpurun_out/exp.json')); k=d['kernels']; g=lambda n: k.get(n,{}).get('ms_per_step',0)
print('min=$1 bucket=$2', round(d['ms_per_step'],2), 'hist', g('part_hist'), 'scan', g('part_scan'), 'scat', g('part_scatter'), 'ins', g('bucket_insert'), 'ans', g('bucket_answer'), 'clear', g('table_clear'), 'node_insert', g('node_insert'), 'count', g('count_first'), 'assign', g('assign_ids'), 'resolve', g('resolve_ids'))"
done
