# partitioned node levels v2 (staged scatter + cooperative tagged tables): bucket size x L2 budget
for cfg in "0x7fffffff 0x40000 48" "0x200000 0x40000 48" "0x200000 0x40000 96" "0x200000 0x20000 48" "0x200000 0x80000 64" "0x200000 0x10000 32"; do
  set -- $cfg
  STB_PART_MIN=$1 STB_PART_BUCKET=$2 STB_PART_L2_MB=$3 python bench.py --no-cpu-baseline --no-e2e --steps 3 --warmup 2 > gpurun_out/exp.json 2>/dev/null
  python -c "
import json; d=json.load(open('gpurun_out/exp.json')); k=d['kernels']; g=lambda n: k.get(n,{}).get('ms_per_step',0)
print('min=$1 bucket=$2 l2=$3', round(d['ms_per_step'],2), 'hist', g('part_hist'), 'scan', g('part_scan'), 'scat', g('part_scatter'), 'proc', g('bucket_process'), 'clear', g('table_clear'), 'node_insert', g('node_insert'), 'count', g('count_first'), 'assign', g('assign_ids'), 'resolve', g('resolve_ids'))"
done
