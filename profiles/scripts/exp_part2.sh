# partitioned node levels, all buckets of a level in (almost) one launch: do tables stay hot by block order?
for bucket in 0x10000 0x40000 0x100000; do
  STB_PART_MIN=0x200000 STB_PART_BUCKET=$bucket STB_PART_L2_MB=100000 python bench.py --no-cpu-baseline --no-e2e --steps 3 --warmup 2 > gpurun_out/exp.json 2>/dev/null
  python -c "
import json; d=json.load(open('gpurun_out/exp.json')); k=d['kernels']; g=lambda n: k.get(n,{}).get('ms_per_step',0)
print('bucket=$bucket', round(d['ms_per_step'],2), 'hist', g('part_hist'), 'scat', g('part_scatter'), 'ins', g('bucket_insert'), 'ans', g('bucket_answer'), 'clear', g('table_clear'), 'launches', d['gpu_launches'])"
done
