#!/bin/bash
# usage: r02_sweep.sh "<opts1>" "<opts2>" ...   each a space-separated list of name=value (or "default")
B="python bench.py --steps 5 --warmup 2 --no-cpu-baseline --no-e2e --no-pipeline"
for o in "$@"; do
  args=""
  if [ "$o" != "default" ]; then for kv in $o; do args="$args --option $kv"; done; fi
  echo "== $o"
  $B $args 2>&1 | python -c "
import json,sys
for line in sys.stdin:
    if line.startswith('{'):
        d=json.loads(line); print(round(d['ms_per_step'],3),'ms', d['gpu_launches'], {k:v['ms_per_step'] for k,v in d['kernels'].items()})
    else: print(line.rstrip()[:300])
"
done
