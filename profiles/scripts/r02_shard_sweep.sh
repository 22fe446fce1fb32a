#!/bin/bash
# usage: r02_shard_sweep.sh N "<opts>" ...
N=$1; shift
for o in "$@"; do
  args=""
  if [ "$o" != "default" ]; then for kv in $o; do args="$args --option $kv"; done; fi
  echo "== N=$N $o"
  python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29551 bench.py --gpus $N --steps 10 --warmup 3 --no-e2e $args 2>/dev/null | python -c "
import json,sys
for line in sys.stdin:
    if line.startswith('{'):
        d=json.loads(line); print(round(d['ms_per_step'],3),'ms', {k:v['ms_per_step'] for k,v in d['kernels'].items()}, d['parity'].get('equal_to_reference'))
"
done
