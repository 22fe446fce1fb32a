#!/bin/bash
# usage: r02_bench_n.sh N [extra bench args]   -> gpurun_out/r02_n$N.json
N=$1; shift
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29551 bench.py --gpus $N --steps 10 --warmup 3 "$@" > gpurun_out/r02_n$N.json 2> gpurun_out/r02_n$N.err
tail -c 400 gpurun_out/r02_n$N.err
python - <<PY
import json
d=json.loads(open('gpurun_out/r02_n$N.json').read())
print('N=$N', round(d['value'],1),'Gbp/s', round(d['ms_per_step'],3),'ms', 'e2e', round(d['e2e']['ms_per_step'],2) if d['e2e'] else None, 'copy', d['e2e']['h2d_copy_alone_ms'] if d['e2e'] else None, d['parity'].get('equal_to_reference'))
print({k:v['ms_per_step'] for k,v in d['kernels'].items()}, round(sum(v['ms_per_step'] for v in d['kernels'].values()),3))
PY
