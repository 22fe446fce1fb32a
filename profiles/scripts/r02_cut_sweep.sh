#!/bin/bash
# usage: r02_cut_sweep.sh N cut1 cut2 ...   (cut = 0: the default)  -> one line per cut
N=$1; shift
for cut in "$@"; do
  opt=""; [ "$cut" != "0" ] && opt="--option cut=$cut"
  python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29551 bench.py --gpus $N --steps 6 --warmup 3 --no-e2e $opt > gpurun_out/r02_cut_n${N}_$cut.json 2> gpurun_out/r02_cut_n${N}_$cut.err
  python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/r02_cut_n${N}_$cut.json').read())
    k={a:round(b['ms_per_step'],3) for a,b in d['kernels'].items()}
    print('N=$N cut=$cut', round(d['ms_per_step'],3),'ms', round(d['value'],1),'Gbp/s', d['parity'].get('equal_to_reference'), k, flush=True)
except Exception as e:
    print('N=$N cut=$cut failed', e, open('gpurun_out/r02_cut_n${N}_$cut.err').read()[-600:])
PY
done
