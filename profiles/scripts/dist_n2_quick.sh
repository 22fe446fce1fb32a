# quick 2-GPU check of the sharded build: parity (dist_gpu_check: both exchanges, two builds each) + bench line
set -e
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29551 tests/dist_gpu_check.py 120000000 2>&1 | grep "dist_gpu_check ok"
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus 2 --no-e2e "$@" > gpurun_out/b.json 2> gpurun_out/b.err
python - <<'PY'
import json
d = json.loads([l for l in open('gpurun_out/b.json') if l.startswith('{')][-1])
print('N=2', round(d['value'], 1), 'Gbp/s', round(d['ms_per_step'], 2), 'ms', d['collectives_per_step'], 'collectives')
for k, v in sorted(d['kernels'].items(), key=lambda kv: -kv[1]['ms_per_step'])[:12]:
    print(f"  {k:28s} {v['ms_per_step']:.3f}")
PY
