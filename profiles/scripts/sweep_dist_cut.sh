for cut in 24 22 20 18; do
STB_DIST_CUT_LOG2=$cut timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus 2 --no-e2e > gpurun_out/b.json 2> gpurun_out/b.err
python -c "
import json,sys; d=json.loads([l for l in open('gpurun_out/b.json') if l.startswith('{')][-1]); print('cut', $cut, round(d['value'],1), round(d['ms_per_step'],2), d['collectives_per_step'], round(sum(v['ms_per_step'] for v in d['kernels'].values()),2))"
done
STB_DIST_TRACE=1 timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus 2 --steps 1 --warmup 3 --no-e2e > gpurun_out/trace_n2.txt 2> gpurun_out/trace_n2.err; grep "dist\]" gpurun_out/trace_n2.txt | tail -36
