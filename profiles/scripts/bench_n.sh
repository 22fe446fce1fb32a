# bench.py at N GPUs exactly as the driver launches it; JSON line -> gpurun_out/r01_bench_n$N.json
N=$1
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus $N --steps 5 --warmup 3 > gpurun_out/r01_bench_n$N.json 2> gpurun_out/r01_bench_n$N.err
python - <<PY
import json
lines = [l for l in open("gpurun_out/r01_bench_n$N.json") if l.strip()]
assert len(lines) == 1, lines
d = json.loads(lines[0])
print("N=$N", round(d["value"], 1), "Gbp/s", round(d["ms_per_step"], 2), "ms; e2e", round(d["e2e"]["value"], 1), "Gbp/s", round(d["e2e"]["ms_per_step"], 2), "ms;", d["collectives_per_step"], "collectives;", d["config"]["sharding"])
print(d["clocks"], d["gpu_launches"], d["roofline"])
PY
