#!/bin/bash
# per-level kernel classes on every rank (why does shard_apply grow with the rank?)  usage: r02_apply_probe.sh N
N=$1
run() { tag=$1; shift; STB_RANK_PROFILES=$tag python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29551 bench.py --gpus $N --no-e2e "$@" > gpurun_out/r02_probe_$tag.json 2> gpurun_out/r02_probe_$tag.err; grep "stb shard" gpurun_out/r02_probe_$tag.err | sort | uniq | head -80 > gpurun_out/r02_probe_${tag}_answers.txt; }
run lvl --steps 5 --warmup 2 --option profile_levels=1
run cnt --steps 1 --warmup 1 --option profile_levels=2
python - <<PY
import json,glob
for tag in ("lvl",):
    rows=[json.load(open(f"gpurun_out/rank_profile_{tag}_n$N"+f"_r{r}.json")) for r in range($N)]
    keys=[]
    for r in rows:
        for k in r:
            if k not in keys: keys.append(k)
    for k in keys:
        if "apply" in k or "dedup" in k or "partition" in k or "collective" in k or "resolve" in k:
            print("%-28s"%k+"".join("%8.3f"%row.get(k,0) for row in rows))
PY
head -12 gpurun_out/r02_probe_cnt_answers.txt
