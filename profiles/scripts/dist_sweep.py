"""torchrun tool: one process group, one synthetic shard per rank, then the sharded build timed
for several settings (cut, exchange) back to back — amortises start-up when GPU time is short.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29561 \
        profiles/scripts/dist_sweep.py --cuts 24,22,20 [--exchange peer,collective] [--bases 3100000000] [--trace]
"""
import argparse
import os
import sys
from pathlib import Path

import torch
import torch.distributed as dist

ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT))
from __graft_entry__ import load_package  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--bases", type=int, default=3_100_000_000)
    ap.add_argument("--cuts", default="24")
    ap.add_argument("--exchange", default="peer")
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--trace", action="store_true")
    args = ap.parse_args()
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    pkg = load_package()
    from genome_compression_b200.dist import CudaStages, DistBuilder, ShardPlan

    S = 12
    n_leaves = args.bases // S
    plan = ShardPlan(n_leaves, world)
    lo, hi = plan.level_range(rank, 0)
    body = torch.empty(max(16, (hi - lo) * S), dtype=torch.uint8, device="cuda")
    if hi > lo:
        pkg.synth_genome(body, args.bases, first=lo * S, count=(hi - lo) * S, seed=42, repeat_permille=500, device=local)
    stream = torch.cuda.current_stream()
    for exchange in args.exchange.split(","):
        for cut in [int(c) for c in args.cuts.split(",")]:
            stages = CudaStages(pkg, S, local, stream=stream.cuda_stream)
            builder = DistBuilder(stages, cut=1 << cut, exchange=exchange)
            for _ in range(3):
                tree = builder.build_from_body(body, args.bases)
            torch.cuda.synchronize()
            dist.barrier()
            stages.ctx.profile(True)
            stages.ctx.profile_reset()
            ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            ev0.record(stream)
            for _ in range(args.steps):
                tree = builder.build_from_body(body, args.bases)
            ev1.record(stream)
            torch.cuda.synchronize()
            ms = torch.tensor([ev0.elapsed_time(ev1) / args.steps], device="cuda", dtype=torch.float64)
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
            prof = stages.ctx.profile_read()
            stages.ctx.profile(False)
            if rank == 0:
                ksum = sum(r["ms"] for r in prof.values()) / args.steps
                top = sorted(prof.items(), key=lambda kv: -kv[1]["ms"])[:8]
                print(f"world={world} exchange={exchange} cut=2^{cut}: {float(ms.item()):.2f} ms = "
                      f"{n_leaves * S / float(ms.item()) / 1e6:.1f} Gbp/s; sharded levels {plan.__class__(n_leaves, world, 1 << cut).sharded_levels()}; "
                      f"rank-0 stage kernels {ksum:.2f} ms: " + ", ".join(f"{k} {v['ms'] / args.steps:.2f}" for k, v in top), flush=True)
            if args.trace:
                builder.trace = True
                builder.build_from_body(body, args.bases)
                builder.trace = False
            del tree
            builder.close()
            dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
