"""Run under torchrun (one rank per GPU, NCCL): sharded build of a synthetic sequence, gathered
on rank 0 and compared with the single-GPU build of the same sequence (and the oracle when small).

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 \
        tests/dist_gpu_check.py [n_bases]
"""
import os
import sys
from pathlib import Path

import torch
import torch.distributed as dist

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))
from conftest import load_package  # noqa: E402


def main():
    n_bases = int(sys.argv[1]) if len(sys.argv) > 1 else 60_000_000
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    pkg = load_package()
    from genome_compression_b200.dist import CudaStages, DistBuilder, ShardPlan

    S = 12
    n_leaves = n_bases // S
    plan = ShardPlan(n_leaves, world)
    lo, hi = plan.level_range(rank, 0)
    body = torch.empty(max(16, (hi - lo) * S), dtype=torch.uint8, device="cuda")
    if hi > lo:
        pkg.synth_genome(body, n_bases, first=lo * S, count=(hi - lo) * S, seed=9, repeat_permille=500, device=local)
    builder = DistBuilder(CudaStages(pkg, S, local))
    tree = builder.build_from_body(body, n_bases)
    full = builder.gather(tree)
    if rank == 0:
        whole = torch.empty(n_bases, dtype=torch.uint8, device="cuda")
        pkg.synth_genome(whole, n_bases, seed=9, repeat_permille=500, device=local)
        single = pkg.SharedTree(S, device=local).build_from_body(whole)
        assert full.layer_counts() == single.layer_counts(), (full.layer_counts(), single.layer_counts())
        assert full.leaf_count() == single.leaf_count() and full.root() == single.root() and full.width() == single.width()
        assert full.serialize() == single.serialize(), "pre-sort stream differs from the single-GPU build"
        full.sort()
        single.sort()
        assert full.serialize() == single.serialize(), "post-sort stream differs from the single-GPU build"
        print(f"dist_gpu_check ok: world={world} bases={n_bases} leaves={full.leaf_count()} nodes={full.node_count()} "
              f"collectives={builder.collectives}", flush=True)
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
