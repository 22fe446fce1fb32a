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
    single = single_pre = single_post = None
    if rank == 0:
        whole = torch.empty(n_bases, dtype=torch.uint8, device="cuda")
        pkg.synth_genome(whole, n_bases, seed=9, repeat_permille=500, device=local)
        single = pkg.SharedTree(S, device=local).build_from_body(whole)
        single_pre = single.serialize()
        single.sort()
        single_post = single.serialize()
        del whole
    # several exchanged node levels (the default cut would leave only the leaf level sharded here);
    # two builds per builder: the second one reuses the peer-mapped exchange memory
    for exchange in ("peer", "collective"):
        builder = DistBuilder(CudaStages(pkg, S, local), cut=1 << 12, exchange=exchange)
        for _ in range(2):
            tree = builder.build_from_body(body, n_bases)
        full = builder.gather(tree)
        if rank == 0:
            check(world, n_bases, exchange, builder, full, single, single_pre, single_post)
        builder.close()
    dist.barrier()
    dist.destroy_process_group()


def check(world, n_bases, exchange, builder, full, single, single_pre, single_post):
    assert full.layer_counts() == single.layer_counts(), (full.layer_counts(), single.layer_counts())
    assert full.leaf_count() == single.leaf_count() and full.root() == single.root() and full.width() == single.width()
    assert full.serialize() == single_pre, "pre-sort stream differs from the single-GPU build"
    full.sort()
    assert full.serialize() == single_post, "post-sort stream differs from the single-GPU build"
    print(f"dist_gpu_check ok: world={world} exchange={exchange} bases={n_bases} leaves={full.leaf_count()} "
          f"nodes={full.node_count()} collectives={builder.collectives}", flush=True)


if __name__ == "__main__":
    main()
