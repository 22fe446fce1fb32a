"""torchrun target: the sharded build over NCCL on real GPUs against the single-GPU stream.
    python -m torch.distributed.run --nproc-per-node N tests/shard_gpu_check.py [n_bases]"""
import hashlib
import os
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))


def main():
    import torch
    import torch.distributed as dist
    from conftest import load_package
    n_bases = int(sys.argv[1]) if len(sys.argv) > 1 else 60_000_000
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    stb = load_package()
    from genome_compression_b200 import shard
    me = shard.create_nccl(12, device=local)
    first, count = me.range(n_bases)
    body = torch.empty(max(count, 16), dtype=torch.uint8, device="cuda")
    if count:
        stb.synth_genome(body, n_bases, first=first, count=count, seed=5, repeat_permille=500, device=local)
    for _ in range(3):  # repeated builds on one handle reuse the arenas
        me.build_from_body(body, n_bases)
    tree = me.gather()
    if rank == 0:
        full = torch.empty(n_bases, dtype=torch.uint8, device="cuda")
        stb.synth_genome(full, n_bases, seed=5, repeat_permille=500, device=local)
        single = stb.SharedTree(12, device=local).build_from_body(full)
        assert tree.layer_counts() == single.layer_counts(), (tree.layer_counts()[:4], single.layer_counts()[:4])
        a, b = tree.serialize(), single.serialize()
        assert hashlib.sha256(a).digest() == hashlib.sha256(b).digest(), "sharded stream differs from the single-GPU stream"
        tree.sort()
        single.sort()
        assert tree.serialize() == single.serialize()
        print(f"shard_gpu_check ok: world {world}, {n_bases} bases, totals {me.layer_totals()[:4]}", flush=True)
    dist.barrier()
    me.close()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
