"""torchrun microbenchmark: all_to_all_single bandwidth per rank vs payload, int64 records."""
import os, time, torch, torch.distributed as dist
local = int(os.environ["LOCAL_RANK"]); torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
world, rank = dist.get_world_size(), dist.get_rank()
for mb in (8, 64, 512):
    n = mb * (1 << 20) // 8
    send = torch.empty(n, dtype=torch.int64, device="cuda").random_()
    recv = torch.empty_like(send)
    per = n // world
    splits = [per] * (world - 1) + [n - per * (world - 1)]
    for name, kw in (("equal", {}), ("splits", {"output_split_sizes": splits, "input_split_sizes": splits})):
        for _ in range(3):
            dist.all_to_all_single(recv, send, **kw)
        torch.cuda.synchronize(); dist.barrier(); t0 = time.perf_counter()
        for _ in range(10):
            dist.all_to_all_single(recv, send, **kw)
        torch.cuda.synchronize(); dt = (time.perf_counter() - t0) / 10
        if rank == 0:
            print(f"world={world} {mb:4d} MB/rank {name:6s} {dt*1e3:7.3f} ms  {mb/1024/dt:7.1f} GB/s per rank (incl. self part)", flush=True)
    # all_reduce for reference
    t = torch.zeros(mb * (1 << 20) // 4, dtype=torch.int32, device="cuda")
    for _ in range(3): dist.all_reduce(t)
    torch.cuda.synchronize(); dist.barrier(); t0 = time.perf_counter()
    for _ in range(10): dist.all_reduce(t)
    torch.cuda.synchronize(); dt = (time.perf_counter() - t0) / 10
    if rank == 0: print(f"world={world} {mb:4d} MB all_reduce {dt*1e3:7.3f} ms", flush=True)
dist.destroy_process_group()
