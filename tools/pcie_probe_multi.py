"""Aggregate D2H bandwidth of N GPUs copying to pinned host memory at the same time (one rank per GPU,
torchrun): what bounds bench.py's e2e at N > 1.  Prints one line on rank 0."""
import os, time, torch, torch.distributed as dist
rank, world, lr = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(lr)
dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
n = 4 * 1024**3
d = torch.empty(n, dtype=torch.uint8, device="cuda")
h = torch.empty(n, dtype=torch.uint8).pin_memory()
res = []
for rep in range(3):
    torch.cuda.synchronize(); dist.barrier(); torch.cuda.synchronize()
    t = time.perf_counter()
    h.copy_(d, non_blocking=True)
    torch.cuda.synchronize()
    dt = torch.tensor([time.perf_counter() - t], dtype=torch.float64, device="cuda")
    dist.all_reduce(dt, op=dist.ReduceOp.MAX)
    res.append(world * n / float(dt[0]) / 1e9)
if rank == 0:
    print(f"D2H, {world} GPUs at once, 4 GiB each to pinned host memory: aggregate {max(res):.1f} GB/s ({max(res)/world:.1f} per GPU); cpus {os.cpu_count()}")
dist.barrier(); dist.destroy_process_group()
