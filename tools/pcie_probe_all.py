"""Pinned-memory copy bandwidth with EVERY GPU of the box copying at once (one process per GPU under torchrun): the ceiling of
the multi-GPU e2e number.  Prints one line per rank and the aggregate."""
import os, time
import torch
import torch.distributed as dist
rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1")); local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
n = 1 << 30
h1 = torch.empty(n, dtype=torch.uint8, pin_memory=True); h2 = torch.empty(n, dtype=torch.uint8, pin_memory=True)
d1 = torch.empty(n, dtype=torch.uint8, device="cuda"); d2 = torch.empty(n, dtype=torch.uint8, device="cuda")
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
def bar():
    torch.cuda.synchronize()
    if world > 1: dist.barrier()
def run(f, reps=8):
    f(); bar(); t0 = time.perf_counter()
    for _ in range(reps): f()
    torch.cuda.synchronize(); dt = (time.perf_counter() - t0) / reps; bar(); return dt
def h2d():
    with torch.cuda.stream(s1): d1.copy_(h1, non_blocking=True)
def d2h():
    with torch.cuda.stream(s2): h2.copy_(d2, non_blocking=True)
def both(): h2d(); d2h()
res = [n / run(f) / 1e9 for f in (h2d, d2h, both)]
t = torch.tensor(res, device="cuda", dtype=torch.float64)
if world > 1:
    allr = [torch.zeros_like(t) for _ in range(world)]; dist.all_gather(allr, t)
else:
    allr = [t]
if rank == 0:
    for r, x in enumerate(allr):
        print("rank %d: H2D %.1f GB/s, D2H %.1f GB/s, both at once %.1f GB/s each direction" % (r, *x.tolist()))
    s = torch.stack(allr).sum(0).tolist()
    print("all %d GPUs at once: H2D %.1f GB/s, D2H %.1f GB/s, both at once %.1f GB/s each direction; host cpus %d" % (world, *s, os.cpu_count()))
if world > 1: dist.destroy_process_group()
