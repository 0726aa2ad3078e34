"""Concurrent host<->device copy bandwidth of N ranks (torchrun): the platform ceiling of bench.py's e2e."""
import os, time, torch, torch.distributed as dist
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
n = 472 * 1000 * 1000
h = torch.empty(n, dtype=torch.uint8, pin_memory=True); h.fill_(1)
d = torch.empty(n, dtype=torch.uint8, device="cuda")
h2 = torch.empty(n // 5, dtype=torch.uint8, pin_memory=True)
d2 = torch.empty(n // 5, dtype=torch.uint8, device="cuda")
s2 = torch.cuda.Stream()
def run(both):
    torch.cuda.synchronize(); dist.barrier(); torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(5):
        d.copy_(h, non_blocking=True)
        if both:
            with torch.cuda.stream(s2):
                h2.copy_(d2, non_blocking=True)
    torch.cuda.synchronize()
    return 5 * n / (time.perf_counter() - t0) / 1e9
for both in (False, True):
    run(both)
    g = torch.tensor([run(both)], device="cuda")
    out = [torch.zeros_like(g) for _ in range(world)]
    dist.all_gather(out, g)
    if rank == 0:
        print("world %d %s: H2D GB/s per rank %s  total %.1f" % (world, "H2D+D2H" if both else "H2D only", ["%.1f" % o.item() for o in out], sum(o.item() for o in out)), flush=True)
dist.destroy_process_group()
