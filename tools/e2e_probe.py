"""e2e of the strv host entry under different host-side settings, one process per GPU (torchrun) —
what limits the 8-GPU end-to-end number? Each setting: 4096 frames per GPU per step, 6 steps."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, torch.distributed as dist
import psxavenc_b200 as pb
from psxavenc_b200 import synth

rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
n, fb, size = 4096, 115200, 20160
frames = np.tile(synth.gen_frames(rank * 64, 64, 320, 240, 3), (n // 64, 1))
h_frames = torch.from_numpy(frames).pin_memory()
h_sizes = torch.full((n,), size, dtype=torch.int32).pin_memory()
h_out = torch.empty((n, size), dtype=torch.uint8, pin_memory=True)
h_res = torch.empty((n, 4), dtype=torch.int32, pin_memory=True)


def barrier():
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
        torch.cuda.synchronize()


def run(label, env):
    for k in ("PSXB200_EXPERIMENT_SKIP_TAIL_ZERO", "PSXB200_HOST_CHUNK"):
        os.environ.pop(k, None)
    os.environ.update(env)
    enc = pb.BsEncoder(0, 320, 240, pb.FDCT_SSE2, max_batch=int(env.get("MAX_BATCH", 256)))
    enc.encode_host_into(n, h_frames, h_sizes, h_out, size, h_res)
    barrier()
    t0 = time.perf_counter()
    for _ in range(6):
        enc.encode_host_into(n, h_frames, h_sizes, h_out, size, h_res)
    dt = time.perf_counter() - t0
    t = torch.tensor([dt], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    enc.close()
    if rank == 0:
        v = world * n * 6 / t.item()
        print("%-44s %9.0f frames/s  %6.1f GB/s host->device in total" % (label, v, v * fb / 1e9), flush=True)


run("default (256-frame chunks, 3 slots)", {})
run("tail zero fill skipped (experiment)", {"PSXB200_EXPERIMENT_SKIP_TAIL_ZERO": "1"})
run("128-frame chunks", {"PSXB200_HOST_CHUNK": "128"})
run("1024-frame chunks", {"MAX_BATCH": "1024", "PSXB200_HOST_CHUNK": "1024"})
# bare copies: all ranks at once, frames in only / frames in + bitstreams out
d_in = torch.empty((n, fb), dtype=torch.uint8, device="cuda")
d_out = torch.empty((n, size), dtype=torch.uint8, device="cuda")
s2 = torch.cuda.Stream()
for label, both in (("bare cudaMemcpy host->device only", False), ("bare host->device + device->host (20160 B/frame)", True)):
    barrier()
    t0 = time.perf_counter()
    for _ in range(6):
        d_in.copy_(h_frames, non_blocking=True)
        if both:
            with torch.cuda.stream(s2):
                h_out.copy_(d_out, non_blocking=True)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    t = torch.tensor([dt], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    if rank == 0:
        v = world * n * 6 / t.item()
        print("%-44s %9.0f frames/s  %6.1f GB/s host->device in total" % (label, v, v * fb / 1e9), flush=True)
if world > 1:
    dist.destroy_process_group()
