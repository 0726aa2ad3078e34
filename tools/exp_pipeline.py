"""Experiment: L2-sized chunks issued round-robin on several streams (each with its own
coefficient plane) vs one big launch pair. Prints ms per 4096-frame step."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import psxavenc_b200 as pb
import bench

n = 4096
dev = torch.device("cuda", 0)
frames = torch.from_numpy(bench.make_frames(n, 0)).to(dev)
sizes = torch.full((n,), 20160, dtype=torch.int32, device=dev)
out = torch.zeros((n, 20160), dtype=torch.uint8, device=dev)
res = torch.zeros((n, 4), dtype=torch.int32, device=dev)

def run(chunk, nstreams, steps=20):
    encs = [pb.BsEncoder(0, 320, 240, 1, max_batch=chunk) for _ in range(nstreams)]
    streams = [torch.cuda.Stream() for _ in range(nstreams)]
    def step():
        for k, first in enumerate(range(0, n, chunk)):
            m = min(chunk, n - first)
            s = streams[k % nstreams]
            encs[k % nstreams].encode_device(m, frames[first:].data_ptr(), sizes[first:].data_ptr(), 20160,
                                             out[first:].data_ptr(), 20160, res[first:].data_ptr(), s.cuda_stream)
    for _ in range(3): step()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(steps): step()
    torch.cuda.synchronize()
    ms = (time.perf_counter() - t0) * 1000 / steps
    for e in encs: e.close()
    return ms

for chunk, ns in [(4096, 1), (1024, 2), (512, 2), (512, 4), (256, 2), (256, 4), (256, 6), (128, 4), (128, 8), (64, 8), (148, 4), (148, 6)]:
    ms = run(chunk, ns)
    print("chunk=%4d streams=%d  ms/step=%.3f  fps=%.0f  plane_MB=%.0f" % (chunk, ns, ms, n / ms * 1000, chunk * ns * 262656 / 1e6), flush=True)
