"""Kernel durations on busy content (noise_bits 6): run under
ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum -k regex:bs_ to see the common and the
BUSY pack kernel separately."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import psxavenc_b200 as pb
from psxavenc_b200 import synth

noise = int(sys.argv[1]) if len(sys.argv) > 1 else 6
n = 4096
base = synth.gen_frames(0, 256, 320, 240, noise)
d_frames = torch.from_numpy(np.tile(base, (n // 256, 1))).cuda()
d_out = torch.zeros((n, 20160), dtype=torch.uint8, device="cuda")
d_res = torch.zeros((n, 4), dtype=torch.int32, device="cuda")
enc = pb.BsEncoder(0, 320, 240, pb.FDCT_SSE2, max_batch=n)
for _ in range(4):
    enc.encode_device(n, d_frames, None, 20160, d_out, 20160, d_res, None)
torch.cuda.synchronize()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record()
for _ in range(5):
    enc.encode_device(n, d_frames, None, 20160, d_out, 20160, d_res, None)
b.record()
torch.cuda.synchronize()
q = d_res[:, 2].cpu().numpy()
print("noise %d: %.3f ms per 4096 frames, quant scales %s" % (noise, a.elapsed_time(b) / 5, dict(zip(*np.unique(q, return_counts=True)))))
