"""Per-kernel time of the BS path against the number of frames per launch (wave quantisation / tail probe).

usage: python tools/wave_probe.py [noise_bits] [n ...]
The pack kernel runs one CTA per frame, four CTAs per SM: 592 frames are one full wave on 148 SMs.
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import psxavenc_b200 as pb
from psxavenc_b200 import synth

noise = int(sys.argv[1]) if len(sys.argv) > 1 else 3
counts = [int(a) for a in sys.argv[2:]] or [592, 1184, 2368, 3552, 4096, 4144, 4736, 8192, 8288]
dev = torch.device("cuda", 0)
W, H, BUDGET = 320, 240, 20160
base = torch.from_numpy(synth.gen_frames(0, 256, W, H, noise)).to(dev)
stream = torch.cuda.current_stream()
for n in counts:
    frames = base.repeat((n + 255) // 256, 1)[:n].contiguous()
    out = torch.zeros((n, BUDGET), dtype=torch.uint8, device=dev)
    res = torch.zeros((n, 4), dtype=torch.int32, device=dev)
    enc = pb.BsEncoder(pb.CODEC_V2, W, H, pb.FDCT_SSE2, max_batch=n)
    step = lambda: enc.encode_device(n, frames, None, BUDGET, out, BUDGET, res, stream.cuda_stream)
    for _ in range(3):
        step()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(stream)
    for _ in range(20):
        step()
    b.record(stream)
    torch.cuda.synchronize()
    whole = a.elapsed_time(b) / 20
    enc.timing(True)
    enc.read_timing()
    for _ in range(10):
        step()
    torch.cuda.synchronize()
    dct_ms, pack_ms, pairs = enc.read_timing()
    enc.close()
    print("n=%5d (%.2f waves)  step %.4f ms  %.3f M frames/s   dct %.4f  pack %.4f   ns/frame: dct %.1f pack %.1f"
          % (n, n / 592.0, whole, n / whole / 1e3, dct_ms / pairs, pack_ms / pairs, dct_ms / pairs / n * 1e6, pack_ms / pairs / n * 1e6),
          flush=True)
    del frames, out, res
