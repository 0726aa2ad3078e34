"""ADPCM throughput on one GPU (device-resident): SPU `vagi x B` and XA stereo streams."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import psxavenc_b200 as pb
from psxavenc_b200 import synth

dev = torch.device("cuda", 0)
stream = torch.cuda.current_stream()

def timed(fn, steps=10):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(stream)
    for _ in range(steps): fn()
    b.record(stream)
    torch.cuda.synchronize()
    return a.elapsed_time(b) / steps

def spu(files, ch=8, count=3584 * 4):
    base = synth.gen_pcm(count, ch, 7)
    d_pcm = torch.from_numpy(base).to(dev).unsqueeze(0).repeat(files, 1, 1).contiguous()
    d_pcm += torch.randint(-64, 64, d_pcm.shape, dtype=torch.int16, device=dev)
    streams = files * ch
    row = 16 * (count // 28)
    d_states = torch.zeros((streams, 24), dtype=torch.uint8, device=dev)
    d_out = torch.zeros((streams, row), dtype=torch.uint8, device=dev)
    def step():
        d_states.zero_()
        assert pb.lib().psxb200_spu_encode_device(streams, d_pcm.data_ptr(), ch, count * ch, count, None, d_states.data_ptr(),
                                                  d_out.data_ptr(), row, stream.cuda_stream) == 0
    ms = timed(step)
    print("SPU  files=%5d chains=%6d  %.3f ms  %.1f Gsamples/s" % (files, streams, ms, streams * count / ms / 1e6), flush=True)

def xa(n, stereo=True, bits=4, sectors=8):
    ch = 2 if stereo else 1
    per = ((112 if bits == 8 else 224) >> (1 if stereo else 0)) * 18
    count = per * sectors
    base = synth.gen_pcm(count + 224, ch, 9)
    d_pcm = torch.from_numpy(base).to(dev).unsqueeze(0).repeat(n, 1, 1).contiguous()
    d_states = torch.zeros((n, 48), dtype=torch.uint8, device=dev)
    d_out = torch.zeros((n, sectors * 2352), dtype=torch.uint8, device=dev)
    def step():
        d_states.zero_()
        rc = pb.lib().psxb200_xa_encode_device(n, 1, int(stereo), 37800, bits, 1, 0, d_pcm.data_ptr(), (count + 224) * ch, count, 0,
                                               d_states.data_ptr(), d_out.data_ptr(), sectors * 2352, stream.cuda_stream)
        assert rc == sectors * 2352, pb.last_error()
    ms = timed(step)
    print("XA   streams=%5d stereo=%d bits=%d  %.3f ms  %.1f Gsamples/s" % (n, stereo, bits, ms, n * count * ch / ms / 1e6), flush=True)

if __name__ == "__main__":
    for files in (1, 16, 128, 1024, 4096):
        spu(files)
    for n in (1, 256, 4096, 16384):
        xa(n)
    xa(4096, stereo=False)
    xa(4096, bits=8)
