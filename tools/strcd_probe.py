"""e2e of psxb200_strcd_encode_host for different group sizes (frames per pipeline chunk)."""
import os, sys, time, ctypes as C
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import psxavenc_b200 as pb
from psxavenc_b200 import synth

fpf, files, samples = 8, 512, 20160
n = fpf * files
frames = torch.from_numpy(np.tile(synth.gen_frames(0, 64, 320, 240, 3), (n // 64, 1))).pin_memory()
pcm_one = np.concatenate([synth.gen_pcm(samples, 2, 5).ravel(), np.zeros(256, np.int16)])
pcm = torch.from_numpy(np.tile(pcm_one, (files, 1))).pin_memory()
params = pb.str_params(pb.FORMAT_STRCD, 1050, 120, framing=1, interleave=8, place_at_lba=1, xa_file=1, frames_per_file=fpf)
size = int(pb.lib().psxb200_strcd_image_bytes(C.byref(params), fpf, 4, 1, samples))
images = torch.empty((files, size), dtype=torch.uint8, pin_memory=True)
res = torch.empty((n, 4), dtype=torch.int32, pin_memory=True)
ref = None
for group in (128, 256, 512, 1024, 2048):
    os.environ["PSXB200_STRCD_GROUP_FRAMES"] = str(group)
    enc = pb.BsEncoder(0, 320, 240, pb.FDCT_SSE2, max_batch=256)

    def call():
        rc = pb.lib().psxb200_strcd_encode_host(enc.handle, files, fpf, frames.data_ptr(), C.byref(params), 37800, 4, 1, pcm.data_ptr(),
                                                pcm.shape[1], samples, None, images.data_ptr(), size, res.data_ptr())
        assert rc == 0, pb.last_error()
    call()
    t0 = time.perf_counter()
    for _ in range(5):
        call()
    dt = time.perf_counter() - t0
    if ref is None:
        ref = images.clone()
    print("group %4d frames: %8.0f frames/s  (%.1f GB/s host->device)  same bytes: %s" % (group, 5 * n / dt, 5 * n * 117720 / dt / 1e9, bool(torch.equal(ref, images))), flush=True)
    enc.close()
