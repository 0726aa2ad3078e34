"""Latency of the DROP-IN symbols (one frame / one audio call at a time, host buffers), i.e. what
the unmodified psxavenc mux loops would see when linked against libpsxav_b200.so."""
import ctypes as C, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import psxavenc_b200 as pb
from psxavenc_b200 import synth

lib = pb.lib()
w, h, n = 320, 240, 2000
frames = synth.gen_frames(0, 64, w, h, 3)
enc = pb.MdecEncoder()
assert lib.init_mdec_encoder(C.byref(enc), pb.CODEC_V2, w, h)
buf = np.zeros(20160, np.uint8)
enc.state.frame_output = buf.ctypes.data_as(C.POINTER(C.c_uint8))
enc.state.frame_max_size = 20160
enc.state.quant_scale_sum = 0
for i in range(50):
    lib.encode_frame_bs(C.byref(enc), frames[i % 64].ctypes.data)
t0 = time.perf_counter()
for i in range(n):
    lib.encode_frame_bs(C.byref(enc), frames[i % 64].ctypes.data)
dt = time.perf_counter() - t0
print("encode_frame_bs drop-in: %.1f us/frame, %.0f frames/s (pageable host memory)" % (dt / n * 1e6, n / dt))
lib.destroy_mdec_encoder(C.byref(enc))

pcm = synth.gen_pcm(3584 * 200, 8, 3)
states = [pb.ChannelState() for _ in range(8)]
out = np.zeros(2048, np.uint8)
t0 = time.perf_counter()
calls = 0
for chunk in range(100):
    for ch in range(8):
        lib.psx_audio_spu_encode(C.addressof(states[ch]), pcm[chunk * 3584:].ctypes.data + 2 * ch, 3584, 8, out.ctypes.data)
        calls += 1
dt = time.perf_counter() - t0
print("psx_audio_spu_encode drop-in (3584-sample chunks, pitch 8): %.1f us/call, %.2f Msamples/s" % (dt / calls * 1e6, calls * 3584 / dt / 1e6))
