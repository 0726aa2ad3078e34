"""Latency of the DROP-IN symbols (one frame / one audio call at a time, host buffers), i.e. what
the unmodified psxavenc mux loops see when linked against libpsxav_b200.so:

  encode_frame_bs             one synchronous frame per call (encode_file_sbs, filefmt.c:643)
  encode_sector_str           the strv sector loop with the decoder's frame queue (filefmt.c:572-613),
                              with and without the look-ahead (PSXB200_STR_LOOKAHEAD=0)
  psx_audio_spu_encode        3584-sample interleave chunks (filefmt.c:335) and single 28-sample
                              blocks (filefmt.c:243)
  psx_audio_xa_encode         one 2352-byte sector per call (filefmt.c:487)
"""
import ctypes as C, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import psxavenc_b200 as pb
from psxavenc_b200 import synth

lib = pb.lib()
w, h, n = 320, 240, 2000
frames = synth.gen_frames(0, 64, w, h, 3)


def new_encoder():
    enc = pb.MdecEncoder()
    assert lib.init_mdec_encoder(C.byref(enc), pb.CODEC_V2, w, h)
    buf = np.zeros(20160, np.uint8)
    enc.state.frame_output = buf.ctypes.data_as(C.POINTER(C.c_uint8))
    enc.state.frame_max_size = 20160
    enc.state.quant_scale_sum = 0
    return enc, buf


enc, buf = new_encoder()
for i in range(50):
    lib.encode_frame_bs(C.byref(enc), frames[i % 64].ctypes.data)
t0 = time.perf_counter()
for i in range(n):
    lib.encode_frame_bs(C.byref(enc), frames[i % 64].ctypes.data)
dt = time.perf_counter() - t0
print("encode_frame_bs drop-in: %.1f us/frame, %.0f frames/s (pageable host memory)" % (dt / n * 1e6, n / dt))
lib.destroy_mdec_encoder(C.byref(enc))


def sector_loop(lookahead, n_frames=600):
    os.environ["PSXB200_STR_LOOKAHEAD"] = "1" if lookahead else "0"
    enc, buf = new_encoder()
    for name, val in (("frame_index", 0), ("frame_data_offset", 0), ("frame_max_size", 0), ("frame_block_base_overflow", 150),
                      ("frame_block_overflow_num", 0), ("frame_block_overflow_den", 15)):
        setattr(enc.state, name, val)
    # the decoder's queue: two frames + the spare slot; consumed frames are retired by moving the rest down
    queue = np.zeros((3, frames.shape[1]), np.uint8)
    queue[0], queue[1] = frames[0], frames[1]
    nxt = 2
    sector = np.zeros(2048, np.uint8)
    t0 = time.perf_counter()
    for s in range(n_frames * 10):
        used = lib.encode_sector_str(C.byref(enc), pb.FORMAT_STRV, 0x8001, queue.ctypes.data, sector.ctypes.data)
        if used:
            queue[0] = queue[1]
            queue[1] = frames[nxt % 64]
            nxt += 1
    dt = time.perf_counter() - t0
    hits, misses = C.c_longlong(0), C.c_longlong(0)
    lib.psxb200_bs_lookahead_stats(enc.state.dct_context, C.byref(hits), C.byref(misses))
    lib.destroy_mdec_encoder(C.byref(enc))
    print("encode_sector_str drop-in, look-ahead %s: %.1f us/frame amortised over its 10 sector calls (%.0f frames/s), %d hits %d misses"
          % ("on " if lookahead else "off", dt / n_frames * 1e6, n_frames / dt, hits.value, misses.value))


sector_loop(False)
sector_loop(True)

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

mono = synth.gen_sine(28 * 20000)
st = pb.ChannelState()
blk = np.zeros(16, np.uint8)
sp, base, op = C.addressof(st), mono.ctypes.data, blk.ctypes.data
t0 = time.perf_counter()
for b in range(20000):
    lib.psx_audio_spu_encode(sp, base + 56 * b, 28, 1, op)
dt = time.perf_counter() - t0
print("psx_audio_spu_encode drop-in (one 28-sample block per call): %.1f us/call, %.2f Msamples/s" % (dt / 20000 * 1e6, 20000 * 28 / dt / 1e6))

xa = pb.XaSettings(1, True, 37800, 4, 1, 0)
stereo = np.concatenate([synth.gen_pcm(2016 * 200, 2, 5).ravel(), np.zeros(512, np.int16)])
state = pb.EncoderState()
sector = np.zeros(2352, np.uint8)
t0 = time.perf_counter()
for k in range(200):
    lib.psx_audio_xa_encode(xa, C.addressof(state), stereo.ctypes.data + 4 * 2016 * k, 2016, k, sector.ctypes.data)
dt = time.perf_counter() - t0
print("psx_audio_xa_encode drop-in (one 2352-byte sector per call): %.1f us/call, %.2f Msamples/s" % (dt / 200 * 1e6, 200 * 4032 / dt / 1e6))

# the same sector loop from plain C (no Python between the calls), with and without caller-side work
import subprocess, tempfile
import oracle
oracle.build()
if os.path.exists(oracle.DROPIN_DRIVER):
    with tempfile.TemporaryDirectory() as tmp:
        path = os.path.join(tmp, "in.nv21")
        frames.tofile(path)
        for work in ("0", "150"):
            print(subprocess.run([oracle.DROPIN_DRIVER, "strvbench", str(w), str(h), "1500", work, path], capture_output=True,
                                 text=True, timeout=300).stdout.strip())
