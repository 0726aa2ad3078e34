"""GPU parity tests for the round-2 entry points: compacted host copies, complete STR/STRCD
sectors and file images, the look-ahead of the drop-in encode_sector_str, the single-process
multi-device entries, and the BASELINE configs that were not run before (`spu` sine driven one
block per call, XA on pre-filled buffers). Everything goes through the C ABI and is compared
bit for bit with the unmodified reference build (oracle/_ref) or the oracle port."""
import ctypes as C

import numpy as np
import pytest

import oracle
import psxavenc_b200 as pb
from psxavenc_b200 import sharding, synth

pytestmark = pytest.mark.gpu


def _device_ids(count=2):
    """`count` device ids for the multi-device entries: distinct GPUs when the box has them, else
    the same GPU several times (two encoders + two worker threads on one device — the host-side
    sharding logic is the same)."""
    have = pb.device_count()
    assert have > 0
    return [i % have for i in range(count)]


# ---- host pipeline: compacted copies, ragged budgets, caller's bytes ----------------------------

def test_bs_host_ragged_budgets_keep_callers_tail(restated):
    """Budgets differ inside a chunk, out_stride is larger than every budget and the output
    buffer is pre-filled: [0, budget) must equal the reference (stream + zero padding), bytes at
    and beyond a frame's own budget must keep the caller's content."""
    w, h, n = 64, 48, 23
    frames = np.stack([synth.gen_frame(i, w, h, 1 + i % 6) for i in range(n)])
    sizes = np.array([2016 * (1 + (i * 5) % 4) + (i % 3) for i in range(n)], np.int32)
    sizes[7] = 9           # cannot fit: q = 64, all zero
    sizes[11] = 4          # below the header size
    stride = 2016 * 4 + 64
    enc = pb.BsEncoder(0, w, h, pb.FDCT_ISLOW, max_batch=4)      # 6 chunks over the 3 slots
    out = np.full((n, stride), 0xAA, np.uint8)
    res = np.zeros((n, 4), np.int32)
    failed = enc.encode_host_into(n, frames, sizes, out, stride, res)
    enc.close()
    exp_out, exp_res = restated.bs_encode_batch(0, w, h, frames, sizes, oracle.FDCT_ISLOW, stride=stride)
    assert failed == int((exp_res[:, 2] >= 64).sum()) >= 2
    assert np.array_equal(res[:, 2], exp_res[:, 2])
    for i in range(n):
        s = max(int(sizes[i]), 0)
        if exp_res[i, 2] < 64:
            assert np.array_equal(res[i], exp_res[i])
            assert np.array_equal(out[i, :s], exp_out[i, :s]), "frame %d" % i
        else:
            assert not out[i, :s].any()
        assert (out[i, s:] == 0xAA).all(), "frame %d: bytes beyond its budget were written" % i


def test_bs_large_budget_multi_chunk(restated):
    """Budgets beyond the shared-memory image limit (global-memory bitstream images) over several
    chunks in flight at once: every pipeline slot needs its own image buffer."""
    rng = np.random.default_rng(5)
    n, size = 7, 2016 * 150
    frames = rng.integers(0, 256, size=(n, 320 * 240 * 3 // 2), dtype=np.uint8)
    frames[1] = synth.gen_frame(1, 320, 240, 5)
    frames[4] = synth.gen_frame(4, 320, 240, 3)
    enc = pb.BsEncoder(1, 320, 240, pb.FDCT_ISLOW, max_batch=2)
    got_out, got_res = enc.encode_host(frames, size)
    enc.close()
    exp_out, exp_res = restated.bs_encode_batch(1, 320, 240, frames, size, oracle.FDCT_ISLOW)
    assert np.array_equal(got_res, exp_res)
    assert np.array_equal(got_out, exp_out)


@pytest.mark.parametrize("codec", [pb.CODEC_V2, pb.CODEC_V3], ids=["v2", "v3"])
def test_bs_cluster_mode_matches_single_cta_mode(restated, codec):
    """Calls with very few frames run the pack kernel as a thread-block cluster per frame (the CTAs split
    the plane groups, exchange block lengths and totals through distributed shared memory and OR their
    bitstream images together); the same frames inside a larger call take the one-CTA-per-frame kernels.
    Both must give the oracle's bytes — easy, typical and busy content (the cluster finishes busy frames
    itself, without the census kernel), ragged budgets."""
    w, h = 320, 240
    few = np.stack([synth.gen_frame(i, w, h, nb) for i, nb in enumerate((0, 3, 6, 3, 5, 6))])
    sizes_few = np.array([20160, 20160, 20160, 16128, 18144, 12096], np.int32)
    filler = synth.gen_frames(100, 34, w, h, 3)
    many = np.concatenate([few, filler])
    sizes_many = np.concatenate([sizes_few, np.full(len(filler), 20160, np.int32)])
    enc = pb.BsEncoder(codec, w, h, pb.FDCT_SSE2, max_batch=64)
    out_few, res_few = enc.encode_host(few, sizes_few, stride=20160)          # 6 frames: cluster mode
    out_many, res_many = enc.encode_host(many, sizes_many, stride=20160)      # 40 frames: one CTA per frame
    enc.close()
    exp_out, exp_res = restated.bs_encode_batch(codec, w, h, few, sizes_few, oracle.FDCT_SSE2, stride=20160)
    assert len(set(exp_res[:, 2])) >= 3 and exp_res[:, 2].max() >= 8
    assert np.array_equal(res_few, exp_res) and np.array_equal(res_many[:len(few)], exp_res)
    for i, sz in enumerate(sizes_few):
        assert np.array_equal(out_few[i, :sz], exp_out[i, :sz]), "cluster mode, frame %d" % i
        assert np.array_equal(out_many[i, :sz], exp_out[i, :sz]), "single-CTA mode, frame %d" % i


def test_bs_first_pass_total_exact_for_huge_levels(restated):
    """The q = 1 pass has its own list walk (level = (y + 1) >> 1, rows of the length table clamped at
    level 63): budgets exactly at, just below and just above the q = 1 stream size of a frame made of
    DCT basis sign patterns (levels up to 462, escape codes) and of a noisy frame must leave q = 1
    exactly where the oracle does."""
    w, h = 64, 48
    n = np.arange(8)
    luma = np.zeros((h, w), np.uint8)
    k = 0
    for by in range(h // 8):
        for bx in range(w // 8):
            u, v = (3 * k + 1) % 8, (5 * k + 2) % 8
            basis = np.outer(np.cos((2 * n + 1) * u * np.pi / 16), np.cos((2 * n + 1) * v * np.pi / 16))
            luma[8 * by:8 * by + 8, 8 * bx:8 * bx + 8] = np.where((basis > 0) ^ (k & 1 == 1), 255, 0)
            k += 1
    chroma = np.tile(np.array([[0, 255], [255, 0]], np.uint8), (h // 4, w // 2))
    extreme = np.concatenate([luma.ravel(), chroma.ravel()])
    noisy = synth.gen_frame(5, w, h, 5)
    for frame in (extreme, noisy):
        lo, hi = 8, 200000
        while lo < hi:   # smallest budget that admits q = 1
            mid = (lo + hi) // 2
            _, r = restated.bs_encode_batch(0, w, h, frame[None], mid, oracle.FDCT_SSE2, stride=200000)
            if r[0, 2] <= 1:
                hi = mid
            else:
                lo = mid + 1
        sizes = np.array([lo - 4, lo - 2, lo - 1, lo, lo + 1, lo + 2, lo + 4], np.int32)
        frames = np.repeat(frame[None], len(sizes), axis=0)
        exp_out, exp_res = restated.bs_encode_batch(0, w, h, frames, sizes, oracle.FDCT_SSE2, stride=int(sizes.max()))
        enc = pb.BsEncoder(pb.CODEC_V2, w, h, pb.FDCT_SSE2, max_batch=8)
        got_out, got_res = enc.encode_host(frames, sizes, stride=int(sizes.max()))
        enc.close()
        assert exp_res[:, 2].min() == 1 and 1 < exp_res[:, 2].max() < 64
        assert np.array_equal(got_res, exp_res)
        for i, sz in enumerate(sizes):
            assert np.array_equal(got_out[i, :sz], exp_out[i, :sz]), "budget %d" % sz



@pytest.mark.parametrize("n", [5, 200], ids=["few-frames-640-threads", "many-frames-320-threads"])
@pytest.mark.parametrize("codec", [0, 1], ids=["v2", "v3"])
def test_bs_busy_content_skips_hopeless_scales(restated, n, codec):
    """Content far too busy for the small quant scales: the first pass overruns its budget early
    and the census rules out the scales a lower bound proves hopeless. The result must still be
    the reference's first-fit quant scale and bytes (mdec.c:663-722), for both CTA widths."""
    w, h = 320, 240
    rng = np.random.default_rng(n + codec)
    frames = np.stack([synth.gen_frame(i, w, h, 6 if i % 3 else 5) for i in range(min(n, 10))])
    frames = np.tile(frames, ((n + 9) // 10, 1))[:n].copy()
    frames[:, ::11] ^= rng.integers(0, 8, size=(n, frames[:, ::11].shape[1]), dtype=np.uint8)
    frames[-1] = rng.integers(0, 256, size=frames.shape[1], dtype=np.uint8)            # white noise: q ~ 40
    sizes = np.array([(20160, 18144, 16128, 14112)[i % 4] for i in range(n)], np.int32)
    enc = pb.BsEncoder(codec, w, h, pb.FDCT_SSE2, max_batch=256)
    got_out, got_res = enc.encode_host(frames, sizes, stride=20160)
    enc.close()
    sel = np.arange(n) if n <= 16 else np.unique(np.concatenate([np.arange(0, n, 13), [n - 2, n - 1]]))
    exp_out, exp_res = restated.bs_encode_batch(codec, w, h, frames[sel], sizes[sel], oracle.FDCT_SSE2, stride=20160)
    assert exp_res[:, 2].min() >= 4 and exp_res[:, 2].max() < 64
    assert np.array_equal(got_res[sel], exp_res)
    for k, i in enumerate(sel):
        assert np.array_equal(got_out[i, :sizes[i]], exp_out[k, :sizes[i]]), "frame %d" % i


# ---- complete STR / STRCD sectors and file images ------------------------------------------------

@pytest.mark.parametrize("fmt", [pb.FORMAT_STRCD, pb.FORMAT_STR], ids=["strcd", "str"])
def test_str_complete_video_sectors(reference, fmt):
    """psxb200_str_encode_host_ex with framing: whole 2352/2336-byte video sectors (sync, BCD
    timecode, subheaders, STR header, payload, FORM1 EDC) against the reference's mux loop
    (encode_sector_str + init_sector_buffer_video + psx_cdrom_calculate_checksums) without audio."""
    w, h, n = 320, 240, 9
    frames = synth.gen_frames(3, n, w, h, 3)
    exp, qsum = reference.str_mux(0, w, h, frames, fmt=fmt, fdct=oracle.FDCT_ISLOW, cd_speed=2, fps_num=15, fps_den=1,
                                  xa_file=3, xa_channel=5)
    size = exp.shape[1]
    enc = pb.BsEncoder(0, w, h, pb.FDCT_ISLOW, max_batch=4)
    params = pb.str_params(fmt, 150, 15, framing=1, xa_file=3, xa_channel=5)
    got = np.zeros_like(exp)
    res = enc.str_encode_host_ex(frames, params, got)
    assert res[:, 2].sum() == qsum
    assert np.array_equal(got, exp)
    if fmt == pb.FORMAT_STRCD:
        # bytes the reference never writes (the ECC area behind the EDC) keep the caller's content
        got2 = np.full_like(exp, 0x5A)
        enc.str_encode_host_ex(frames, params, got2)
        assert np.array_equal(got2[:, :0x81C], exp[:, :0x81C]) and (got2[:, 0x81C:] == 0x5A).all()
        # continuation in two calls
        k = 4
        first = int(pb.lib().psxb200_str_sector_count(k, 1, 150, 15))
        got3 = np.zeros_like(exp)
        enc.str_encode_host_ex(frames[:k], params, got3[:first])
        enc.str_encode_host_ex(frames[k:], pb.str_params(fmt, 150, 15, first_frame_index=1 + k, framing=1, xa_file=3, xa_channel=5),
                               got3[first:])
        assert np.array_equal(got3, exp)
    assert size == (2352 if fmt == pb.FORMAT_STRCD else 2336)
    enc.close()


def _strcd_case(n_frames, seed, noise=3, w=320, h=240):
    frames = synth.gen_frames(seed, n_frames, w, h, noise)
    return frames


@pytest.mark.parametrize("fmt,trailing", [(pb.FORMAT_STRCD, False), (pb.FORMAT_STRCD, True), (pb.FORMAT_STR, False)],
                         ids=["strcd", "strcd-trailing-audio", "str"])
def test_strcd_file_images(reference, fmt, trailing):
    """BASELINE config `strcd` end to end on the GPU: psxb200_strcd_encode_host builds the muxed
    .str image (video sectors with budgets 16128,18144,18144,... + one 37800 Hz 4-bit stereo XA
    sector per 8, both produced concurrently) for several files at once; each image must equal
    what the reference's encode_file_str loop writes for that file."""
    w, h, fpf, n_files = 320, 240, 6, 7
    interleave = 8
    enc = pb.BsEncoder(0, w, h, pb.FDCT_ISLOW, max_batch=3)      # 2 files per group -> 4 groups over the 3 slots
    frames = np.concatenate([_strcd_case(fpf, 10 * f, noise=2 + f % 3) for f in range(n_files)])
    params = pb.str_params(fmt, 150 * (interleave - 1), 15 * interleave, interleave=interleave, trailing_audio=int(trailing),
                           xa_file=1, xa_channel=2)
    # how many sectors the video takes decides how much audio the mux loop consumes
    probe, _ = reference.str_mux(0, w, h, frames[:fpf], fmt=fmt, pcm=np.zeros(2 * 2016 * 40, np.int16), n_samples=2016 * 40,
                                 trailing_audio=trailing, xa_channel=2)
    n_sectors = probe.shape[0]
    audio_slot0 = interleave - 1 if trailing else 0
    audio_sectors = len([s for s in range(n_sectors) if s % interleave == audio_slot0])
    samples = audio_sectors * 2016 - 700          # the last audio sector is partial
    pcm = np.stack([synth.gen_pcm(samples + 256, 2, 40 + f)[:samples + 256].ravel() for f in range(n_files)])
    images, res = enc.strcd_encode_host(frames, fpf, params, pcm=pcm, samples_per_file=samples)
    assert images.shape == (n_files, n_sectors * probe.shape[1])
    for f in range(n_files):
        exp, qsum = reference.str_mux(0, w, h, frames[f * fpf:(f + 1) * fpf], fmt=fmt, pcm=pcm[f, :2 * samples + 256],
                                      n_samples=samples, trailing_audio=trailing, xa_channel=2)
        got = images[f].reshape(n_sectors, -1)
        assert exp.shape == got.shape
        bad = [s for s in range(n_sectors) if not np.array_equal(got[s], exp[s])]
        assert not bad, "file %d: sectors %s differ" % (f, bad[:8])
        assert res[f * fpf:(f + 1) * fpf, 2].sum() == qsum
    enc.close()


def test_strcd_device_entries_two_streams(reference):
    """The same image assembled by the caller from the two device entry points on two CUDA
    streams (what bench.py's strcd leg times): psxb200_str_encode_device_ex with LBA placement
    and psxb200_xa_encode_device_ex with a sector stride and LBA step."""
    torch = pytest.importorskip("torch")
    w, h, fpf, interleave = 320, 240, 5, 8
    frames = synth.gen_frames(50, fpf, w, h, 4)
    probe, _ = reference.str_mux(0, w, h, frames, fmt=pb.FORMAT_STRCD, pcm=np.zeros(2 * 2016 * 40, np.int16), n_samples=2016 * 40)
    n_sectors = probe.shape[0]
    audio_sectors = (n_sectors + interleave - 1) // interleave
    samples = audio_sectors * 2016
    pcm = synth.gen_pcm(samples, 2, 9)
    exp, _ = reference.str_mux(0, w, h, frames, fmt=pb.FORMAT_STRCD, pcm=pcm, n_samples=samples)
    enc = pb.BsEncoder(0, w, h, pb.FDCT_ISLOW, max_batch=8)
    params = pb.str_params(pb.FORMAT_STRCD, 150 * 7, 15 * 8, framing=1, interleave=8, place_at_lba=1, xa_file=1, xa_channel=0)
    d_frames = torch.from_numpy(frames).cuda()
    d_pcm = torch.from_numpy(np.concatenate([pcm.ravel(), np.zeros(512, np.int16)])).cuda()
    d_image = torch.zeros((n_sectors, 2352), dtype=torch.uint8, device="cuda")
    d_res = torch.zeros((fpf, 4), dtype=torch.int32, device="cuda")
    d_states = torch.zeros(48, dtype=torch.uint8, device="cuda")
    torch.cuda.synchronize()
    s_video, s_audio = torch.cuda.Stream(), torch.cuda.Stream()
    rc = pb.lib().psxb200_str_encode_device_ex(enc.handle, fpf, d_frames.data_ptr(), C.byref(params), d_image.data_ptr(),
                                               d_res.data_ptr(), s_video.cuda_stream)
    assert rc == 0, pb.last_error()
    rc = pb.lib().psxb200_xa_encode_device_ex(1, 1, 1, 37800, 4, 1, 0, d_pcm.data_ptr(), 0, samples, 0, interleave,
                                              d_states.data_ptr(), d_image.data_ptr(), 0, interleave * 2352, s_audio.cuda_stream)
    assert rc == audio_sectors * 2352, pb.last_error()
    torch.cuda.synchronize()
    got = d_image.cpu().numpy()
    bad = [s for s in range(n_sectors) if not np.array_equal(got[s], exp[s])]
    assert not bad, bad[:8]
    enc.close()


# ---- drop-in encode_sector_str with look-ahead ----------------------------------------------------

def _drive_sector_loop(lib, frames_queue_fn, n_sectors, w, h, num, den, fmt):
    """encode_file_strspu's video loop (filefmt.c:546-630) with the decoder's frame queue semantics:
    encode_sector_str is always handed the queue head, consumed frames are retired by moving the
    rest down (retire_av_data, decoding.c:536-558)."""
    lib.init_mdec_encoder.restype = C.c_bool
    lib.init_mdec_encoder.argtypes = [C.POINTER(pb.MdecEncoder), C.c_int, C.c_int, C.c_int]
    lib.encode_sector_str.restype = C.c_int
    lib.encode_sector_str.argtypes = [C.POINTER(pb.MdecEncoder), C.c_int, C.c_uint16, C.c_void_p, C.c_void_p]
    lib.destroy_mdec_encoder.argtypes = [C.POINTER(pb.MdecEncoder)]
    enc = pb.MdecEncoder()
    assert lib.init_mdec_encoder(C.byref(enc), 0, w, h)
    frame_buf = np.zeros(2016 * 16, np.uint8)
    enc.state.frame_output = frame_buf.ctypes.data_as(C.POINTER(C.c_uint8))
    enc.state.frame_index = 0
    enc.state.frame_data_offset = 0
    enc.state.frame_max_size = 0
    enc.state.frame_block_base_overflow = num
    enc.state.frame_block_overflow_num = 0
    enc.state.frame_block_overflow_den = den
    enc.state.quant_scale_sum = 0
    size = {pb.FORMAT_STRV: 2048, pb.FORMAT_STR: 2336, pb.FORMAT_STRCD: 2352}[fmt]
    out = np.zeros((n_sectors, size), np.uint8)
    queue = frames_queue_fn()
    used_total = 0
    for s in range(n_sectors):
        used = lib.encode_sector_str(C.byref(enc), fmt, 0x8001, queue.ctypes.data, out[s].ctypes.data)
        if used:
            queue[:-used] = queue[used:].copy()      # retire: memmove the queue down
            used_total += used
    hits, misses = C.c_longlong(0), C.c_longlong(0)
    if hasattr(lib, "psxb200_bs_lookahead_stats"):
        lib.psxb200_bs_lookahead_stats(enc.state.dct_context, C.byref(hits), C.byref(misses))
    qsum = enc.state.quant_scale_sum
    lib.destroy_mdec_encoder(C.byref(enc))
    return out, used_total, qsum, hits.value, misses.value


def test_sector_str_lookahead_hits_and_matches_reference(reference, monkeypatch):
    monkeypatch.setenv("PSXB200_FDCT", "sse2")
    monkeypatch.delenv("PSXB200_STR_LOOKAHEAD", raising=False)
    w, h, n = 320, 240, 12
    frames = synth.gen_frames(0, n, w, h, 3)
    n_sectors = int(pb.lib().psxb200_str_sector_count(n, 1, 1050, 120))

    def queue():      # the queue keeps one spare slot behind its frames (decoding.c:448-451)
        return np.concatenate([frames, np.zeros((1, frames.shape[1]), np.uint8)])

    ours, used_a, q_a, hits, misses = _drive_sector_loop(pb.lib(), queue, n_sectors, w, h, 1050, 120, pb.FORMAT_STRV)
    theirs, used_b, q_b, _, _ = _drive_sector_loop(reference.lib, queue, n_sectors, w, h, 1050, 120, pb.FORMAT_STRV)
    assert used_a == used_b == n and q_a == q_b
    assert np.array_equal(ours, theirs)
    assert hits == n - 1 and misses == 0, (hits, misses)


def test_sector_str_lookahead_miss_falls_back(reference, monkeypatch):
    """The frame behind the current one changes before it is needed (a caller that does not keep
    a queue): the speculation must be discarded, the output must still be the reference's."""
    monkeypatch.setenv("PSXB200_FDCT", "sse2")
    w, h, n = 64, 48, 8
    frames = np.stack([synth.gen_frame(i, w, h, 2 + i % 4) for i in range(n)])
    lib = pb.lib()
    enc = pb.MdecEncoder()
    assert lib.init_mdec_encoder(C.byref(enc), 0, w, h)
    frame_buf = np.zeros(2016 * 4, np.uint8)
    enc.state.frame_output = frame_buf.ctypes.data_as(C.POINTER(C.c_uint8))
    for name, val in (("frame_index", 0), ("frame_data_offset", 0), ("frame_max_size", 0), ("frame_block_base_overflow", 3),
                      ("frame_block_overflow_num", 0), ("frame_block_overflow_den", 2), ("quant_scale_sum", 0)):
        setattr(enc.state, name, val)
    slot = np.zeros((2, frames.shape[1]), np.uint8)      # the frame handed in + scratch behind it
    rng = np.random.default_rng(1)
    n_sectors = int(lib.psxb200_str_sector_count(n, 1, 3, 2))
    out = np.zeros((n_sectors, 2048), np.uint8)
    k = 0
    for s in range(n_sectors):
        if enc.state.frame_data_offset >= enc.state.frame_max_size:
            slot[0] = frames[k]
            slot[1] = rng.integers(0, 256, size=frames.shape[1], dtype=np.uint8)     # never the real next frame
            k += 1
        lib.encode_sector_str(C.byref(enc), pb.FORMAT_STRV, 0x8001, slot.ctypes.data, out[s].ctypes.data)
    hits, misses = C.c_longlong(0), C.c_longlong(0)
    lib.psxb200_bs_lookahead_stats(enc.state.dct_context, C.byref(hits), C.byref(misses))
    lib.destroy_mdec_encoder(C.byref(enc))
    exp, used, _ = reference.str_encode(0, w, h, frames, n_sectors, 3, 2, fmt=pb.FORMAT_STRV, fdct=oracle.FDCT_SSE2,
                                        max_frame_size=2016 * 4)
    assert used == n and k == n
    assert np.array_equal(out, exp)
    assert hits.value == 0 and misses.value == n - 1


# ---- one process, several devices -----------------------------------------------------------------

def test_bs_multi_device_host_entry(restated):
    w, h, n = 320, 240, 37
    frames = synth.gen_frames(0, n, w, h, 3)
    sizes = sharding.frame_budgets(n, 1050, 120)
    multi = pb.BsMultiEncoder(0, w, h, pb.FDCT_ISLOW, max_batch=5, device_ids=_device_ids(2))
    assert multi.n_devices == 2
    out = np.full((n, 18144), 0x11, np.uint8)
    res = np.zeros((n, 4), np.int32)
    assert multi.encode_host_into(n, frames, sizes, out, 18144, res) == 0
    exp_out, exp_res = restated.bs_encode_batch(0, w, h, frames, sizes, oracle.FDCT_ISLOW, stride=18144)
    assert np.array_equal(res, exp_res)
    for i in range(n):
        assert np.array_equal(out[i, :sizes[i]], exp_out[i, :sizes[i]]) and (out[i, sizes[i]:] == 0x11).all()
    # STR sectors over the devices: one stream, contiguous frame ranges, sectors land in place
    count = int(pb.lib().psxb200_str_sector_count(n, 1, 1050, 120))
    sectors = np.zeros((count, 2048), np.uint8)
    multi.str_encode_host_ex(frames, pb.str_params(pb.FORMAT_STRV, 1050, 120), sectors)
    single = pb.BsEncoder(0, w, h, pb.FDCT_ISLOW, max_batch=8)
    exp_sectors, _ = single.str_encode_host(frames, 1, 1050, 120, fmt=pb.FORMAT_STRV)
    single.close()
    assert np.array_equal(sectors, exp_sectors)
    multi.close()


def test_strcd_images_over_devices(reference):
    """psxb200_bs_multi_strcd_encode_host: whole files dealt out over the devices."""
    w, h, fpf, n_files, interleave = 320, 240, 4, 5, 8
    frames = synth.gen_frames(7, fpf * n_files, w, h, 3)
    params = pb.str_params(pb.FORMAT_STRCD, 150 * 7, 15 * 8, interleave=interleave, xa_file=1, xa_channel=0)
    probe, _ = reference.str_mux(0, w, h, frames[:fpf], fmt=pb.FORMAT_STRCD, pcm=np.zeros(2 * 2016 * 40, np.int16), n_samples=2016 * 40)
    n_sectors = probe.shape[0]
    samples = ((n_sectors + interleave - 1) // interleave) * 2016
    pcm = np.stack([np.concatenate([synth.gen_pcm(samples, 2, 90 + f).ravel(), np.zeros(256, np.int16)]) for f in range(n_files)])
    size = int(pb.lib().psxb200_strcd_image_bytes(C.byref(params), fpf, 4, 1, samples))
    assert size == n_sectors * 2352
    images = np.zeros((n_files, size), np.uint8)
    res = np.zeros((fpf * n_files, 4), np.int32)
    multi = pb.BsMultiEncoder(0, w, h, pb.FDCT_ISLOW, max_batch=4, device_ids=_device_ids(2))
    rc = pb.lib().psxb200_bs_multi_strcd_encode_host(multi.handle, n_files, fpf, frames.ctypes.data, C.byref(params), 37800, 4, 1,
                                                     pcm.ctypes.data, pcm.shape[1], samples, None, images.ctypes.data, size,
                                                     res.ctypes.data)
    assert rc == 0, pb.last_error()
    for f in range(n_files):
        exp, qsum = reference.str_mux(0, w, h, frames[f * fpf:(f + 1) * fpf], fmt=pb.FORMAT_STRCD, pcm=pcm[f], n_samples=samples)
        assert np.array_equal(images[f], exp.ravel()), "file %d" % f
        assert res[f * fpf:(f + 1) * fpf, 2].sum() == qsum
    multi.close()


def test_spu_multi_device_channel_split(restated):
    """`vagi`: one 8-channel file, chain c on device c mod G; and `vagi x B`: whole files per device."""
    ch, count = 8, 3584 + 280
    pcm = synth.gen_pcm(count, ch, 21)
    got, states = pb.spu_encode_host_multi(pcm, ch, ch, 0, count, device_ids=_device_ids(3))
    for c in range(ch):
        st = oracle.ChannelState()
        exp = restated.spu_encode(st, pcm, count, ch, offset=c)
        assert np.array_equal(got[c], exp), "channel %d" % c
        assert (states[c].prev1, states[c].prev2) == (st.prev1, st.prev2)
    files = 5
    many = np.stack([synth.gen_pcm(count, ch, 30 + f) for f in range(files)])
    got, states = pb.spu_encode_host_multi(many, files * ch, ch, count * ch, count, device_ids=_device_ids(2))
    for f in range(files):
        for c in range(ch):
            st = oracle.ChannelState()
            assert np.array_equal(got[f * ch + c], restated.spu_encode(st, many[f], count, ch, offset=c)), (f, c)
            assert states[f * ch + c].prev1 == st.prev1


@pytest.mark.parametrize("count,files,boundary", [(28 * 20 + 5, 1000, 464), (28 * 256 + 5, 260, 128)], ids=["short-chains", "long-chains"])
def test_spu_host_call_cut_into_pipelined_runs(restated, count, files, boundary):
    """A host call with thousands of whole interleave groups is cut into runs of groups that rotate over
    several streams (capi_audio.cu): every chain must come out as if encoded alone, including the
    partial last group, a buffer that ends with the last chain's last sample, a gap between the groups
    and non-zero incoming states."""
    ch = 8
    stride = count * ch + 24                      # groups do not touch: a few unused samples between them
    rng = np.random.default_rng(5)
    pcm = rng.integers(-32768, 32768, size=files * stride, dtype=np.int16)
    n_streams = files * ch - 3                    # last group holds 5 of its 8 chains
    pcm = pcm[:(files - 1) * stride + (count - 1) * ch + 5]   # and the buffer ends with the last chain's last sample
    states = (pb.ChannelState * n_streams)()
    for s in range(0, n_streams, 7):
        states[s].prev1, states[s].prev2 = int(rng.integers(-3000, 3000)), int(rng.integers(-3000, 3000))
    before = [(st.prev1, st.prev2) for st in states]
    launches = pb.launch_count()
    got, states = pb.spu_encode_host(pcm, n_streams, ch, stride, count, states=states)
    assert pb.launch_count() - launches >= 2, "expected the call to be cut into several launches"
    edge = boundary * ch                          # first chain of the second run
    for s in list(range(0, n_streams, 131)) + [edge - 1, edge, 2 * edge - 1, 2 * edge, n_streams - 1]:
        f, c = divmod(s, ch)
        st = oracle.ChannelState()
        st.prev1, st.prev2 = before[s]
        exp = restated.spu_encode(st, pcm[f * stride:], count, ch, offset=c)
        assert np.array_equal(got[s], exp), "chain %d" % s
        assert (states[s].prev1, states[s].prev2) == (st.prev1, st.prev2), "state of chain %d" % s


# ---- BASELINE config `spu`: the sine, one block per call ---------------------------------------------

def test_spu_sine_driven_like_encode_file_spu(reference):
    """configs[0]: mono 22050 Hz 440 Hz sine through psx_audio_spu_encode the way encode_file_spu
    drives it (filefmt.c:212-292): a leading zero block, then ONE <= 28-sample block per call with
    the loop flags patched in, a trailing LOOP_TRAP block, padding to 64 bytes."""
    seconds = 4
    pcm = synth.gen_sine(22050 * seconds + 13)       # ragged last block

    def encode_file_spu(lib):
        lib.psx_audio_spu_encode.restype = C.c_int
        lib.psx_audio_spu_encode.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p]
        state = pb.ChannelState()
        blocks = [np.zeros(16, np.uint8)]
        block = np.zeros(16, np.uint8)
        done = 0
        while done < len(pcm):
            length = min(28, len(pcm) - done)
            n = lib.psx_audio_spu_encode(C.addressof(state), pcm[done:].ctypes.data, length, 1, block.ctypes.data)
            assert n == 16
            done += length
            b = block.copy()
            if done >= len(pcm):
                b[1] |= 1            # PSX_AUDIO_SPU_LOOP_END (filefmt.c:251-254, no loop point)
            blocks.append(b)
        trap = np.zeros(16, np.uint8)
        trap[1] = 5                  # PSX_AUDIO_SPU_LOOP_TRAP (filefmt.c:271-278)
        blocks.append(trap)
        data = np.concatenate(blocks)
        pad = (-len(data)) % 64
        return np.concatenate([data, np.zeros(pad, np.uint8)]), (state.prev1, state.prev2)

    ours, st_a = encode_file_spu(pb.lib())
    theirs, st_b = encode_file_spu(reference.lib)
    assert st_a == st_b
    assert np.array_equal(ours, theirs)
    assert len(ours) % 64 == 0 and len(ours) >= 16 * (len(pcm) // 28 + 2)


# ---- XA on pre-filled output buffers ----------------------------------------------------------------

@pytest.mark.parametrize("fmt", [0, 1], ids=["xa", "xacd"])
@pytest.mark.parametrize("stereo", [False, True], ids=["mono", "stereo"])
@pytest.mark.parametrize("bits", [4, 8])
def test_xa_prefilled_output_buffers(reference, fmt, stereo, bits):
    """The reference leaves some bytes of the output alone (bytes 8-15 of 8-bit sound groups are
    copied 8-11 -> 12-15 from whatever was there, adpcm.c:322; the 2336-byte format ORs the coding
    byte, adpcm.c:277-288; the 20 pad bytes of a group-less tail) and computes the EDC over them.
    filefmt.c:453 hands in an uninitialised stack buffer — so compare on 0xAA and random fills."""
    per_sector = (112 if bits == 8 else 224) // (2 if stereo else 1) * 18
    count = per_sector * 3 - 333
    ch = 2 if stereo else 1
    pcm = np.concatenate([synth.gen_pcm(count, ch, 60 + bits + fmt).ravel(), np.zeros(600, np.int16)])
    size = 2336 if fmt == 0 else 2352
    rng = np.random.default_rng(bits * 4 + fmt * 2 + int(stereo))
    for fill in (np.full(3 * size, 0xAA, np.uint8), rng.integers(0, 256, size=3 * size, dtype=np.uint8)):
        a, b = fill.copy(), fill.copy()
        sa, sb = pb.EncoderState(), oracle.new_states()
        ours = pb.XaSettings(fmt, stereo, 37800, bits, 1, 2)
        na = pb.lib().psx_audio_xa_encode(ours, C.addressof(sa), pcm.ctypes.data, count, 11, a.ctypes.data)
        got_b = reference.xa_encode(fmt, stereo, 37800, bits, 1, 2, sb, pcm, count, 11, out=b)
        assert na == len(got_b) == 3 * size
        assert np.array_equal(a, b), np.nonzero(a != b)[0][:16]
        assert (sa.left.prev1, sa.left.prev2, sa.right.prev1, sa.right.prev2) == (sb[0].prev1, sb[0].prev2, sb[1].prev1, sb[1].prev2)


def test_xa_batch_prefilled_and_multi_device(reference):
    """The batched XA entry on pre-filled buffers, over two devices."""
    n_streams, count = 5, 2016 * 2 + 100
    pcm = np.stack([np.concatenate([synth.gen_pcm(count, 2, 80 + s).ravel(), np.zeros(512, np.int16)]) for s in range(n_streams)])
    size = 3 * 2352
    out = np.full((n_streams, size), 0x3C, np.uint8)
    states = (pb.EncoderState * n_streams)()
    ids = _device_ids(2)
    rc = pb.lib().psxb200_xa_encode_host_multi(2, (C.c_int * 2)(*ids), n_streams, 1, 1, 37800, 4, 1, 0, pcm.ctypes.data,
                                               pcm.shape[1], count, 5, C.addressof(states), out.ctypes.data, size)
    assert rc == size, pb.last_error()
    for s in range(n_streams):
        exp = np.full(size, 0x3C, np.uint8)
        st = oracle.new_states()
        reference.xa_encode(1, True, 37800, 4, 1, 0, st, pcm[s], count, 5, out=exp)
        assert np.array_equal(out[s], exp), "stream %d" % s
        assert states[s].right.prev1 == st[1].prev1
