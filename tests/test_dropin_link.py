"""Link-level drop-in: tests/dropin_driver.c is compiled against the REFERENCE's own headers
(mdec.h, libpsxav.h) and linked with libpsxav_b200.so in place of mdec.c/adpcm.c. The CPU test
proves it compiles and links (declarations and struct layouts agree); the GPU tests run it the
way the reference's mux loops drive the boundary and compare the files with the oracle."""
import os
import subprocess

import numpy as np
import pytest

import oracle
from psxavenc_b200 import build as pb_build, synth


@pytest.fixture(scope="module")
def driver():
    pb_build.build()
    oracle.build()
    if not os.path.exists(oracle.DROPIN_DRIVER):
        pytest.skip("dropin_driver not built (needs /root/reference headers)")
    return oracle.DROPIN_DRIVER


def test_driver_links_against_product_library(driver):
    out = subprocess.run(["ldd", driver], capture_output=True, text=True).stdout
    assert "libpsxav_b200.so" in out and "not found" not in out.split("libpsxav_b200.so")[1].splitlines()[0]
    syms = subprocess.run(["nm", "-D", "--undefined-only", driver], capture_output=True, text=True).stdout
    for name in ("init_mdec_encoder", "encode_frame_bs", "encode_sector_str", "destroy_mdec_encoder", "psx_audio_spu_encode"):
        assert name in syms


@pytest.mark.gpu
def test_driver_sbs_loop(driver, restated, tmp_path):
    """encode_file_sbs: fixed 8192-byte frames, BS v3, islow FDCT (the drop-in default)."""
    w, h, n = 640, 480, 3
    frames = np.stack([synth.gen_smooth_frame(i, w, h) for i in range(n)])
    frames.tofile(tmp_path / "in.nv21")
    subprocess.run([driver, "sbs", str(w), str(h), "1", "8192", str(n), str(tmp_path / "in.nv21"), str(tmp_path / "out.bin")],
                   check=True, timeout=120)
    got = np.fromfile(tmp_path / "out.bin", dtype=np.uint8).reshape(n, 8192)
    exp, res = restated.bs_encode_batch(1, w, h, frames, 8192, oracle.FDCT_ISLOW)
    assert (res[:, 2] < 64).all()
    assert np.array_equal(got, exp)


@pytest.mark.gpu
def test_driver_strv_loop(driver, reference, tmp_path):
    """encode_file_strspu video branch at strcd-like 1050/120 sectors per frame, vs the
    reference's own encode_sector_str."""
    w, h, n_frames, sectors = 320, 240, 4, 35
    frames = synth.gen_frames(0, n_frames, w, h, 4)
    frames.tofile(tmp_path / "in.nv21")
    subprocess.run([driver, "strv", str(w), str(h), "0", "1050", "120", str(sectors), str(tmp_path / "in.nv21"),
                    str(tmp_path / "out.bin")], check=True, timeout=120)
    got = np.fromfile(tmp_path / "out.bin", dtype=np.uint8).reshape(sectors, 2048)
    exp, used, _ = reference.str_encode(0, w, h, frames, sectors, 1050, 120, fmt=9, fdct=oracle.FDCT_ISLOW)
    assert used == n_frames
    assert np.array_equal(got, exp)


@pytest.mark.gpu
def test_driver_spui_loop(driver, restated, tmp_path):
    """encode_file_spui: 8 channels, 2048-byte interleave, leading dummy block, ragged tail."""
    ch, interleave, total = 8, 2048, 3584 * 2 + 1000
    pcm = synth.gen_pcm(total, ch, 12)
    pcm.tofile(tmp_path / "in.pcm")
    subprocess.run([driver, "spui", str(ch), str(interleave), str(total), str(tmp_path / "in.pcm"), str(tmp_path / "out.bin")],
                   check=True, timeout=120)
    got = np.fromfile(tmp_path / "out.bin", dtype=np.uint8)
    states = [oracle.ChannelState() for _ in range(ch)]
    chunks, done, k = [], 0, 0
    while done < total:
        length = min(3584, total - done)
        chunk = np.zeros((ch, interleave), np.uint8)
        skip = 0
        if k == 0:
            skip, length = 16, length - 28
        for c in range(ch):
            blk = restated.spu_encode(states[c], pcm[done:], length, ch, offset=c)
            chunk[c, skip:skip + len(blk)] = blk
        chunks.append(chunk.ravel())
        done += length
        k += 1
    assert np.array_equal(got, np.concatenate(chunks))


@pytest.mark.gpu
def test_driver_multi_device_entry(driver, restated, tmp_path):
    """A plain C host (compiled against the reference's headers plus the additive psxb200_* layer)
    driving two devices from one process through psxb200_bs_multi_encode_host with pinned buffers
    from psxb200_pinned_alloc; output and result rows byte-equal to the oracle."""
    w, h, n, size = 320, 240, 41, 18144
    frames = synth.gen_frames(5, n, w, h, 3)
    frames.tofile(tmp_path / "in.nv21")
    subprocess.run([driver, "multi", str(w), str(h), "1", str(size), str(n), "2", str(tmp_path / "in.nv21"), str(tmp_path / "out.bin")],
                   check=True, timeout=120)
    raw = np.fromfile(tmp_path / "out.bin", dtype=np.uint8)
    got = raw[:n * size].reshape(n, size)
    res = raw[n * size:].view(np.int32).reshape(n, 4)
    exp, exp_res = restated.bs_encode_batch(1, w, h, frames, size, oracle.FDCT_ISLOW)
    assert np.array_equal(res, exp_res)
    assert np.array_equal(got, exp)


def _mux_strcd(lib, frames, pcm, n_sectors, w, h):
    """The sector schedule of encode_file_str (filefmt.c:391-520) for -t strcd defaults: 2x speed,
    15 fps, 37800 Hz 4-bit stereo XA -> interleave 8 (1 audio + 7 video sectors), driven through
    the drop-in symbols of `lib` (ours or the reference build). Sector init / EDC of video
    sectors belong to filefmt.c + cdrom.c (out of scope) and are left out on both sides."""
    import ctypes as C
    import psxavenc_b200 as pb
    settings_cls = pb.XaSettings
    xa = settings_cls(1, True, 37800, 4, 1, 0)                      # XACD sectors
    interleave, video_per_block, per_sector = 8, 7, 2016
    enc = pb.MdecEncoder()
    lib.init_mdec_encoder.restype = C.c_bool
    lib.init_mdec_encoder.argtypes = [C.POINTER(pb.MdecEncoder), C.c_int, C.c_int, C.c_int]
    lib.encode_sector_str.restype = C.c_int
    lib.encode_sector_str.argtypes = [C.POINTER(pb.MdecEncoder), C.c_int, C.c_uint16, C.c_void_p, C.c_void_p]
    lib.destroy_mdec_encoder.argtypes = [C.POINTER(pb.MdecEncoder)]
    lib.psx_audio_xa_encode.restype = C.c_int
    lib.psx_audio_xa_encode.argtypes = [settings_cls, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p]
    assert lib.init_mdec_encoder(C.byref(enc), 0, w, h)
    enc.state.frame_block_base_overflow = 150 * video_per_block * 1      # (75 * speed) * video sectors * fps_den
    enc.state.frame_block_overflow_den = interleave * 15
    frame_buf = np.zeros(2016 * 9, np.uint8)
    enc.state.frame_output = frame_buf.ctypes.data_as(C.POINTER(C.c_uint8))
    enc.state.frame_index = 0
    enc.state.frame_data_offset = 0
    enc.state.frame_max_size = 0
    enc.state.frame_block_overflow_num = 0
    enc.state.quant_scale_sum = 0
    audio_state = pb.EncoderState()
    out = np.zeros((n_sectors, 2352), np.uint8)
    # one spare slot behind the frames, as in the decoder's queue (decoding.c:448-451)
    frames = np.concatenate([frames, np.zeros((1, frames.shape[1]), np.uint8)])
    frame_pos, sample_pos = 0, 0
    for s in range(n_sectors):
        if s % interleave > 0:       # video sector (filefmt.c:458-461)
            frame_pos += lib.encode_sector_str(C.byref(enc), pb.FORMAT_STRCD, 0x8001, frames[frame_pos:].ctypes.data,
                                               out[s].ctypes.data)
        else:
            n = lib.psx_audio_xa_encode(xa, C.addressof(audio_state), pcm[sample_pos:].ctypes.data, per_sector, s,
                                        out[s].ctypes.data)
            assert n == 2352
            sample_pos += per_sector
    qsum = enc.state.quant_scale_sum
    lib.destroy_mdec_encoder(C.byref(enc))
    return out, frame_pos, qsum


@pytest.mark.gpu
def test_strcd_mux_through_dropin_symbols(reference, monkeypatch):
    """BASELINE config `strcd`: 320x240 v2 video with budgets 16128,18144,18144,... interleaved with
    37800 Hz 4-bit stereo XA, muxed by the same loop through our symbols and the reference's."""
    import psxavenc_b200 as pb
    monkeypatch.setenv("PSXB200_FDCT", "sse2")      # the reference build here runs ff_fdct_sse2
    w, h, n_sectors = 320, 240, 64
    frames = synth.gen_frames(0, 8, w, h, 3)
    pcm = synth.gen_pcm(2016 * 9, 2, 77)
    ours, used_a, q_a = _mux_strcd(pb.lib(), frames, pcm, n_sectors, w, h)
    theirs, used_b, q_b = _mux_strcd(reference.lib, frames, pcm, n_sectors, w, h)
    assert used_a == used_b == 7 and q_a == q_b
    assert np.array_equal(ours, theirs)
    assert ours[1, 0x18:0x1A].tobytes() == b"\x60\x01"      # STR header where FORMAT_STRCD puts it
