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
