"""Differential tests: our C restatement against the unmodified reference build
(oracle/_ref), on random and edge inputs. Skipped where the reference cannot be built."""
import numpy as np
import pytest

import oracle
from psxavenc_b200 import synth


def _blocks(rng, n):
    b = rng.integers(-128, 128, size=(n, 64), dtype=np.int16)
    b[0] = 127
    b[1] = -128
    b[2] = np.tile([127, -128], 32)
    b[3] = np.repeat([127, -128], 32)
    b[4] = (np.arange(64) % 8) * 30 - 105
    b[5] = (np.arange(64) // 8) * 30 - 105
    return b


@pytest.mark.parametrize("variant,mode", [(oracle.FDCT_ISLOW, 1), (oracle.FDCT_SSE2, 0)], ids=["islow", "sse2"])
def test_fdct_models_match_libavcodec(restated, reference, variant, mode):
    """The two FDCT models against the libavcodec 62.11.100 binary."""
    if not reference.has_libavcodec:
        pytest.skip("no libavcodec binary linked")
    blocks = _blocks(np.random.default_rng(7), 200000)
    assert np.array_equal(restated.fdct(variant, blocks), reference.fdct(mode, blocks))


@pytest.mark.parametrize("codec", [0, 1, 2])
@pytest.mark.parametrize("fdct", [oracle.FDCT_ISLOW, oracle.FDCT_SSE2], ids=["islow", "sse2"])
def test_bs_random_frames(restated, reference, codec, fdct):
    rng = np.random.default_rng(100 + codec)
    w, h = 64, 48
    frames = []
    for i in range(12):
        base = synth.gen_frame(i, w, h, noise_bits=i % 7).astype(np.int32)
        amp = 4 + 11 * i   # growing wide-band noise pushes q up and hits escapes/clamps
        base = base + rng.integers(-amp, amp, size=base.shape)
        frames.append(np.clip(base, 0, 255).astype(np.uint8))
    frames.append(np.full(w * h * 3 // 2, 0, np.uint8))
    frames.append(np.full(w * h * 3 // 2, 255, np.uint8))
    chk = np.zeros((h * 3 // 2, w), np.uint8)
    chk[::2, ::2] = 255
    chk[1::2, 1::2] = 255
    frames.append(chk.ravel())
    frames = np.stack(frames)
    sizes = np.array([2016 * (1 + i % 3) for i in range(len(frames))], np.int32)
    o1, r1 = restated.bs_encode_batch(codec, w, h, frames, sizes, fdct, stride=6048)
    o2, r2 = reference.bs_encode_batch(codec, w, h, frames, sizes, fdct, stride=6048)
    assert np.array_equal(r1, r2)
    assert np.array_equal(o1, o2)
    assert len(set(r1[:, 2])) > 3   # several different quant scales exercised


@pytest.mark.parametrize("pitch", [1, 2, 8])
def test_spu_random(restated, reference, pitch):
    rng = np.random.default_rng(pitch)
    for count in (0, 1, 27, 28, 29, 56, 1000):
        pcm = rng.integers(-32768, 32768, size=(max(count, 1) * pitch,), dtype=np.int16)
        s1, s2 = oracle.ChannelState(), oracle.ChannelState()
        a = restated.spu_encode(s1, pcm, count, pitch)
        b = reference.spu_encode(s2, pcm, count, pitch)
        assert np.array_equal(a, b)
        assert (s1.prev1, s1.prev2, s1.mse) == (s2.prev1, s2.prev2, s2.mse)


def test_spu_extremes(restated, reference):
    for pcm in (np.full(280, 32767, np.int16), np.full(280, -32768, np.int16),
                np.tile(np.array([32767, -32768], np.int16), 140), np.zeros(280, np.int16)):
        s1, s2 = oracle.ChannelState(), oracle.ChannelState()
        assert np.array_equal(restated.spu_encode(s1, pcm, 280, 1), reference.spu_encode(s2, pcm, 280, 1))


@pytest.mark.parametrize("stereo", [False, True])
@pytest.mark.parametrize("bits", [4, 8])
@pytest.mark.parametrize("fmt", [0, 1])
def test_xa_random(restated, reference, stereo, bits, fmt):
    rng = np.random.default_rng(bits + fmt)
    ch = 2 if stereo else 1
    for count in (1, 100, 2016, 4032, 5000):
        pcm = np.concatenate([rng.integers(-20000, 20000, size=(count, ch), dtype=np.int16),
                              np.zeros((4100, ch), np.int16)])
        s1, s2 = oracle.new_states(), oracle.new_states()
        a = restated.xa_encode(fmt, stereo, 18900 if bits == 8 else 37800, bits, 3, 5, s1, pcm, count, 1234, finalize=True)
        b = reference.xa_encode(fmt, stereo, 18900 if bits == 8 else 37800, bits, 3, 5, s2, pcm, count, 1234, finalize=True)
        assert np.array_equal(a, b)
        assert bytes(s1) == bytes(s2)


@pytest.mark.parametrize("codec", [0, 1, 2], ids=["v2", "v3", "v3dc"])
@pytest.mark.parametrize("fdct", [oracle.FDCT_ISLOW, oracle.FDCT_SSE2], ids=["islow", "sse2"])
def test_bs_baseline_shapes(restated, reference, codec, fdct):
    """The port against the unmodified reference at the BASELINE shapes themselves: 320x240 with
    the strv and strcd budgets over easy / typical / hard content, and 640x480 with the sbs
    budget — so that GPU tests which compare with the port inherit the reference's bytes."""
    w, h = 320, 240
    frames = np.stack([synth.gen_frame(i, w, h, noise_bits=(0, 3, 5, 6)[i % 4]) for i in range(8)])
    sizes = np.array([20160, 18144, 16128, 20160, 18144, 16128, 20160, 18144], np.int32)
    o1, r1 = restated.bs_encode_batch(codec, w, h, frames, sizes, fdct, stride=20160)
    o2, r2 = reference.bs_encode_batch(codec, w, h, frames, sizes, fdct, stride=20160)
    assert np.array_equal(r1, r2) and len(set(r1[:, 2])) >= 3
    for i, s in enumerate(sizes):     # the reference stores one stray byte AT frame_max_size when a pass overflows
        assert np.array_equal(o1[i, :s], o2[i, :s]), "frame %d" % i
    if codec:
        w, h = 640, 480
        frames = np.stack([synth.gen_smooth_frame(i, w, h, amplitude=20 + 9 * i, fx=0.006 + 0.002 * i, fy=0.009 + 0.001 * i)
                           for i in range(3)])
        o1, r1 = restated.bs_encode_batch(codec, w, h, frames, 8192, fdct)
        o2, r2 = reference.bs_encode_batch(codec, w, h, frames, 8192, fdct)
        assert np.array_equal(r1, r2) and np.array_equal(o1, o2)
