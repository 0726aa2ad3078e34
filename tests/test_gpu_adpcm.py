"""GPU parity tests for the SPU/XA-ADPCM path against the CPU oracle, bit-exact."""
import ctypes as C

import numpy as np
import pytest

import oracle
import psxavenc_b200 as pb
from psxavenc_b200 import synth
from tests import kat
from tests.gpu_backend import GpuBackend

pytestmark = pytest.mark.gpu
KAT = kat.load_kat()


@pytest.fixture(scope="module")
def gpu():
    assert pb.device_count() > 0, "GPU tests need a CUDA device"
    return GpuBackend()


@pytest.mark.parametrize("case", KAT["spu"], ids=lambda c: c["name"])
def test_spu_kat(gpu, case):
    pcm = synth.gen_pcm(case["n"], case["ch"], case["seed"])
    states = [pb.ChannelState() for _ in range(case["ch"])]
    out = np.concatenate([gpu.spu_encode(states[c], pcm, case["count"], case["ch"], offset=c)
                          for c in range(case["ch"])])
    assert len(out) == case["len"]
    assert "%016x" % kat.fnv(out) == case["hash"]
    if "prev1" in case:
        assert (states[0].prev1, states[0].prev2) == (case["prev1"], case["prev2"])
        assert out[:16].tobytes().hex() == case["first_block"]


@pytest.mark.parametrize("case", KAT["xa"], ids=lambda c: "ch%d-%dbit-f%d" % (c["ch"], c["bits"], c["format"]))
def test_xa_kat(gpu, case):
    st = pb.EncoderState()
    out = gpu.xa_encode(case["format"], case["ch"] == 2, 37800, case["bits"], 1, 2, st, kat.xa_input(case), case["n"],
                        7, finalize=True)
    assert len(out) == case["len"]
    assert "%016x" % kat.fnv(out) == case["hash"]


@pytest.mark.parametrize("pitch", [1, 2, 8])
def test_spu_dropin_random_and_chunked_state(gpu, restated, pitch):
    """Random full-scale PCM, ragged tails, and state carried across calls the way
    encode_file_spui feeds 3584-sample chunks (filefmt.c:319-341)."""
    rng = np.random.default_rng(pitch)
    for count in (1, 27, 28, 29, 56, 1000):
        pcm = rng.integers(-32768, 32768, size=(count * pitch,), dtype=np.int16)
        s1, s2 = oracle.ChannelState(), pb.ChannelState()
        a = restated.spu_encode(s1, pcm, count, pitch)
        b = gpu.spu_encode(s2, pcm, count, pitch)
        assert np.array_equal(a, b)
        assert (s1.prev1, s1.prev2, s1.mse, s1.qerr) == (s2.prev1, s2.prev2, s2.mse, s2.qerr)
    pcm = synth.gen_pcm(3584 * 3 + 100, pitch, 9)
    s1, s2 = oracle.ChannelState(), pb.ChannelState()
    for first in range(0, len(pcm), 3584):
        count = min(3584, len(pcm) - first)
        a = restated.spu_encode(s1, pcm[first:], count, pitch, offset=pitch - 1)
        b = gpu.spu_encode(s2, pcm[first:], count, pitch, offset=pitch - 1)
        assert np.array_equal(a, b)
        assert (s1.prev1, s1.prev2, s1.mse) == (s2.prev1, s2.prev2, s2.mse)


def test_spu_extremes(gpu, restated):
    for pcm in (np.full(280, 32767, np.int16), np.full(280, -32768, np.int16),
                np.tile(np.array([32767, -32768], np.int16), 140), np.zeros(280, np.int16),
                np.tile(np.array([32767, 32767, -32768, -32768], np.int16), 70)):
        s1, s2 = oracle.ChannelState(), pb.ChannelState()
        assert np.array_equal(restated.spu_encode(s1, pcm, 280, 1), gpu.spu_encode(s2, pcm, 280, 1))
        assert (s1.prev1, s1.prev2, s1.mse) == (s2.prev1, s2.prev2, s2.mse)


def test_spu_batched_streams(gpu, restated):
    """psxb200_spu_encode_host: B interleaved 8-channel files = 8*B independent chains
    (the `vagi x B` workload), each equal to the oracle run on that channel alone."""
    files, ch, count = 5, 8, 3584 + 28 * 3 + 5
    pcm = np.stack([synth.gen_pcm(count, ch, 20 + f) for f in range(files)])   # [files, count, ch]
    out, states = pb.spu_encode_host(pcm, files * ch, ch, count * ch, count)
    for f in range(files):
        for c in range(ch):
            st = oracle.ChannelState()
            exp = restated.spu_encode(st, pcm[f], count, ch, offset=c)
            s = f * ch + c
            assert np.array_equal(out[s], exp), (f, c)
            assert (states[s].prev1, states[s].prev2, states[s].mse) == (st.prev1, st.prev2, st.mse)


def test_spu_device_api_with_ragged_counts(gpu, restated):
    torch = pytest.importorskip("torch")
    n, cap = 37, 28 * 40
    rng = np.random.default_rng(2)
    pcm = rng.integers(-30000, 30000, size=(n, cap), dtype=np.int16)
    counts = rng.integers(0, cap + 1, size=n).astype(np.int32)
    counts[0], counts[1], counts[2] = 0, cap, 1
    d_pcm = torch.from_numpy(pcm).cuda()
    d_counts = torch.from_numpy(counts).cuda()
    d_states = torch.zeros((n, 24), dtype=torch.uint8, device="cuda")
    d_out = torch.zeros((n, 16 * 40), dtype=torch.uint8, device="cuda")
    rc = pb.lib().psxb200_spu_encode_device(n, d_pcm.data_ptr(), 1, cap, cap, d_counts.data_ptr(), d_states.data_ptr(),
                                            d_out.data_ptr(), 16 * 40, None)
    assert rc == 0
    torch.cuda.synchronize()
    out = d_out.cpu().numpy()
    states = d_states.cpu().numpy()
    for s in range(n):
        st = oracle.ChannelState()
        exp = restated.spu_encode(st, pcm[s], int(counts[s]), 1)
        assert np.array_equal(out[s, :len(exp)], exp), s
        assert not out[s, len(exp):].any()
        assert bytes(states[s]) == bytes(st), s


@pytest.mark.parametrize("stereo", [False, True], ids=["mono", "stereo"])
@pytest.mark.parametrize("bits", [4, 8])
@pytest.mark.parametrize("fmt", [0, 1], ids=["xa", "xacd"])
def test_xa_dropin_random(gpu, restated, stereo, bits, fmt):
    """Ragged tails (incl. the stereo limit quirk of adpcm.c:204-211), multi-sector calls and
    state carried sector by sector as encode_file_xa does (filefmt.c:175-197)."""
    rng = np.random.default_rng(bits + fmt)
    ch = 2 if stereo else 1
    freq = 18900 if bits == 8 else 37800
    for count in (1, 13, 100, 2016, 4032, 5000):
        pcm = np.concatenate([rng.integers(-20000, 20000, size=(count, ch), dtype=np.int16),
                              rng.integers(-20000, 20000, size=(300, ch), dtype=np.int16)])   # live data past the end
        s1, s2 = oracle.new_states(), pb.EncoderState()
        a = restated.xa_encode(fmt, stereo, freq, bits, 3, 5, s1, pcm, count, 1234, finalize=True)
        b = gpu.xa_encode(fmt, stereo, freq, bits, 3, 5, s2, pcm, count, 1234, finalize=True)
        assert np.array_equal(a, b), count
        assert bytes(s1) == bytes(s2)
    per_sector = ((112 if bits == 8 else 224) >> (1 if stereo else 0)) * 18
    pcm = np.concatenate([synth.gen_pcm(per_sector * 3 + 77, ch, 31), np.zeros((per_sector, ch), np.int16)])
    s1, s2 = oracle.new_states(), pb.EncoderState()
    for k, first in enumerate(range(0, per_sector * 3 + 77, per_sector)):
        count = min(per_sector, per_sector * 3 + 77 - first)
        a = restated.xa_encode(fmt, stereo, freq, bits, 1, 0, s1, pcm[first:], count, k)
        b = gpu.xa_encode(fmt, stereo, freq, bits, 1, 0, s2, pcm[first:], count, k)
        assert np.array_equal(a, b), k
        assert bytes(s1) == bytes(s2)


def test_xa_batched_streams(gpu, restated):
    """psxb200_xa_encode_host over independent stereo streams (strcd audio of several files)."""
    n, count = 6, 2016 * 2 + 500
    pcm = np.stack([np.concatenate([synth.gen_pcm(count, 2, 40 + s), np.zeros((224, 2), np.int16)]) for s in range(n)])
    out, states = pb.xa_encode_host(pcm, n, pcm.shape[1] * 2, count, fmt=pb.FORMAT_XACD, stereo=True, bits=4,
                                    file_number=1, channel_number=0, lba=100)
    for s in range(n):
        st = oracle.new_states()
        exp = restated.xa_encode(1, True, 37800, 4, 1, 0, st, pcm[s], count, 100)
        assert np.array_equal(out[s], exp), s
        assert bytes(states[s]) == bytes(st)


@pytest.mark.parametrize("bits", [4, 8])
@pytest.mark.parametrize("n", [1, 2, 5])
def test_xa_batched_mono_streams(gpu, restated, bits, n):
    """Mono streams share a warp in pairs (one per half-warp): even, odd and single counts."""
    count = 4032 + 777
    pad = 224
    pcm = np.stack([np.concatenate([synth.gen_pcm(count, 1, 70 + s).ravel(), np.zeros(pad, np.int16)]) for s in range(n)])
    out, states = pb.xa_encode_host(pcm, n, pcm.shape[1], count, fmt=pb.FORMAT_XACD, stereo=False, bits=bits,
                                    file_number=2, channel_number=3, lba=7)
    for s in range(n):
        st = oracle.new_states()
        exp = restated.xa_encode(1, False, 37800, bits, 2, 3, st, pcm[s], count, 7)
        assert np.array_equal(out[s], exp), s
        assert bytes(states[s]) == bytes(st)
