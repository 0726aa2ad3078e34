"""CPU-side checks of the drop-in boundary: the C-ABI library loads, exports every symbol
include/psxav_b200.h declares (and the binding table covers them all), struct layouts match
the reference's, and the host-only helpers compute the reference's values. No kernel runs."""
import ctypes as C
import os
import re

import pytest

import oracle
import psxavenc_b200 as pb
from psxavenc_b200 import build as pb_build

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    pb_build.build()
    return pb.lib()


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "psxav_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    names = re.findall(r"\b(psxb200_\w+|psx_audio_\w+|init_mdec_encoder|destroy_mdec_encoder|encode_frame_bs|encode_sector_str)\s*\(", text)
    return sorted(set(names))


def test_header_symbols_exported(lib):
    names = declared_symbols()
    assert len(names) >= 25
    for name in names:
        assert hasattr(lib, name), "library does not export %s" % name
    assert sorted(pb.SYMBOLS) == names, "binding table and header disagree"


def test_struct_layouts_match_reference():
    # mdec.h:32-63 on LP64 (cross-checked against the compiled reference in the next test)
    assert pb.MdecEncoderState.frame_output.offset == 40
    assert pb.MdecEncoderState.bytes_used.offset == 48
    assert pb.MdecEncoderState.quant_scale_sum.offset == 64
    assert pb.MdecEncoderState.dct_context.offset == 72
    assert C.sizeof(pb.MdecEncoderState) == 152
    assert pb.MdecEncoder.state.offset == 16 and C.sizeof(pb.MdecEncoder) == 168
    assert C.sizeof(pb.ChannelState) == 24 == C.sizeof(oracle.ChannelState)
    assert pb.ChannelState.mse.offset == 8 and pb.ChannelState.prev1.offset == 16
    assert C.sizeof(pb.EncoderState) == 48
    assert C.sizeof(pb.XaSettings) == 24 == C.sizeof(oracle.XaSettingsRef)
    assert C.sizeof(pb.BsResult) == 16


def test_reference_struct_sizes(reference):
    """The same numbers straight from the compiled reference headers."""
    fn = getattr(reference.lib, "ref_sizeof_mdec_encoder", None)
    if fn is None:
        pytest.skip("reference driver without size probe")
    fn.restype = C.c_size_t
    assert fn() == C.sizeof(pb.MdecEncoder)


@pytest.mark.parametrize("fmt", [0, 1])
@pytest.mark.parametrize("stereo", [False, True])
@pytest.mark.parametrize("bits", [4, 8])
@pytest.mark.parametrize("freq", [18900, 37800])
def test_size_helpers_match_reference(lib, reference, fmt, stereo, bits, freq):
    ours = pb.XaSettings(fmt, stereo, freq, bits, 1, 2)
    theirs = oracle.XaSettingsRef(fmt, stereo, freq, bits, 1, 2)
    for name in ("psx_audio_xa_get_buffer_size_per_sector", "psx_audio_xa_get_samples_per_sector",
                 "psx_audio_xa_get_sector_interleave"):
        ref_fn = getattr(reference.lib, name)
        ref_fn.restype = C.c_uint32
        ref_fn.argtypes = [oracle.XaSettingsRef]
        assert getattr(lib, name)(ours) == ref_fn(theirs), name
    ref_fn = reference.lib.psx_audio_xa_get_buffer_size
    ref_fn.restype = C.c_uint32
    ref_fn.argtypes = [oracle.XaSettingsRef, C.c_int]
    for count in (0, 1, 2015, 2016, 2017, 4032, 100000):
        assert lib.psx_audio_xa_get_buffer_size(ours, count) == ref_fn(theirs, count)
    reference.lib.psx_audio_spu_get_buffer_size.restype = C.c_uint32
    for count in (0, 1, 27, 28, 29, 3584, 1323000):
        assert lib.psx_audio_spu_get_buffer_size(count) == reference.lib.psx_audio_spu_get_buffer_size(count)


def test_xa_finalize_matches_reference(lib, reference):
    import numpy as np
    rng = np.random.default_rng(3)
    for fmt, length in ((0, 2336), (0, 4672), (1, 2352), (1, 7056), (1, 100)):
        buf = rng.integers(0, 256, size=8000, dtype=np.uint8)
        a, b = buf.copy(), buf.copy()
        lib.psx_audio_xa_encode_finalize(pb.XaSettings(fmt, True, 37800, 4, 0, 0), a.ctypes.data + 16, length)
        reference.lib.psx_audio_xa_encode_finalize(oracle.XaSettingsRef(fmt, True, 37800, 4, 0, 0), b.ctypes.data + 16, length)
        assert np.array_equal(a, b)


def test_no_device_fails_loudly(lib):
    """Without a GPU the product refuses to run; it never computes on the CPU."""
    if lib.psxb200_device_count() > 0:
        pytest.skip("a CUDA device is present")
    with pytest.raises(pb.Psxb200Error, match="no CUDA device"):
        pb.BsEncoder(pb.CODEC_V2, 320, 240)
    with pytest.raises(pb.Psxb200Error):
        import numpy as np
        pb.spu_encode_host(np.zeros(28, np.int16), 1, 1, 28, 28)


def test_bad_arguments_rejected(lib):
    for args in ((0, 100, 240), (0, 320, 100), (3, 320, 240), (-1, 320, 240), (0, 0, 0)):
        with pytest.raises(pb.Psxb200Error):
            pb.BsEncoder(*args)


def test_product_never_imports_oracle():
    """Boundary hygiene: nothing under psxavenc_b200/ or include/ references oracle/."""
    bad = []
    for base in ("psxavenc_b200", "include"):
        for dirpath, _, files in os.walk(os.path.join(ROOT, base)):
            for f in files:
                if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp")):
                    text = open(os.path.join(dirpath, f)).read()
                    if re.search(r"\bimport oracle\b|from oracle\b|oracle/|psx_oracle|orc_", text):
                        bad.append(os.path.join(dirpath, f))
    assert not bad, bad
