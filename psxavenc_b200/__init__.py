"""psxavenc_b200 — host-side binding of libpsxav_b200.so, the B200 (sm_100a) MDEC/BS and
SPU/XA-ADPCM encode core.

This module is a thin ctypes mirror of include/psxav_b200.h: the product is the shared
library (CUDA kernels + C ABI); Python is only used by the tests and the benchmark. It
never falls back to a CPU implementation: loading fails loudly when the library has not
been built (python -m psxavenc_b200.build) and every entry point needs a CUDA device.

Reference interface mirrored (WonderfulToolchain/psxavenc): psxavenc/mdec.h:65-74
(init_mdec_encoder / encode_frame_bs / encode_sector_str / destroy_mdec_encoder) and
libpsxav/libpsxav.h:73-101 (psx_audio_spu_encode, psx_audio_xa_encode and helpers).
"""
import ctypes as C
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
# PSXB200_LIB points at an alternative build of the same library (A/B experiments)
LIB_PATH = os.environ.get("PSXB200_LIB") or os.path.join(HERE, "libpsxav_b200.so")

FDCT_ISLOW, FDCT_SSE2 = 0, 1
CODEC_V2, CODEC_V3, CODEC_V3DC = 0, 1, 2
FORMAT_XA, FORMAT_XACD, FORMAT_STR, FORMAT_STRCD, FORMAT_STRV, FORMAT_SBS = 0, 1, 6, 7, 9, 10
PIX_RGB24, PIX_BGR24, PIX_RGBA, PIX_BGRA, PIX_YUV420P = 0, 1, 2, 3, 4


class BsResult(C.Structure):
    """psxb200_bs_result_t"""
    _fields_ = [("bytes_used", C.c_int), ("blocks_used", C.c_int), ("quant_scale", C.c_int),
                ("uncomp_hwords_used", C.c_int)]


class ChannelState(C.Structure):
    """psx_audio_encoder_channel_state_t (libpsxav.h:53-57)"""
    _fields_ = [("qerr", C.c_int), ("mse", C.c_uint64), ("prev1", C.c_int), ("prev2", C.c_int)]


class EncoderState(C.Structure):
    """psx_audio_encoder_state_t (libpsxav.h:59-62)"""
    _fields_ = [("left", ChannelState), ("right", ChannelState)]


class XaSettings(C.Structure):
    """psx_audio_xa_settings_t (libpsxav.h:44-51)"""
    _fields_ = [("format", C.c_int), ("stereo", C.c_bool), ("frequency", C.c_int),
                ("bits_per_sample", C.c_int), ("file_number", C.c_int), ("channel_number", C.c_int)]


class StrParams(C.Structure):
    """psxb200_str_params_t"""
    _fields_ = [("format", C.c_int), ("first_frame_index", C.c_int), ("sectors_num", C.c_int), ("sectors_den", C.c_int),
                ("video_id", C.c_int), ("framing", C.c_int), ("xa_file", C.c_int), ("xa_channel", C.c_int),
                ("interleave", C.c_int), ("trailing_audio", C.c_int), ("place_at_lba", C.c_int),
                ("lba_origin", C.c_longlong), ("frames_per_file", C.c_int), ("file_stride", C.c_longlong)]


def str_params(fmt, sectors_num, sectors_den, first_frame_index=1, video_id=0x8001, framing=0, xa_file=1, xa_channel=0,
               interleave=1, trailing_audio=0, place_at_lba=0, lba_origin=0, frames_per_file=0, file_stride=0):
    return StrParams(fmt, first_frame_index, sectors_num, sectors_den, video_id, framing, xa_file, xa_channel, interleave,
                     trailing_audio, place_at_lba, lba_origin, frames_per_file, file_stride)


class MdecEncoderState(C.Structure):
    """mdec_encoder_state_t (mdec.h:32-55)"""
    _fields_ = [("frame_index", C.c_int), ("frame_data_offset", C.c_int), ("frame_max_size", C.c_int),
                ("frame_block_base_overflow", C.c_int), ("frame_block_overflow_num", C.c_int),
                ("frame_block_overflow_den", C.c_int), ("block_type", C.c_int),
                ("last_dc_values", C.c_int16 * 3), ("bits_value", C.c_uint16), ("bits_left", C.c_int),
                ("frame_output", C.POINTER(C.c_uint8)), ("bytes_used", C.c_int), ("blocks_used", C.c_int),
                ("uncomp_hwords_used", C.c_int), ("quant_scale", C.c_int), ("quant_scale_sum", C.c_int),
                ("dct_context", C.c_void_p), ("ac_huffman_map", C.c_void_p), ("dc_huffman_map", C.c_void_p),
                ("coeff_clamp_map", C.c_void_p), ("dct_block_lists", C.c_void_p * 6)]


class MdecEncoder(C.Structure):
    """mdec_encoder_t (mdec.h:57-63)"""
    _fields_ = [("video_codec", C.c_int), ("video_width", C.c_int), ("video_height", C.c_int),
                ("state", MdecEncoderState)]


# every symbol include/psxav_b200.h declares: name -> (restype, argtypes)
_P = C.c_void_p
SYMBOLS = {
    "init_mdec_encoder": (C.c_bool, [C.POINTER(MdecEncoder), C.c_int, C.c_int, C.c_int]),
    "destroy_mdec_encoder": (None, [C.POINTER(MdecEncoder)]),
    "encode_frame_bs": (None, [C.POINTER(MdecEncoder), _P]),
    "encode_sector_str": (C.c_int, [C.POINTER(MdecEncoder), C.c_int, C.c_uint16, _P, _P]),
    "psx_audio_xa_get_buffer_size": (C.c_uint32, [XaSettings, C.c_int]),
    "psx_audio_spu_get_buffer_size": (C.c_uint32, [C.c_int]),
    "psx_audio_xa_get_buffer_size_per_sector": (C.c_uint32, [XaSettings]),
    "psx_audio_xa_get_samples_per_sector": (C.c_uint32, [XaSettings]),
    "psx_audio_xa_get_sector_interleave": (C.c_uint32, [XaSettings]),
    "psx_audio_xa_encode": (C.c_int, [XaSettings, _P, _P, C.c_int, C.c_int, _P]),
    "psx_audio_xa_encode_simple": (C.c_int, [XaSettings, _P, C.c_int, C.c_int, _P]),
    "psx_audio_spu_encode": (C.c_int, [_P, _P, C.c_int, C.c_int, _P]),
    "psx_audio_spu_encode_simple": (C.c_int, [_P, C.c_int, _P, C.c_int]),
    "psx_audio_xa_encode_finalize": (None, [XaSettings, _P, C.c_int]),
    "psxb200_device_count": (C.c_int, []),
    "psxb200_last_error": (C.c_char_p, []),
    "psxb200_launch_count": (C.c_ulonglong, []),
    "psxb200_bs_create": (_P, [C.c_int, C.c_int, C.c_int, C.c_int, C.c_int]),
    "psxb200_bs_destroy": (None, [_P]),
    "psxb200_bs_device": (C.c_int, [_P]),
    "psxb200_bs_frame_bytes": (C.c_longlong, [_P]),
    "psxb200_pinned_alloc": (_P, [C.c_size_t]),
    "psxb200_pinned_free": (None, [_P]),
    "psxb200_str_slot_range": (C.c_int, [C.POINTER(StrParams), C.c_int, C.POINTER(C.c_longlong), C.POINTER(C.c_longlong)]),
    "psxb200_str_encode_device_ex": (C.c_int, [_P, C.c_int, _P, C.POINTER(StrParams), _P, _P, _P]),
    "psxb200_str_encode_host_ex": (C.c_int, [_P, C.c_int, _P, C.POINTER(StrParams), _P, _P]),
    "psxb200_strcd_image_bytes": (C.c_longlong, [C.POINTER(StrParams), C.c_int, C.c_int, C.c_int, C.c_int]),
    "psxb200_strcd_encode_host": (C.c_int, [_P, C.c_int, C.c_int, _P, C.POINTER(StrParams), C.c_int, C.c_int, C.c_int,
                                            _P, C.c_long, C.c_int, _P, _P, C.c_longlong, _P]),
    "psxb200_bs_multi_create": (_P, [C.c_int] * 6 + [_P]),
    "psxb200_bs_multi_destroy": (None, [_P]),
    "psxb200_bs_multi_device_count": (C.c_int, [_P]),
    "psxb200_bs_multi_encode_host": (C.c_int, [_P, C.c_int, _P, _P, _P, C.c_size_t, _P]),
    "psxb200_bs_multi_str_encode_host": (C.c_int, [_P, C.c_int, _P, C.POINTER(StrParams), _P, _P]),
    "psxb200_bs_multi_strcd_encode_host": (C.c_int, [_P, C.c_int, C.c_int, _P, C.POINTER(StrParams), C.c_int, C.c_int, C.c_int,
                                                     _P, C.c_long, C.c_int, _P, _P, C.c_longlong, _P]),
    "psxb200_nv21_scratch_bytes": (C.c_size_t, [C.c_int] * 5),
    "psxb200_nv21_from_device": (C.c_int, [C.c_int, C.c_int, C.c_int, _P, C.c_size_t] + [C.c_int] * 5 + [_P, _P, _P]),
    "psxb200_bs_lookahead_stats": (None, [_P, C.POINTER(C.c_longlong), C.POINTER(C.c_longlong)]),
    "psxb200_spu_encode_host_multi": (C.c_int, [C.c_int, _P, C.c_int, _P, C.c_int, C.c_long, C.c_int, _P, _P, C.c_long]),
    "psxb200_xa_encode_device_ex": (C.c_int, [C.c_int] * 7 + [_P, C.c_long, C.c_int, C.c_int, C.c_int, _P, _P, C.c_long, C.c_long, _P]),
    "psxb200_xa_encode_host_multi": (C.c_int, [C.c_int, _P] + [C.c_int] * 7 + [_P, C.c_long, C.c_int, C.c_int, _P, _P, C.c_long]),
    "psxb200_bs_timing_enable": (None, [_P, C.c_int]),
    "psxb200_bs_timing_read": (C.c_int, [_P, C.POINTER(C.c_double), C.POINTER(C.c_double), C.POINTER(C.c_int)]),
    "psxb200_bs_encode_device": (C.c_int, [_P, C.c_int, _P, _P, C.c_int, _P, C.c_size_t, _P, _P]),
    "psxb200_bs_encode_host": (C.c_int, [_P, C.c_int, _P, _P, _P, C.c_size_t, _P]),
    "psxb200_str_sector_count": (C.c_longlong, [C.c_int] * 4),
    "psxb200_str_encode_device": (C.c_int, [_P, C.c_int, _P] + [C.c_int] * 5 + [_P, _P, _P]),
    "psxb200_str_encode_host": (C.c_int, [_P, C.c_int, _P] + [C.c_int] * 5 + [_P, _P]),
    "psxb200_spu_encode_device": (C.c_int, [C.c_int, _P, C.c_int, C.c_long, C.c_int, _P, _P, _P, C.c_long, _P]),
    "psxb200_spu_encode_host": (C.c_int, [C.c_int, _P, C.c_int, C.c_long, C.c_int, _P, _P, C.c_long]),
    "psxb200_xa_encode_device": (C.c_int, [C.c_int] * 7 + [_P, C.c_long, C.c_int, C.c_int, _P, _P, C.c_long, _P]),
    "psxb200_xa_encode_host": (C.c_int, [C.c_int] * 7 + [_P, C.c_long, C.c_int, C.c_int, _P, _P, C.c_long]),
}

_lib = None


def lib():
    """The loaded shared library; raises when it has not been built (no fallback)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError("%s is missing: build it with `python -m psxavenc_b200.build` "
                               "(psxavenc_b200 has no CPU fallback)" % LIB_PATH)
        handle = C.CDLL(LIB_PATH)
        for name, (restype, argtypes) in SYMBOLS.items():
            fn = getattr(handle, name)
            fn.restype = restype
            fn.argtypes = argtypes
        _lib = handle
    return _lib


def last_error():
    return lib().psxb200_last_error().decode()


def device_count():
    return lib().psxb200_device_count()


def launch_count():
    return int(lib().psxb200_launch_count())


class Psxb200Error(RuntimeError):
    pass


def _check(rc, what):
    if rc < 0:
        raise Psxb200Error("%s failed: %s" % (what, last_error()))
    return rc


def _ptr(a):
    """Raw address of a numpy array, an int address, or anything with data_ptr() (torch)."""
    if a is None:
        return None
    if isinstance(a, int):
        return a
    if isinstance(a, np.ndarray):
        return a.ctypes.data
    if hasattr(a, "data_ptr"):
        return a.data_ptr()
    return C.addressof(a)


class BsEncoder:
    """Batched MDEC/BS frame encoder (psxb200_bs_*): the GPU form of init_mdec_encoder +
    encode_frame_bs (mdec.c:512-755) for many frames per call."""

    def __init__(self, codec, width, height, fdct=FDCT_ISLOW, max_batch=256):
        self.codec, self.width, self.height, self.fdct = codec, width, height, fdct
        self.frame_bytes = width * height * 3 // 2
        self.handle = lib().psxb200_bs_create(codec, width, height, fdct, max_batch)
        if not self.handle:
            raise Psxb200Error("psxb200_bs_create failed: %s" % last_error())

    def close(self):
        if self.handle:
            lib().psxb200_bs_destroy(self.handle)
            self.handle = None

    __del__ = close

    def timing(self, on):
        """Bracket every internal kernel launch with CUDA events (benchmarks only)."""
        lib().psxb200_bs_timing_enable(self.handle, int(on))

    def read_timing(self):
        """-> (fdct kernel ms, pack kernel ms, launch pairs) since the last read."""
        a, b, n = C.c_double(), C.c_double(), C.c_int()
        _check(lib().psxb200_bs_timing_read(self.handle, C.byref(a), C.byref(b), C.byref(n)), "bs_timing_read")
        return a.value, b.value, n.value

    def encode_device(self, n, d_frames, d_max_sizes, max_size_bound, d_out, out_stride, d_results, stream=None):
        """All pointers are device addresses (ints or torch tensors). Asynchronous on `stream`."""
        return _check(lib().psxb200_bs_encode_device(self.handle, n, _ptr(d_frames), _ptr(d_max_sizes), max_size_bound,
                                                     _ptr(d_out), out_stride, _ptr(d_results), stream), "bs_encode_device")

    def encode_host_into(self, n, h_frames, h_max_sizes, h_out, out_stride, h_results):
        """Host buffers by address (numpy / pinned torch tensors); returns number of failed frames."""
        return _check(lib().psxb200_bs_encode_host(self.handle, n, _ptr(h_frames), _ptr(h_max_sizes), _ptr(h_out),
                                                   out_stride, _ptr(h_results)), "bs_encode_host")

    def str_encode_host(self, frames, first_frame_index, sectors_num, sectors_den, fmt=FORMAT_STRV, video_id=0x8001,
                        sectors=None):
        """psxb200_str_encode_host -> (sectors[n_sectors, sector_size] uint8, res[n, 4] int32)."""
        frames = np.ascontiguousarray(frames, dtype=np.uint8).reshape(-1, self.frame_bytes)
        n = frames.shape[0]
        count = int(lib().psxb200_str_sector_count(n, first_frame_index, sectors_num, sectors_den))
        size = {FORMAT_STRV: 2048, FORMAT_STR: 2336, FORMAT_STRCD: 2352}[fmt]
        if sectors is None:
            sectors = np.zeros((count, size), dtype=np.uint8)
        res = np.zeros((n, 4), dtype=np.int32)
        _check(lib().psxb200_str_encode_host(self.handle, n, frames.ctypes.data, fmt, first_frame_index, sectors_num,
                                             sectors_den, video_id, sectors.ctypes.data, res.ctypes.data), "str_encode_host")
        return sectors, res

    def str_encode_host_ex(self, frames, params, sectors):
        """psxb200_str_encode_host_ex into the caller's `sectors` array -> res[n, 4] int32."""
        frames = np.ascontiguousarray(frames, dtype=np.uint8).reshape(-1, self.frame_bytes)
        res = np.zeros((frames.shape[0], 4), dtype=np.int32)
        _check(lib().psxb200_str_encode_host_ex(self.handle, frames.shape[0], frames.ctypes.data, C.byref(params),
                                                sectors.ctypes.data, res.ctypes.data), "str_encode_host_ex")
        return res

    def strcd_encode_host(self, frames, frames_per_file, params, pcm=None, samples_per_file=0, frequency=37800, bits=4,
                          stereo=True, states=None):
        """psxb200_strcd_encode_host -> (images[n_files, image_bytes] uint8, res[n, 4] int32)."""
        frames = np.ascontiguousarray(frames, dtype=np.uint8).reshape(-1, self.frame_bytes)
        n_files = frames.shape[0] // frames_per_file
        size = int(lib().psxb200_strcd_image_bytes(C.byref(params), frames_per_file, bits, int(stereo),
                                                   samples_per_file if pcm is not None else 0))
        images = np.zeros((n_files, size), dtype=np.uint8)
        res = np.zeros((frames.shape[0], 4), dtype=np.int32)
        pcm_stride = 0
        if pcm is not None:
            pcm = np.ascontiguousarray(pcm, dtype=np.int16).reshape(n_files, -1)
            pcm_stride = pcm.shape[1]
        _check(lib().psxb200_strcd_encode_host(self.handle, n_files, frames_per_file, frames.ctypes.data, C.byref(params),
                                               frequency, bits, int(stereo), _ptr(pcm), pcm_stride, samples_per_file,
                                               None if states is None else C.addressof(states), images.ctypes.data, size,
                                               res.ctypes.data), "strcd_encode_host")
        return images, res

    def encode_host(self, frames, max_sizes, stride=None):
        """frames: uint8 [n, 1.5*W*H]; -> (out[n, stride] uint8, res[n, 4] int32)."""
        frames = np.ascontiguousarray(frames, dtype=np.uint8).reshape(-1, self.frame_bytes)
        n = frames.shape[0]
        max_sizes = np.ascontiguousarray(np.broadcast_to(np.asarray(max_sizes, dtype=np.int32), (n,)))
        stride = int(stride or max_sizes.max())
        out = np.zeros((n, stride), dtype=np.uint8)
        res = np.zeros((n, 4), dtype=np.int32)
        self.encode_host_into(n, frames, max_sizes, out, stride, res)
        return out, res


class BsMultiEncoder:
    """psxb200_bs_multi_*: one process, one encoder + worker thread per device."""

    def __init__(self, codec, width, height, fdct=FDCT_ISLOW, max_batch=256, n_devices=0, device_ids=None):
        self.frame_bytes = width * height * 3 // 2
        ids = None
        if device_ids is not None:
            ids = (C.c_int * len(device_ids))(*device_ids)
            n_devices = len(device_ids)
        self.handle = lib().psxb200_bs_multi_create(codec, width, height, fdct, max_batch, n_devices, ids)
        if not self.handle:
            raise Psxb200Error("psxb200_bs_multi_create failed: %s" % last_error())
        self.n_devices = lib().psxb200_bs_multi_device_count(self.handle)

    def close(self):
        if self.handle:
            lib().psxb200_bs_multi_destroy(self.handle)
            self.handle = None

    __del__ = close

    def encode_host_into(self, n, h_frames, h_max_sizes, h_out, out_stride, h_results):
        return _check(lib().psxb200_bs_multi_encode_host(self.handle, n, _ptr(h_frames), _ptr(h_max_sizes), _ptr(h_out),
                                                         out_stride, _ptr(h_results)), "bs_multi_encode_host")

    def str_encode_host_ex(self, frames, params, sectors):
        frames = np.ascontiguousarray(frames, dtype=np.uint8).reshape(-1, self.frame_bytes)
        res = np.zeros((frames.shape[0], 4), dtype=np.int32)
        _check(lib().psxb200_bs_multi_str_encode_host(self.handle, frames.shape[0], frames.ctypes.data, C.byref(params),
                                                      sectors.ctypes.data, res.ctypes.data), "bs_multi_str_encode_host")
        return res


def spu_encode_host_multi(samples, n_streams, pitch, group_stride, sample_count, n_devices=0, device_ids=None, states=None):
    """psxb200_spu_encode_host_multi -> (out[n_streams, 16*ceil(count/28)] uint8, states)."""
    samples = np.ascontiguousarray(samples, dtype=np.int16)
    if states is None:
        states = (ChannelState * n_streams)()
    row = 16 * ((sample_count + 27) // 28)
    out = np.zeros((n_streams, row), dtype=np.uint8)
    ids = None
    if device_ids is not None:
        ids = (C.c_int * len(device_ids))(*device_ids)
        n_devices = len(device_ids)
    _check(lib().psxb200_spu_encode_host_multi(n_devices, ids, n_streams, samples.ctypes.data, pitch, group_stride,
                                               sample_count, C.addressof(states), out.ctypes.data, row), "spu_encode_host_multi")
    return out, states


def spu_encode_host(samples, n_streams, pitch, group_stride, sample_count, states=None):
    """psxb200_spu_encode_host. samples: int16 array; states: (ChannelState * n_streams) or None (zeroed).
    -> (out[n_streams, 16*ceil(count/28)] uint8, states)."""
    samples = np.ascontiguousarray(samples, dtype=np.int16)
    if states is None:
        states = (ChannelState * n_streams)()
    row = 16 * ((sample_count + 27) // 28)
    out = np.zeros((n_streams, row), dtype=np.uint8)
    _check(lib().psxb200_spu_encode_host(n_streams, samples.ctypes.data, pitch, group_stride, sample_count,
                                         C.addressof(states), out.ctypes.data, row), "spu_encode_host")
    return out, states


def xa_sectors(stereo, bits, sample_count):
    jump = 112 if bits == 8 else 224
    total = sample_count * 2 if stereo else sample_count
    return ((total + jump - 1) // jump + 17) // 18


def xa_encode_host(samples, n_streams, in_stride, sample_count, fmt=FORMAT_XACD, stereo=True, frequency=37800, bits=4,
                   file_number=1, channel_number=0, lba=0, states=None, out=None):
    """psxb200_xa_encode_host -> (out[n_streams, sectors*size] uint8, states)."""
    samples = np.ascontiguousarray(samples, dtype=np.int16)
    if states is None:
        states = (EncoderState * n_streams)()
    row = xa_sectors(stereo, bits, sample_count) * (2336 if fmt == 0 else 2352)
    if out is None:
        out = np.zeros((n_streams, row), dtype=np.uint8)
    n = _check(lib().psxb200_xa_encode_host(n_streams, fmt, int(stereo), frequency, bits, file_number, channel_number,
                                            samples.ctypes.data, in_stride, sample_count, lba, C.addressof(states),
                                            out.ctypes.data, out.shape[1] if out.ndim == 2 else row), "xa_encode_host")
    assert n == row
    return out, states
