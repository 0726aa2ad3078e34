"""ctypes front-end of the TEST ORACLE (see oracle/psx_oracle.h).

TEST INFRASTRUCTURE ONLY. Importable from tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs; the product package psxavenc_b200 never imports it.

Two back-ends with the same Python surface:
  Restated()  oracle/liboracle.so        — our C restatement of the reference algorithm
  Reference() oracle/_ref/libpsxav_ref.so — the unmodified reference C sources, compiled
              from /root/reference by oracle/Makefile (prebuilt file travels to the GPU box)
"""
import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
RESTATED_SO = os.path.join(HERE, "liboracle.so")
REFERENCE_SO = os.path.join(HERE, "_ref", "libpsxav_ref.so")
DROPIN_DRIVER = os.path.join(HERE, "_ref", "dropin_driver")

FDCT_ISLOW, FDCT_SSE2 = 0, 1
CODEC_V2, CODEC_V3, CODEC_V3DC = 0, 1, 2


def build(force=False):
    """Compile liboracle.so and (when /root/reference is present) _ref/libpsxav_ref.so."""
    args = ["make", "-C", HERE] + (["-B"] if force else []) + ["all", "dropin"]
    subprocess.run(args, check=True, stdout=subprocess.DEVNULL)


class ChannelState(C.Structure):
    """psx_audio_encoder_channel_state_t (libpsxav.h:53-57)."""
    _fields_ = [("qerr", C.c_int), ("mse", C.c_uint64), ("prev1", C.c_int), ("prev2", C.c_int)]


class XaSettingsRef(C.Structure):
    """psx_audio_xa_settings_t (libpsxav.h:44-51), passed BY VALUE to the reference."""
    _fields_ = [("format", C.c_int), ("stereo", C.c_bool), ("frequency", C.c_int),
                ("bits_per_sample", C.c_int), ("file_number", C.c_int), ("channel_number", C.c_int)]


class XaSettingsOrc(C.Structure):
    _fields_ = [("format", C.c_int), ("stereo", C.c_int), ("frequency", C.c_int),
                ("bits_per_sample", C.c_int), ("file_number", C.c_int), ("channel_number", C.c_int)]


def _u8(a):
    return a.ctypes.data_as(C.POINTER(C.c_uint8))


def _i16(a):
    return a.ctypes.data_as(C.POINTER(C.c_int16))


def _i32(a):
    return a.ctypes.data_as(C.POINTER(C.c_int))


class _Base:
    def xa_sector_size(self, fmt):
        return 2336 if fmt == 0 else 2352

    def xa_buffer_size(self, fmt, stereo, bits, sample_count):
        per = ((112 if bits == 8 else 224) >> (1 if stereo else 0)) * 18
        return ((sample_count + per - 1) // per) * self.xa_sector_size(fmt)


class Restated(_Base):
    """Our CPU restatement (kind: "port")."""
    kind = "port"

    def __init__(self):
        if not os.path.exists(RESTATED_SO):
            build()
        self.lib = C.CDLL(RESTATED_SO)
        self.lib.orc_fnv1a64.restype = C.c_uint64
        self.lib.orc_fnv1a64.argtypes = [C.c_void_p, C.c_long]
        self.lib.orc_edc_crc32.restype = C.c_uint32
        self.lib.orc_edc_crc32.argtypes = [C.c_void_p, C.c_int]
        self.lib.orc_bs_encode_batch.argtypes = [C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p,
                                                 C.c_void_p, C.c_void_p, C.c_long, C.c_void_p]
        self.lib.orc_spu_encode.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p]
        self.lib.orc_xa_encode.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p]
        self.lib.orc_xa_finalize.argtypes = [C.c_void_p, C.c_void_p, C.c_int]
        self.lib.orc_fdct_batch.argtypes = [C.c_int, C.c_void_p, C.c_int]

    def fdct(self, variant, blocks):
        out = np.ascontiguousarray(blocks, dtype=np.int16).copy()
        self.lib.orc_fdct_batch(variant, out.ctypes.data, out.size // 64)
        return out

    def fnv(self, data):
        a = np.ascontiguousarray(np.frombuffer(bytes(data), dtype=np.uint8)) if not isinstance(data, np.ndarray) else np.ascontiguousarray(data)
        return int(self.lib.orc_fnv1a64(a.ctypes.data, a.nbytes))

    def edc(self, data):
        a = np.ascontiguousarray(data, dtype=np.uint8)
        return int(self.lib.orc_edc_crc32(a.ctypes.data, a.nbytes))

    def bs_encode_batch(self, codec, width, height, frames, max_sizes, fdct=FDCT_ISLOW, stride=None):
        """-> (out[n, stride] uint8, res[n, 4] int32 = bytes_used, blocks_used, q, uncomp_hwords)."""
        frames = np.ascontiguousarray(frames, dtype=np.uint8).reshape(-1, width * height * 3 // 2)
        n = frames.shape[0]
        max_sizes = np.ascontiguousarray(np.broadcast_to(np.asarray(max_sizes, dtype=np.int32), (n,)))
        stride = int(stride or max_sizes.max())
        out = np.zeros((n, stride), dtype=np.uint8)
        res = np.zeros((n, 4), dtype=np.int32)
        self.lib.orc_bs_encode_batch(codec, fdct, width, height, n, frames.ctypes.data, max_sizes.ctypes.data,
                                     out.ctypes.data, stride, res.ctypes.data)
        return out, res

    def spu_encode(self, state, samples, sample_count, pitch, offset=0):
        """samples: int16 array; encodes samples[offset::pitch][:sample_count]. state: ChannelState."""
        samples = np.ascontiguousarray(samples, dtype=np.int16).ravel()
        out = np.zeros(16 * ((sample_count + 27) // 28), dtype=np.uint8)
        n = self.lib.orc_spu_encode(C.byref(state), samples.ctypes.data + 2 * offset, sample_count, pitch,
                                    out.ctypes.data)
        assert n == out.size
        return out

    def xa_encode(self, fmt, stereo, frequency, bits, file_number, channel_number, states, samples,
                  sample_count, lba, out=None, finalize=False):
        """states: (ChannelState * 2). samples must be padded to whole sound groups."""
        cfg = XaSettingsOrc(fmt, int(stereo), frequency, bits, file_number, channel_number)
        samples = np.ascontiguousarray(samples, dtype=np.int16).ravel()
        if out is None:
            out = np.zeros(self.xa_buffer_size(fmt, stereo, bits, sample_count), dtype=np.uint8)
        n = self.lib.orc_xa_encode(C.byref(cfg), C.byref(states), samples.ctypes.data, sample_count, lba,
                                   out.ctypes.data)
        if finalize:
            self.lib.orc_xa_finalize(C.byref(cfg), out.ctypes.data, n)
        return out[:n]


class Reference(_Base):
    """The unmodified reference sources (kind: "reference")."""
    kind = "reference"
    # fdct modes of ref_driver.c
    LAVC_DEFAULT, LAVC_ISLOW, MODEL_ISLOW, MODEL_SSE2 = 0, 1, 2, 3

    def __init__(self):
        if not os.path.exists(REFERENCE_SO):
            build()
        if not os.path.exists(REFERENCE_SO):
            raise FileNotFoundError(REFERENCE_SO + " (needs /root/reference to build)")
        self.lib = C.CDLL(REFERENCE_SO)
        self.lib.ref_bs_open.restype = C.c_void_p
        self.lib.ref_bs_open.argtypes = [C.c_int] * 4
        self.lib.ref_bs_close.argtypes = [C.c_void_p]
        self.lib.ref_bs_encode_batch.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p,
                                                 C.c_long, C.c_void_p]
        self.lib.ref_fdct_blocks.argtypes = [C.c_int, C.c_void_p, C.c_int]
        self.lib.ref_str_encode.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_int,
                                            C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_void_p]
        self.lib.psx_audio_spu_encode.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p]
        self.lib.psx_audio_xa_encode.argtypes = [XaSettingsRef, C.c_void_p, C.c_void_p, C.c_int, C.c_int,
                                                 C.c_void_p]
        self.lib.psx_audio_xa_encode_finalize.argtypes = [XaSettingsRef, C.c_void_p, C.c_int]
        self.has_libavcodec = bool(self.lib.ref_has_libavcodec())

    def fdct_mode(self, fdct):
        """Map FDCT_ISLOW/FDCT_SSE2 to the libavcodec-backed mode when the binary is linked."""
        if self.has_libavcodec:
            return self.LAVC_ISLOW if fdct == FDCT_ISLOW else self.LAVC_DEFAULT
        return self.MODEL_ISLOW if fdct == FDCT_ISLOW else self.MODEL_SSE2

    def fdct(self, mode, blocks):
        out = np.ascontiguousarray(blocks, dtype=np.int16).copy()
        self.lib.ref_fdct_blocks(mode, out.ctypes.data, out.size // 64)
        return out

    def bs_encode_batch(self, codec, width, height, frames, max_sizes, fdct=FDCT_ISLOW, stride=None, mode=None):
        frames = np.ascontiguousarray(frames, dtype=np.uint8).reshape(-1, width * height * 3 // 2)
        n = frames.shape[0]
        max_sizes = np.ascontiguousarray(np.broadcast_to(np.asarray(max_sizes, dtype=np.int32), (n,)))
        stride = int(stride or max_sizes.max())
        out = np.zeros((n, stride), dtype=np.uint8)
        res = np.zeros((n, 4), dtype=np.int32)
        h = self.lib.ref_bs_open(codec, width, height, self.fdct_mode(fdct) if mode is None else mode)
        assert h
        try:
            self.lib.ref_bs_encode_batch(h, n, frames.ctypes.data, max_sizes.ctypes.data, out.ctypes.data, stride,
                                         res.ctypes.data)
        finally:
            self.lib.ref_bs_close(h)
        return out, res

    def str_encode(self, codec, width, height, frames, n_sectors, overflow_base, overflow_den, fmt=9,
                   video_id=0x8001, sector_size=2048, fdct=FDCT_ISLOW, max_frame_size=2016 * 16):
        """encode_sector_str over a video-only stream (mdec.c:757-836; format 9 = FORMAT_STRV)."""
        frames = np.ascontiguousarray(frames, dtype=np.uint8)
        out = np.zeros((n_sectors, sector_size), dtype=np.uint8)
        frame_buf = np.zeros(max_frame_size, dtype=np.uint8)
        used_per = np.zeros(n_sectors, dtype=np.int32)
        h = self.lib.ref_bs_open(codec, width, height, self.fdct_mode(fdct))
        try:
            used = self.lib.ref_str_encode(h, fmt, video_id, frames.ctypes.data, n_sectors, overflow_base,
                                           overflow_den, frame_buf.ctypes.data, frame_buf.size, out.ctypes.data,
                                           sector_size, used_per.ctypes.data)
        finally:
            self.lib.ref_bs_close(h)
        return out, used, used_per

    def str_mux(self, codec, width, height, frames, fmt=7, pcm=None, n_samples=0, frequency=37800, bits=4, stereo=True,
                cd_speed=2, fps_num=15, fps_den=1, trailing_audio=False, video_id=0x8001, xa_file=1, xa_channel=0,
                fdct=FDCT_ISLOW, max_sectors=None):
        """The sector loop of encode_file_str (filefmt.c:391-520; format 6 = STR, 7 = STRCD) over
        in-memory frames and PCM -> (image[n_sectors, sector_size] uint8, quant_scale_sum)."""
        frames = np.ascontiguousarray(frames, dtype=np.uint8).reshape(-1, width * height * 3 // 2)
        n_frames = frames.shape[0]
        size = 2352 if fmt == 7 else 2336
        if max_sectors is None:
            max_sectors = 64 + 16 * n_frames * max(1, (75 * cd_speed * fps_den + fps_num - 1) // fps_num)
        out = np.zeros((max_sectors, size), dtype=np.uint8)
        frame_buf = np.zeros(2016 * 64, dtype=np.uint8)
        qsum = C.c_int(0)
        if pcm is not None:
            pcm = np.ascontiguousarray(pcm, dtype=np.int16).ravel()
            # psx_audio_xa_encode may read a little past the last sample frame (adpcm.c:193-233)
            pcm = np.concatenate([pcm, np.zeros(512, np.int16)])
        self.lib.ref_str_mux.argtypes = [C.c_void_p] + [C.c_int] * 4 + [C.c_void_p, C.c_int, C.c_void_p] + [C.c_int] * 8 + \
                                        [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p]
        h = self.lib.ref_bs_open(codec, width, height, self.fdct_mode(fdct))
        try:
            n = self.lib.ref_str_mux(h, fmt, video_id, xa_file, xa_channel, frames.ctypes.data, n_frames,
                                     None if pcm is None else pcm.ctypes.data, n_samples if pcm is not None else 0, frequency,
                                     bits, int(stereo), cd_speed, fps_num, fps_den, int(trailing_audio), frame_buf.ctypes.data,
                                     out.ctypes.data, max_sectors, C.byref(qsum))
        finally:
            self.lib.ref_bs_close(h)
        assert n >= 0, "ref_str_mux ran out of room"
        return out[:n], qsum.value

    def spu_encode(self, state, samples, sample_count, pitch, offset=0):
        samples = np.ascontiguousarray(samples, dtype=np.int16).ravel()
        out = np.zeros(16 * ((sample_count + 27) // 28), dtype=np.uint8)
        n = self.lib.psx_audio_spu_encode(C.byref(state), samples.ctypes.data + 2 * offset, sample_count, pitch,
                                          out.ctypes.data)
        assert n == out.size
        return out

    def xa_encode(self, fmt, stereo, frequency, bits, file_number, channel_number, states, samples,
                  sample_count, lba, out=None, finalize=False):
        cfg = XaSettingsRef(fmt, bool(stereo), frequency, bits, file_number, channel_number)
        samples = np.ascontiguousarray(samples, dtype=np.int16).ravel()
        if out is None:
            out = np.zeros(self.xa_buffer_size(fmt, stereo, bits, sample_count), dtype=np.uint8)
        # (re)state the prototype: tests that drive this library through other struct classes reset it
        self.lib.psx_audio_xa_encode.argtypes = [XaSettingsRef, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p]
        self.lib.psx_audio_xa_encode_finalize.argtypes = [XaSettingsRef, C.c_void_p, C.c_int]
        n = self.lib.psx_audio_xa_encode(cfg, C.byref(states), samples.ctypes.data, sample_count, lba,
                                         out.ctypes.data)
        if finalize:
            self.lib.psx_audio_xa_encode_finalize(cfg, out.ctypes.data, n)
        return out[:n]


def new_states():
    """Zeroed psx_audio_encoder_state_t (left, right) as the callers create it (filefmt.c:172-173)."""
    return (ChannelState * 2)()
