"""Adapter giving the C-ABI GPU library the oracle's Python surface, so the same known-answer
and differential checks run against oracle.Restated, oracle.Reference and the B200 path.
Audio goes through the DROP-IN symbols (psx_audio_spu_encode / psx_audio_xa_encode), video
through the batched host entry point."""
import ctypes as C

import numpy as np

import psxavenc_b200 as pb


class GpuBackend:
    kind = "gpu"

    def __init__(self, max_batch=64):
        self.lib = pb.lib()
        self.max_batch = max_batch
        self._encoders = {}

    def encoder(self, codec, width, height, fdct):
        key = (codec, width, height, fdct)
        if key not in self._encoders:
            self._encoders[key] = pb.BsEncoder(codec, width, height, fdct, self.max_batch)
        return self._encoders[key]

    def bs_encode_batch(self, codec, width, height, frames, max_sizes, fdct=pb.FDCT_ISLOW, stride=None):
        return self.encoder(codec, width, height, fdct).encode_host(frames, max_sizes, stride)

    def spu_encode(self, state, samples, sample_count, pitch, offset=0):
        samples = np.ascontiguousarray(samples, dtype=np.int16).ravel()
        out = np.zeros(16 * ((sample_count + 27) // 28), dtype=np.uint8)
        n = self.lib.psx_audio_spu_encode(C.addressof(state), samples.ctypes.data + 2 * offset, sample_count, pitch,
                                          out.ctypes.data)
        assert n == out.size
        return out

    def xa_encode(self, fmt, stereo, frequency, bits, file_number, channel_number, states, samples, sample_count,
                  lba, out=None, finalize=False):
        cfg = pb.XaSettings(fmt, bool(stereo), frequency, bits, file_number, channel_number)
        samples = np.ascontiguousarray(samples, dtype=np.int16).ravel()
        if out is None:
            out = np.zeros(self.lib.psx_audio_xa_get_buffer_size(cfg, sample_count), dtype=np.uint8)
        n = self.lib.psx_audio_xa_encode(cfg, C.addressof(states), samples.ctypes.data, sample_count, lba,
                                         out.ctypes.data)
        if finalize:
            self.lib.psx_audio_xa_encode_finalize(cfg, out.ctypes.data, n)
        return out[:n]
