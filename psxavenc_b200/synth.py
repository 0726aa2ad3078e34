"""Deterministic synthetic inputs (integer-only generators of SURVEY.md Appendix B).

The same bytes are fed to the CPU oracle and to the GPU path. NV21 frames are what the
reference's decoder hands to encode_frame_bs (decoding.c:293; mdec.c:593-594): a W*H luma
plane followed by a W*(H/2) plane of interleaved Cr/Cb bytes. PCM is interleaved int16.
"""
import numpy as np

_LCG_A = 1664525
_LCG_C = 1013904223
_M32 = 0xFFFFFFFF


def lcg_sequence(seed, count):
    """First `count` outputs of x = x*1664525 + 1013904223 (mod 2^32), starting after `seed`."""
    if count <= 0:
        return np.zeros(0, dtype=np.uint32)
    seq = np.empty(count, dtype=np.uint64)
    seq[0] = (int(seed) * _LCG_A + _LCG_C) & _M32
    filled, a, c = 1, _LCG_A, _LCG_C  # x[k+filled] = a*x[k] + c
    while filled < count:
        n = min(filled, count - filled)
        seq[filled:filled + n] = (seq[:n] * np.uint64(a) + np.uint64(c)) & np.uint64(_M32)
        filled += n
        c = (c * (a + 1)) & _M32
        a = (a * a) & _M32
    return seq.astype(np.uint32)


def tri(v, period):
    m = np.mod(v, period)
    return np.where(m < period // 2, m, period - m)


def gen_frame(n, width, height, noise_bits=0):
    """One NV21 frame (uint8, 1.5*W*H bytes): moving triangle gradients plus optional noise."""
    x = np.arange(width, dtype=np.int64)[None, :]
    y = np.arange(height, dtype=np.int64)[:, None]
    v = 64 + tri(x + 2 * n, 256) // 2 + tri(2 * y + n, 192) // 2
    if noise_bits:
        seed = (0x1234567 + n * 2654435761) & _M32
        noise = lcg_sequence(seed, width * height).astype(np.int64) >> (32 - noise_bits)
        v = v + noise.reshape(height, width)
    luma = np.minimum(v, 255).astype(np.uint8)

    cx = np.arange(width // 2, dtype=np.int64)[None, :]
    cy = np.arange(height // 2, dtype=np.int64)[:, None]
    chroma = np.empty((height // 2, width), dtype=np.uint8)
    chroma[:, 0::2] = np.broadcast_to(112 + tri(cx + n, 128) // 2, (height // 2, width // 2))
    chroma[:, 1::2] = np.broadcast_to(144 - tri(cy + 3 * n, 96) // 2, (height // 2, width // 2))
    return np.concatenate([luma.ravel(), chroma.ravel()])


def gen_frames(first, count, width, height, noise_bits=0):
    return np.stack([gen_frame(first + i, width, height, noise_bits) for i in range(count)])


def gen_smooth_frame(n, width, height, amplitude=60.0, fx=0.013, fy=0.017):
    """Sinusoidal low-detail frame for tight budgets (sbs: 640x480 into 8192 bytes)."""
    x = np.arange(width)[None, :]
    y = np.arange(height)[:, None]
    luma = 128 + amplitude * np.sin(fx * (x + 3 * n)) * np.cos(fy * (y + 2 * n))
    luma = np.clip(np.rint(luma), 0, 255).astype(np.uint8)
    cx = np.arange(width // 2)[None, :]
    cy = np.arange(height // 2)[:, None]
    chroma = np.empty((height // 2, width), dtype=np.uint8)
    chroma[:, 0::2] = np.clip(np.rint(128 + 30 * np.sin(0.011 * (cx + n)) + 0 * cy), 0, 255)
    chroma[:, 1::2] = np.clip(np.rint(128 + 30 * np.cos(0.009 * (cy + n)) + 0 * cx), 0, 255)
    return np.concatenate([luma.ravel(), chroma.ravel()])


def gen_pcm(count, channels, seed):
    """Interleaved int16 PCM [count, channels]: two triangle partials plus 8-bit noise."""
    i = np.arange(count, dtype=np.int64)[:, None]
    c = np.arange(channels, dtype=np.int64)[None, :]
    noise = lcg_sequence((0xC0FFEE + seed) & _M32, count * channels).astype(np.int64) >> 24
    v = (tri(i * (37 + 11 * c), 4096) - 1024) * 12 + (tri(i * (5 + c), 1024) - 256) * 8
    v = v + noise.reshape(count, channels) - 128
    return v.astype(np.int16)


def gen_sine(count, frequency=440.0, rate=22050.0, amplitude=12000.0):
    """Mono sine used by the `spu` config (BASELINE.json configs[0])."""
    n = np.arange(count, dtype=np.float64)
    return np.rint(amplitude * np.sin(2.0 * np.pi * frequency * n / rate)).astype(np.int16)


def fnv1a64(data):
    """FNV-1a 64 over a bytes-like object."""
    h = 1469598103934665603
    for b in bytes(data):
        h = ((h ^ b) * 1099511628211) & 0xFFFFFFFFFFFFFFFF
    return h
