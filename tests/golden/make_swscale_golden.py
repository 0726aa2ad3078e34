"""Generates tests/golden/swscale_nv21.npz: inputs and libswscale outputs for the colour
conversion / scaling front end (psxb200_nv21_from_device), the way the reference's decoder drives
libswscale (psxavenc/decoding.c:286-311, 463-475): SWS_BICUBIC to AV_PIX_FMT_NV21 with
sws_setColorspaceDetails(.., ITU601 table, dstRange = 1).

Needs the libswscale binary that ships in this image (opencv's bundled FFmpeg 8.0); run as
  LD_LIBRARY_PATH=<site-packages>/opencv_python_headless.libs python tests/golden/make_swscale_golden.py
The GPU tests only read the committed .npz."""
import ctypes as C
import glob
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
LIBS = "/opt/prime-rl/.venv/lib/python3.12/site-packages/opencv_python_headless.libs/"
PIX = {"rgb24": 2, "bgr24": 3, "rgba": 26, "bgra": 28, "yuv420p": 0}   # AVPixelFormat
AV_PIX_FMT_NV21, SWS_BICUBIC, SWS_CS_ITU601 = 24, 4, 5


def load_swscale():
    C.CDLL(glob.glob(LIBS + "libavutil-*")[0], mode=C.RTLD_GLOBAL)
    sws = C.CDLL(glob.glob(LIBS + "libswscale-*")[0])
    sws.sws_getContext.restype = C.c_void_p
    sws.sws_getContext.argtypes = [C.c_int] * 7 + [C.c_void_p] * 3
    sws.sws_getCoefficients.restype = C.POINTER(C.c_int)
    sws.sws_getCoefficients.argtypes = [C.c_int]
    sws.sws_setColorspaceDetails.argtypes = [C.c_void_p, C.POINTER(C.c_int), C.c_int, C.POINTER(C.c_int)] + [C.c_int] * 4
    sws.sws_scale.argtypes = [C.c_void_p, C.POINTER(C.c_void_p), C.POINTER(C.c_int), C.c_int, C.c_int,
                              C.POINTER(C.c_void_p), C.POINTER(C.c_int)]
    return sws


def swscale_nv21(sws, fmt, planes, pitches, sw, sh, dw, dh, src_full_range):
    ctx = sws.sws_getContext(sw, sh, PIX[fmt], dw, dh, AV_PIX_FMT_NV21, SWS_BICUBIC, None, None, None)
    assert ctx
    # decoding.c:299-309: inv_table = coefficients of the stream's colorspace (unspecified -> default)
    sws.sws_setColorspaceDetails(ctx, sws.sws_getCoefficients(2), int(src_full_range), sws.sws_getCoefficients(SWS_CS_ITU601),
                                 1, 0, 1 << 16, 1 << 16)
    out = np.zeros(dw * dh * 3 // 2, np.uint8)
    src = (C.c_void_p * 4)(*[p.ctypes.data for p in planes] + [None] * (4 - len(planes)))
    ss = (C.c_int * 4)(*list(pitches) + [0] * (4 - len(pitches)))
    dst = (C.c_void_p * 4)(out.ctypes.data, out.ctypes.data + dw * dh, None, None)
    ds = (C.c_int * 4)(dw, dw, 0, 0)
    sws.sws_scale(ctx, src, ss, 0, sh, dst, ds)
    return out


def picture(w, h, channels, seed, noise):
    """Smooth colour gradients plus `noise` levels of per-pixel noise."""
    rng = np.random.default_rng(seed)
    x = np.arange(w)[None, :, None]
    y = np.arange(h)[:, None, None]
    c = np.arange(channels)[None, None, :]
    v = 128 + 90 * np.sin((x * (1.0 + 0.3 * c) + seed) / 17.0) * np.cos((y * (1.0 + 0.2 * c) - seed) / 13.0) + 30 * np.sin((x + y) / 5.0 + c)
    v = v + rng.integers(-noise, noise + 1, size=(h, w, channels))
    return np.clip(np.rint(v), 0, 255).astype(np.uint8)


CASES = [   # name, fmt, src w, h, dst w, h, src full range, noise
    ("rgb24_same", "rgb24", 64, 48, 64, 48, 1, 12),
    ("bgr24_same_noisy", "bgr24", 64, 48, 64, 48, 1, 80),
    ("rgba_down2", "rgba", 128, 96, 64, 48, 1, 20),
    ("bgra_down_odd", "bgra", 150, 100, 64, 48, 1, 10),
    ("rgb24_up", "rgb24", 40, 30, 64, 48, 1, 6),
    ("yuv420p_limited_same", "yuv420p", 64, 48, 64, 48, 0, 10),
    ("yuv420p_limited_down", "yuv420p", 160, 120, 80, 64, 0, 10),
    ("yuv420p_full_up", "yuv420p", 48, 32, 64, 48, 1, 4),
]


def main():
    sws = load_swscale()
    data = {}
    for name, fmt, sw, sh, dw, dh, full, noise in CASES:
        if fmt == "yuv420p":
            yuv = picture(sw, sh, 3, len(name), noise)
            lo, hi_y, hi_c = (0, 255, 255) if full else (16, 235, 240)
            yp = np.clip(yuv[..., 0], lo, hi_y).astype(np.uint8)
            up = np.clip(yuv[::2, ::2, 1], lo, hi_c).astype(np.uint8).copy()
            vp = np.clip(yuv[::2, ::2, 2], lo, hi_c).astype(np.uint8).copy()
            out = swscale_nv21(sws, fmt, [yp, up, vp], [sw, sw // 2, sw // 2], sw, sh, dw, dh, full)
            src = np.concatenate([yp.ravel(), up.ravel(), vp.ravel()])
        else:
            ch = 4 if fmt in ("rgba", "bgra") else 3
            img = picture(sw, sh, ch, len(name), noise)
            out = swscale_nv21(sws, fmt, [img], [sw * ch], sw, sh, dw, dh, full)
            src = img.ravel()
        data[name + "/src"] = src
        data[name + "/nv21"] = out
        data[name + "/meta"] = np.array([list(PIX).index(fmt), sw, sh, dw, dh, full], np.int32)
    np.savez_compressed(os.path.join(HERE, "swscale_nv21.npz"), **data)
    print("wrote", os.path.join(HERE, "swscale_nv21.npz"), {k: v.shape for k, v in data.items() if k.endswith("nv21")})


if __name__ == "__main__":
    main()
