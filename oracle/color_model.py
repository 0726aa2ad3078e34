"""numpy restatement of the colour conversion / scaling front end (psxavenc_b200/csrc/color_convert.cu)
— TEST INFRASTRUCTURE ONLY. The reference implementation of this step is libswscale
(psxavenc/decoding.c:286-311, 463-475), a third-party dependency whose source is not in the
reference tree; its published structure (per-pixel RGB->YCbCr, chroma of pixel pairs averaged,
separable cubic B=0 C=0.6 resampling stretched by the scale ratio, aligned sample centres,
replicated edges) is restated here in float64 and pinned against the libswscale 9.1 binary's
outputs in tests/golden/swscale_nv21.npz (tests/test_color_model.py)."""
import numpy as np

PIX_RGB24, PIX_BGR24, PIX_RGBA, PIX_BGRA, PIX_YUV420P = 0, 1, 2, 3, 4


def cubic06(t):
    t = np.abs(t)
    return np.where(t < 1, (1.4 * t - 2.4) * t * t + 1, np.where(t < 2, ((-0.6 * t + 3) * t - 4.8) * t + 2.4, 0.0))


def resample_matrix(src, dst):
    """[dst, src] weights: cubic stretched by max(ratio, 1), centre (i + .5) * ratio - .5, edges replicated."""
    ratio = src / dst
    stretch = max(ratio, 1.0)
    m = np.zeros((dst, src))
    for i in range(dst):
        centre = (i + 0.5) * ratio - 0.5
        first = int(np.ceil(centre - 2 * stretch))
        last = int(np.floor(centre + 2 * stretch))
        ks = np.arange(first, last + 1)
        w = cubic06((ks - centre) / stretch)
        w = w / w.sum()
        for k, wk in zip(np.clip(ks, 0, src - 1), w):
            m[i, k] += wk
    return m


def to_nv21(pixfmt, src, sw, sh, dw, dh, full_range):
    src = np.asarray(src, np.uint8)
    if pixfmt == PIX_YUV420P and (sw, sh) == (dw, dh):
        # libswscale's unscaled planar -> semi-planar path is a plain re-interleave: no range
        # conversion even when the ranges differ (probed on the 9.1 binary)
        n, c = sw * sh, (sw // 2) * (sh // 2)
        chroma = np.empty(2 * c, np.uint8)
        chroma[0::2] = src[n + c:n + 2 * c]
        chroma[1::2] = src[n:n + c]
        return np.concatenate([src[:n], chroma])
    # chroma that went through libswscale's limited -> full range expansion (every RGB source, and
    # limited-range YUV) comes out truncated instead of rounded (probed: a uniform -0.5 bias)
    chroma_bias = 0.0 if (pixfmt == PIX_YUV420P and full_range) else -0.5
    if pixfmt == PIX_YUV420P:
        y = src[:sw * sh].reshape(sh, sw).astype(np.float64)
        u = src[sw * sh:sw * sh + (sw // 2) * (sh // 2)].reshape(sh // 2, sw // 2).astype(np.float64) - 128
        v = src[sw * sh + (sw // 2) * (sh // 2):].reshape(sh // 2, sw // 2).astype(np.float64) - 128
        if not full_range:
            y = (y - 16) * (255.0 / 219.0)
            u = u * (255.0 / 224.0)
            v = v * (255.0 / 224.0)
        cr, cb = v, u
    else:
        bpp = 4 if pixfmt >= PIX_RGBA else 3
        img = src.reshape(sh, sw, bpp).astype(np.float64)
        bgr = pixfmt in (PIX_BGR24, PIX_BGRA)
        r, g, b = img[..., 2 if bgr else 0], img[..., 1], img[..., 0 if bgr else 2]
        y = 0.299 * r + 0.587 * g + 0.114 * b
        if dw // 2 <= sw // 2:
            # chroma of horizontally adjacent pixel pairs is averaged first ...
            cw = (sw + 1) // 2
            idx1 = np.minimum(2 * np.arange(cw) + 1, sw - 1)
            pair = lambda a: 0.5 * (a[:, 0::2][:, :cw] + a[:, idx1])
        else:
            # ... unless that would leave fewer chroma samples than the destination has
            pair = lambda a: a
        rp, gp, bp = pair(r), pair(g), pair(b)
        yp = 0.299 * rp + 0.587 * gp + 0.114 * bp
        cr = (rp - yp) * (0.5 / (1 - 0.299))
        cb = (bp - yp) * (0.5 / (1 - 0.114))
    luma = resample_matrix(sh, dh) @ y @ resample_matrix(sw, dw).T
    mv, mh = resample_matrix(cr.shape[0], dh // 2), resample_matrix(cr.shape[1], dw // 2)
    cr2, cb2 = mv @ cr @ mh.T + 128 + chroma_bias, mv @ cb @ mh.T + 128 + chroma_bias
    q = lambda a: np.clip(np.rint(a), 0, 255).astype(np.uint8)
    chroma = np.empty((dh // 2, dw), np.uint8)
    chroma[:, 0::2] = q(cr2)
    chroma[:, 1::2] = q(cb2)
    return np.concatenate([q(luma).ravel(), chroma.ravel()])
