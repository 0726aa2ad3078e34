"""GPU: psxb200_nv21_from_device (the libswscale step of psxavenc/decoding.c:
286-311, 463-475) against the libswscale outputs in tests/golden/swscale_nv21.npz — every sample
within +-1 (float32 here, 15-bit fixed point there) — and against the numpy restatement."""
import os

import numpy as np
import pytest

import psxavenc_b200 as pb
from oracle import color_model

pytestmark = pytest.mark.gpu
GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "swscale_nv21.npz")
DATA = np.load(GOLDEN)
NAMES = sorted({k.split("/")[0] for k in DATA.files})


def convert(fmt, src, sw, sh, dw, dh, full, n=1):
    torch = pytest.importorskip("torch")
    bpp = {pb.PIX_YUV420P: 1}.get(fmt, 4 if fmt >= pb.PIX_RGBA else 3)
    d_src = torch.from_numpy(np.tile(src, n)).cuda()
    d_out = torch.zeros(n * dw * dh * 3 // 2, dtype=torch.uint8, device="cuda")
    scratch = torch.zeros(pb.lib().psxb200_nv21_scratch_bytes(fmt, n, sw, sh, dw) // 4 + 2, dtype=torch.float32, device="cuda")
    rc = pb.lib().psxb200_nv21_from_device(fmt, full, n, d_src.data_ptr(), len(src), sw, sh, sw * bpp, dw, dh, d_out.data_ptr(),
                                           scratch.data_ptr(), None)
    assert rc == 0, pb.last_error()
    torch.cuda.synchronize()
    return d_out.cpu().numpy().reshape(n, -1)


@pytest.mark.parametrize("name", NAMES)
def test_nv21_from_device_matches_libswscale(name):
    fmt, sw, sh, dw, dh, full = (int(v) for v in DATA[name + "/meta"])
    got = convert(fmt, DATA[name + "/src"], sw, sh, dw, dh, full, n=3)
    assert (got == got[0]).all()                      # every frame of the batch
    exp = DATA[name + "/nv21"].astype(np.int32)
    diff = np.abs(got[0].astype(np.int32) - exp)
    assert diff.max() <= 1, "max |gpu - libswscale| = %d" % diff.max()
    assert (diff == 0).mean() >= 0.97
    model = color_model.to_nv21(fmt, DATA[name + "/src"], sw, sh, dw, dh, full).astype(np.int32)
    dm = np.abs(got[0].astype(np.int32) - model)
    assert dm.max() <= 1 and (dm == 0).mean() >= 0.995      # float32 vs float64 rounding ties only


def test_nv21_feeds_the_encoder():
    """RGB picture -> NV21 on the device -> BS frame, without leaving HBM."""
    torch = pytest.importorskip("torch")
    import oracle
    w, h = 320, 240
    rng = np.random.default_rng(4)
    x, y = np.meshgrid(np.arange(2 * w), np.arange(2 * h))
    rgb = np.stack([128 + 100 * np.sin(x / 40.0), 128 + 90 * np.cos(y / 33.0), 128 + 60 * np.sin((x + y) / 57.0)], -1)
    rgb = np.clip(rgb + rng.integers(-4, 5, size=rgb.shape), 0, 255).astype(np.uint8)
    d_rgb = torch.from_numpy(rgb.ravel()).cuda()
    d_nv21 = torch.zeros(w * h * 3 // 2, dtype=torch.uint8, device="cuda")
    scratch = torch.zeros(pb.lib().psxb200_nv21_scratch_bytes(pb.PIX_RGB24, 1, 2 * w, 2 * h, w) // 4 + 2, dtype=torch.float32, device="cuda")
    assert pb.lib().psxb200_nv21_from_device(pb.PIX_RGB24, 1, 1, d_rgb.data_ptr(), rgb.size, 2 * w, 2 * h, 6 * w, w, h,
                                             d_nv21.data_ptr(), scratch.data_ptr(), None) == 0
    enc = pb.BsEncoder(0, w, h, pb.FDCT_ISLOW, max_batch=1)
    d_out = torch.zeros(20160, dtype=torch.uint8, device="cuda")
    d_res = torch.zeros(4, dtype=torch.int32, device="cuda")
    enc.encode_device(1, d_nv21, None, 20160, d_out, 20160, d_res, None)
    torch.cuda.synchronize()
    nv21 = d_nv21.cpu().numpy()
    exp_out, exp_res = oracle.Restated().bs_encode_batch(0, w, h, nv21[None], 20160, oracle.FDCT_ISLOW)
    assert np.array_equal(d_res.cpu().numpy(), exp_res[0]) and np.array_equal(d_out.cpu().numpy(), exp_out[0])
    enc.close()
