"""GPU: the tcgen05 implicit-GEMM convolution against the CUDA-core kernel on identical bf16 operands, layer shape by
layer shape (every distinct geometry CRAFT / CRNN use), and against a float64 NumPy convolution."""
import ctypes as C

import numpy as np
import pytest
import torch
import torch.nn.functional as F

from bbocr_b200 import _lib

pytestmark = pytest.mark.gpu

SHAPES = [
    # N, H, W, C1, C2, cout, k, pad, dil
    (1, 24, 40, 64, 0, 64, 3, 1, 1),        # conv1_2-like, partial 16x8 tiles
    (2, 16, 16, 128, 0, 256, 3, 1, 1),      # two N tiles of 128
    (1, 9, 13, 1024, 512, 512, 1, 0, 1),    # upconv1 1x1 over a channel concat (flat view, two tensor maps)
    (1, 12, 20, 512, 0, 1024, 3, 6, 6),     # fc6: dilation 6, padding 6
    (1, 16, 24, 32, 0, 16, 3, 1, 1),        # conv_cls: BK = 32 (SWIZZLE_64B), N = 16
    (1, 16, 24, 32, 0, 32, 3, 1, 1),
    (3, 4, 20, 256, 0, 256, 2, 0, 1),       # CRNN conv6: 2x2, no padding, 32x4 tiles
    (5, 1, 15, 256, 0, 2048, 1, 0, 1),      # LSTM input projection
    (2, 1, 63, 256, 0, 97, 1, 0, 1),        # Prediction: cout padded to 112
    (3, 32, 36, 64, 0, 128, 3, 1, 1),       # CRNN conv3, batch of crops
    (1, 30, 40, 256, 512, 256, 1, 0, 1),    # upconv2 1x1 concat
    (1, 8, 8, 512, 0, 512, 3, 1, 1),
    (1, 70, 44, 64, 0, 64, 3, 1, 1),        # conv_res.cu: two stacked M tiles per patch, partial tiles on both edges
    (2, 40, 24, 128, 0, 64, 3, 1, 1),       # conv_res.cu: two 64-channel k-blocks per tile
    (1, 34, 20, 64, 0, 32, 3, 1, 1),        # conv_res.cu: N = 32
    (1, 48, 32, 128, 0, 128, 3, 1, 1),      # conv_res.cu: two resident channel slices of 64
    (1, 36, 28, 32, 0, 32, 3, 1, 1),        # conv_res.cu: 32-channel k-block (SWIZZLE_64B patch), conv_cls geometry
    (1, 34, 26, 32, 0, 64, 1, 0, 1),        # conv_res.cu: 1x1 over the 32-channel gathered stem (conv1_1)
]


def dbg_conv(handle, x1, x2, w, b, pad, dil, relu, force_generic):
    L = handle.L
    N, H, W, C1 = x1.shape
    C2 = 0 if x2 is None else x2.shape[3]
    cout, cin, kh, kw = w.shape
    OH, OW = H + 2 * pad - dil * (kh - 1), W + 2 * pad - dil * (kw - 1)
    out = np.empty((N, OH, OW, cout), np.float32)
    p = lambda a: a.ctypes.data_as(C.c_void_p) if a is not None else None
    rc = L.bbocr_dbg_conv(handle._h, p(x1), C.c_int(C1), p(x2), C.c_int(C2), C.c_int(N), C.c_int(H), C.c_int(W), p(w), p(b),
                          C.c_int(cout), C.c_int(kh), C.c_int(kw), C.c_int(pad), C.c_int(dil), C.c_int(int(relu)),
                          C.c_int(int(force_generic)), p(out))
    handle._check(rc)
    return out


@pytest.mark.parametrize("shape", SHAPES)
def test_tc_conv_matches_cuda_core_and_numpy(handle, shape):
    N, H, W, C1, C2, cout, k, pad, dil = shape
    rng = np.random.default_rng(abs(hash(shape)) % (2 ** 32))
    x1 = np.ascontiguousarray(rng.standard_normal((N, H, W, C1)), np.float32)
    x2 = np.ascontiguousarray(rng.standard_normal((N, H, W, C2)), np.float32) if C2 else None
    w = np.ascontiguousarray(rng.standard_normal((cout, C1 + C2, k, k)) / np.sqrt((C1 + C2) * k * k), np.float32)
    b = np.ascontiguousarray(rng.standard_normal(cout) * 0.1, np.float32)
    handle.set_precision(_lib.PREC_BF16)
    try:
        tc = dbg_conv(handle, x1, x2, w, b, pad, dil, True, False)
        cc = dbg_conv(handle, x1, x2, w, b, pad, dil, True, True)
    finally:
        handle.set_precision(_lib.PREC_FP32)
    # same bf16 operands, FP32 accumulation in both: only summation order and the final bf16 rounding may differ
    assert np.abs(tc - cc).max() <= 2 ** -7 * max(1.0, np.abs(cc).max())
    # and both agree with an exact convolution of the bf16-rounded operands
    xb = torch.from_numpy(np.concatenate([x1] + ([x2] if C2 else []), 3)).bfloat16().double().permute(0, 3, 1, 2)
    wb = torch.from_numpy(w).bfloat16().double()
    ref = F.relu(F.conv2d(xb, wb, torch.from_numpy(b).double(), padding=pad, dilation=dil)).permute(0, 2, 3, 1).numpy()
    assert np.abs(tc - ref).max() <= 2 ** -7 * max(1.0, np.abs(ref).max())


@pytest.mark.parametrize("shape", [s for s in SHAPES if s[4] == 0])
def test_split_precision_conv_is_fp32_class(handle, shape):
    """3 x bf16 (hi/lo) convolution on the tensor cores: x_hi*w_hi + x_lo*w_hi + x_hi*w_lo, FP32 accumulate -> matches an
    exact convolution of the FP32 operands to ~1e-4 relative (vs ~1e-2 for plain bf16)."""
    N, H, W, C1, C2, cout, k, pad, dil = shape
    rng = np.random.default_rng(abs(hash(shape)) % (2 ** 32))
    x1 = np.ascontiguousarray(rng.standard_normal((N, H, W, C1)), np.float32)
    w = np.ascontiguousarray(rng.standard_normal((cout, C1, k, k)) / np.sqrt(C1 * k * k), np.float32)
    b = np.ascontiguousarray(rng.standard_normal(cout) * 0.1, np.float32)
    handle.set_precision(_lib.PREC_BF16)
    try:
        got = dbg_conv(handle, x1, None, w, b, pad, dil, True, 2)
    finally:
        handle.set_precision(_lib.PREC_FP32)
    ref = F.relu(F.conv2d(torch.from_numpy(x1).double().permute(0, 3, 1, 2), torch.from_numpy(w).double(),
                          torch.from_numpy(b).double(), padding=pad, dilation=dil)).permute(0, 2, 3, 1).numpy()
    err = np.abs(got - ref).max() / max(1.0, np.abs(ref).max())
    assert err < 2e-4, err


def test_resident_weight_kernel_in_subprocess():
    """conv_res.cu (resident weights + halo patch) only takes layers with >= 75 k output pixels by default; BBOCR_RES=2
    forces it for every supported geometry, so the same shape sweep runs through it (mode latched per process)."""
    import os
    import subprocess
    import sys
    env = dict(os.environ, BBOCR_RES="2")
    r = subprocess.run([sys.executable, "-m", "pytest", __file__, "-q", "-m", "gpu", "-k", "matches_cuda_core"], env=env,
                       capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
