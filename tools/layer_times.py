#!/usr/bin/env python
"""CUDA-event timing of single conv layers through the debug hook (BF16 mode), CRAFT shapes at 1920x1440."""
import ctypes as C, os, sys, time
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import numpy as np, torch
from bbocr_b200 import _lib
h = _lib.Handle(0); h.set_precision(_lib.PREC_BF16)
L = h.L
shapes = [("conv1_1", 1440, 1920, 32, 64, 1), ("cls0", 720, 960, 32, 32), ("conv1_2", 1440, 1920, 64, 64), ("conv2_1", 720, 960, 64, 128), ("conv2_2", 720, 960, 128, 128), ("conv3_1", 360, 480, 128, 256), ("up3b", 360, 480, 128, 64),
          ("up4b", 720, 960, 64, 32), ("conv3_2", 360, 480, 256, 256), ("conv4_2", 180, 240, 512, 512)]
for name, H, W, ci, co, *kk in shapes:
    k = kk[0] if kk else 3
    rng = np.random.default_rng(0)
    x = rng.standard_normal((1, H, W, ci)).astype(np.float32)
    w = (rng.standard_normal((co, ci, k, k)) / np.sqrt(ci * k * k)).astype(np.float32)
    b = np.zeros(co, np.float32); out = np.empty((1, H, W, co), np.float32)
    p = lambda a: a.ctypes.data_as(C.c_void_p)
    h.enable_conv_timing(True)
    for rep in range(3):
        rc = L.bbocr_dbg_conv(h._h, p(x), C.c_int(ci), None, C.c_int(0), C.c_int(1), C.c_int(H), C.c_int(W), p(w), p(b), C.c_int(co), C.c_int(k), C.c_int(k), C.c_int(k // 2), C.c_int(1), C.c_int(1), C.c_int(0), p(out))
        h._check(rc)
        ms, n, fl = h.conv_stats()
    print(f"{name}: {ms*1e3:8.1f} us  {fl/ms/1e9:7.1f} TFLOP/s", flush=True)
