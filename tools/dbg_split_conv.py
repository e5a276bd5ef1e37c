#!/usr/bin/env python
"""One split-precision convolution through the debug hook against a float64 reference (tools for compute-sanitizer runs).
usage: python tools/dbg_split_conv.py H W Cin Cout [k=3]"""
import ctypes as C, os, sys
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import numpy as np
from bbocr_b200 import _lib
H, W, ci, co = [int(v) for v in sys.argv[1:5]]
k = int(sys.argv[5]) if len(sys.argv) > 5 else 3
h = _lib.Handle(0); h.set_precision(_lib.PREC_BF16)
rng = np.random.default_rng(0)
x = rng.standard_normal((1, H, W, ci)).astype(np.float32)
w = (rng.standard_normal((co, ci, k, k)) / np.sqrt(ci * k * k)).astype(np.float32)
b = rng.standard_normal(co).astype(np.float32)
out = np.empty((1, H, W, co), np.float32)
p = lambda a: a.ctypes.data_as(C.c_void_p)
rc = h.L.bbocr_dbg_conv(h._h, p(x), C.c_int(ci), None, C.c_int(0), C.c_int(1), C.c_int(H), C.c_int(W), p(w), p(b), C.c_int(co), C.c_int(k), C.c_int(k),
                        C.c_int(k // 2), C.c_int(1), C.c_int(1), C.c_int(2), p(out))
h._check(rc)
import torch
ref = torch.nn.functional.conv2d(torch.from_numpy(x).permute(0, 3, 1, 2).double(), torch.from_numpy(w).double(), torch.from_numpy(b).double(), padding=k // 2)
ref = torch.relu(ref).permute(0, 2, 3, 1).numpy()
print("max abs err", float(np.abs(out - ref).max()), "ref max", float(np.abs(ref).max()))
