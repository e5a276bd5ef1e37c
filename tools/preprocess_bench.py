#!/usr/bin/env python
"""Stage-1 throughput on BASELINE config 3 (synthetic 4032x3024 phone photos, resident in HBM): the reference chain
`preprocess_for_book_cover` (image_preprocessor.py:147-160) through bbocr_preprocess_u8 with device pointers, and the
auto-crop heuristic (enhanced_extractor.py:239-372) through bbocr_autocrop_rect.  Prints one JSON line.

Algorithmic bytes: 12 B per input pixel for the chain (read BGR 3 + write/read/read/write of the x1.5 gray
plane 4 x 2.25).  For auto-crop the dependency minimum is read BGR 3 + the equalised plane written once and read twice
(CLAHE needs all tile histograms first) 3 + 4 B label traffic per pixel = 10 B per pixel."""
import ctypes as C
import json
import os
import sys
import time

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import numpy as np
import torch

from bbocr_b200 import _lib, synth
from bbocr_b200.preprocess import CURRENT, pp_params

n_photos = int(sys.argv[1]) if len(sys.argv) > 1 else 12
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
H, W = 3024, 4032
h = _lib.Handle(0)
base = [torch.from_numpy(np.ascontiguousarray(synth.phone_photo(3001 + i, W, H))).cuda() for i in range(min(n_photos, 4))]
photos = [base[i % len(base)].clone() for i in range(n_photos)]          # 36.6 MB each: > L2 (126 MB) from 4 photos on
p = pp_params(CURRENT, 0)
dH, dW = int(H * 1.5), int(W * 1.5)
outs = [torch.empty((dH, dW), dtype=torch.uint8, device="cuda") for _ in range(2)]
bouts = [torch.empty((dH, dW), dtype=torch.uint8, device="cuda") for _ in range(n_photos)]
oh, ow = C.c_int(), C.c_int()


def chain(i):
    rc = h.L.bbocr_preprocess_u8(h._h, C.c_void_p(photos[i].data_ptr()), C.c_int(H), C.c_int(W), C.c_int(W * 3), C.c_int(1),
                                 C.byref(p), C.c_void_p(outs[i & 1].data_ptr()), C.c_int(1), C.byref(oh), C.byref(ow))
    assert rc == 0, h.L.bbocr_last_error(h._h)


def chain_batch():
    h.preprocess_batch_dev([t.data_ptr() for t in photos], H, W, p, [bouts[i].data_ptr() for i in range(n_photos)])


rect = (C.c_int32 * 4)()
found = C.c_int()


def crop(i):
    rc = h.L.bbocr_autocrop_rect(h._h, C.c_void_p(photos[i].data_ptr()), C.c_int(H), C.c_int(W), C.c_int(3), C.c_int(W * 3),
                                 C.c_int(1), C.c_int(16), rect, C.byref(found), None, None, None, C.c_int(0), None, None)
    assert rc == 0, h.L.bbocr_last_error(h._h)


def timed(fn):
    for i in range(min(3, n_photos)):
        fn(i)
    torch.cuda.synchronize()
    best = None
    for _ in range(reps):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for i in range(n_photos):
            fn(i)
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        best = dt if best is None or dt < best else best
    return best / n_photos


peaks = json.load(open(os.path.join(os.path.dirname(__file__), "..", "MEASURED_PEAKS.json"))) if os.path.exists(
    os.path.join(os.path.dirname(__file__), "..", "MEASURED_PEAKS.json")) else {}
hbm = float(peaks.get("hbm_gbs", 6557.4))
scan_out = torch.empty((H, W), dtype=torch.uint8, device="cuda")


def scan(i):
    h.preprocess_scan_dev(photos[i].data_ptr(), H, W, scan_out.data_ptr())


t_scan = timed(scan)
t_chain = timed(chain)
chain_batch()
torch.cuda.synchronize()
t_batch = None
for _ in range(reps):
    t0 = time.perf_counter()
    chain_batch()
    torch.cuda.synchronize()
    dt = (time.perf_counter() - t0) / n_photos
    t_batch = dt if t_batch is None or dt < t_batch else t_batch
for i in range(n_photos):                                     # the batched call must give the same bytes as the single one
    chain(i)
    assert torch.equal(outs[i & 1], bouts[i]), i
t_crop = timed(crop)
px = H * W
print(json.dumps({
    "workload": f"{n_photos} synthetic {W}x{H} phone photos resident in HBM (BASELINE config[2]); timing = host wall clock around "
                "the blocking C-ABI calls, best of %d passes" % reps,
    "preprocess_chain": {"ms_per_photo": t_chain * 1e3, "photos_per_s": 1 / t_chain, "algorithmic_GBps": 12 * px / t_chain / 1e9,
                         "frac_of_hbm_peak": 12 * px / t_chain / 1e9 / hbm,
                         "launches_per_photo": h.L.bbocr_preprocess_launches_per_image()},
    "preprocess_chain_batched": {"ms_per_photo": t_batch * 1e3, "photos_per_s": 1 / t_batch,
                                 "algorithmic_GBps": 12 * px / t_batch / 1e9, "frac_of_hbm_peak": 12 * px / t_batch / 1e9 / hbm,
                                 "api": "bbocr_preprocess_batch_u8, device pointers, photos spread over 8 streams"},
    "scan_chain": {"what": "BASELINE config[2] literally: gray -> CLAHE -> deskew -> adaptive threshold (bbocr_preprocess_scan_u8); "
                           "8 B per pixel algorithmic (SURVEY.md 8d)", "ms_per_photo": t_scan * 1e3, "photos_per_s": 1 / t_scan,
                   "algorithmic_GBps": 8 * px / t_scan / 1e9, "frac_of_hbm_peak": 8 * px / t_scan / 1e9 / hbm},
    "autocrop": {"ms_per_photo": t_crop * 1e3, "photos_per_s": 1 / t_crop, "algorithmic_GBps": 10 * px / t_crop / 1e9,
                 "frac_of_hbm_peak": 10 * px / t_crop / 1e9 / hbm},
    "hbm_peak_GBps": hbm}))
