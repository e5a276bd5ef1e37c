#!/usr/bin/env python
"""One batched readtext over N synthetic 1920x1440 title pages through bbocr_readtext_batch (detector lanes + grouped
recogniser), for ncu launch lists.  usage: python tools/profile_batch.py [pages=16] [precision=bf16]"""
import os, sys, time
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import torch
import bbocr_b200
from bbocr_b200 import synth

n = int(sys.argv[1]) if len(sys.argv) > 1 else 16
prec = sys.argv[2] if len(sys.argv) > 2 else "bf16"
reader = bbocr_b200.Reader(["en"], gpu=0, verbose=False, precision=prec)
pages = [synth.title_page(2001 + i, 1920, 1440) for i in range(n)]
dev = [torch.from_numpy(p).cuda() for p in pages]
torch.cuda.synchronize()
t0 = time.perf_counter()
res, stats = reader.readtext_device([t.data_ptr() for t in dev], 1440, 1920)
dt = time.perf_counter() - t0
print(n, "pages", sum(len(r) for r in res), "regions", sum(s["n_crops"] for s in stats), "crops", reader.handle.launch_count(),
      "launches", f"{dt * 1e3:.1f} ms")
