#!/usr/bin/env python
"""One readtext over N synthetic 1920x1440 title pages (bench workload), for ncu launch lists / captures.
usage: python tools/profile_page.py [pages=2] [precision=bf16]"""
import os, sys
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import torch
import bbocr_b200
from bbocr_b200 import synth

n = int(sys.argv[1]) if len(sys.argv) > 1 else 2
prec = sys.argv[2] if len(sys.argv) > 2 else "bf16"
reader = bbocr_b200.Reader(["en"], gpu=0, verbose=False, precision=prec)
pages = [synth.title_page(2001 + i, 1920, 1440) for i in range(n)]
dev = [torch.from_numpy(p).cuda() for p in pages]
torch.cuda.synchronize()
for t in dev:                                   # one page at a time: a serial launch list per page
    res, stats = reader.readtext_device([t.data_ptr()], 1440, 1920)
    print(len(res[0]), stats, reader.handle.launch_count())
torch.cuda.synchronize()
