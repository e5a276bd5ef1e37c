#!/usr/bin/env python
"""Where the end-to-end (host arrays in, Python tuples out) time of a 64-page batch goes: C call vs Python shim."""
import os, sys, time
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import numpy as np, torch
import bbocr_b200
from bbocr_b200 import synth
n = int(sys.argv[1]) if len(sys.argv) > 1 else 64
reader = bbocr_b200.Reader(["en"], gpu=0, verbose=False, precision="bf16")
pages = [synth.title_page(2001 + i, 1920, 1440) for i in range(n)]
dev = [torch.from_numpy(p).cuda() for p in pages]
ptrs = [t.data_ptr() for t in dev]
h = reader.handle
p, _ = reader._params({})
host_tuples = [(pg, None, 1440, 1920) for pg in pages]
dev_tuples = [(ptr, None, 1440, 1920) for ptr in ptrs]
for rep in range(4):
    t0 = time.perf_counter(); raw = h.readtext_raw(dev_tuples, p, on_device=True); t1 = time.perf_counter()
    res = [reader._format(r) for r, _ in raw]; t2 = time.perf_counter()
    raw = h.readtext_raw(host_tuples, p); t3 = time.perf_counter()
    res = reader.readtext_batched(pages); t4 = time.perf_counter()
    print(f"rep {rep}: resident raw {1e3*(t1-t0):6.1f} ms (+format {1e3*(t2-t1):5.1f})  host raw {1e3*(t3-t2):6.1f} ms  readtext_batched {1e3*(t4-t3):6.1f} ms"
          f"  -> {n/(t1-t0):6.1f} / {n/(t3-t2):6.1f} / {n/(t4-t3):6.1f} pages/s", flush=True)
