#!/usr/bin/env python
"""Per-stage wall-clock of readtext over the bench batch (diagnostic)."""
import ctypes as C, os, sys, time
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import torch
import bbocr_b200
from bbocr_b200 import synth
n = int(sys.argv[1]) if len(sys.argv) > 1 else 32
reader = bbocr_b200.Reader(["en"], gpu=0, verbose=False, precision="bf16")
dev = [torch.from_numpy(synth.title_page(2001 + i, 1920, 1440)).cuda() for i in range(n)]
ptrs = [t.data_ptr() for t in dev]
h = reader.handle
out = (C.c_double * 5)()
for rep in range(4):
    t0 = time.perf_counter(); reader.readtext_device(ptrs, 1440, 1920); dt = time.perf_counter() - t0
    h.L.bbocr_dbg_stage_ms(h._h, out)
    print(f"rep {rep}: {n / dt:7.1f} pages/s  per-page ms: craft-enqueue {out[0]/n:.2f}  det(wait CRAFT) {out[1]/n:.2f}  boxes+crops {out[2]/n:.2f}  "
          f"rec1 {out[3]/n:.2f}  rec2 {out[4]/n:.2f}", flush=True)
