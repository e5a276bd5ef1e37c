#!/usr/bin/env python
"""The in-memory extractor glue on phone photos: extract_text_with_ocr = preprocess_for_book_cover ->
[auto-crop] -> OCR-input cap (PIL thumbnail 1600) -> readtext -> joined text, host BGR array in, string out.
The reference does the same through a PNG, a JPEG and three decodes (~1.7 s preprocessing + ~1 s I/O per photo on the CPU
before EasyOCR starts).  Prints one JSON line."""
import json
import os
import sys
import time

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import bbocr_b200
from bbocr_b200 import extractor, synth

n = int(sys.argv[1]) if len(sys.argv) > 1 else 6
reader = bbocr_b200.Reader(["en"], gpu=0, verbose=False, precision="bf16")
photos = [synth.phone_photo(3001 + i) for i in range(min(n, 3))]
photos = [photos[i % len(photos)] for i in range(n)]
out = {}
for name, kw in (("plain", {}), ("with_auto_crop", {"crop_for_ocr": True, "crop_margin": 16})):
    extractor.extract_text_with_ocr(reader, photos[0], image_index=0, **kw)
    t0 = time.perf_counter()
    chars = 0
    for p in photos:
        chars += len(extractor.extract_text_with_ocr(reader, p, image_index=0, **kw))
    dt = (time.perf_counter() - t0) / n
    out[name] = {"ms_per_photo": dt * 1e3, "photos_per_s": 1 / dt, "chars": chars}
print(json.dumps({"workload": f"{n} synthetic 4032x3024 phone photos, host BGR arrays, one photo per call (the extractor's own call pattern)",
                  **out}))
