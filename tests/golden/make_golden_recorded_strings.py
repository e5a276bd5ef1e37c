#!/usr/bin/env python
"""The reference's only recorded EasyOCR outputs (never asserted upstream): the joined text of
pipeline_components/img_to_json/ocr_testing/results/json/ocr_comparison_*.json -> a small fixture
tests/golden/easyocr_recorded_strings.json for tests/test_oracle_pin.py.  Run in the build container (reads /root/reference)."""
import glob
import json
import os

HERE = os.path.dirname(os.path.abspath(__file__))
REF = "/root/reference/pipeline_components/img_to_json/ocr_testing"
out = []
for f in sorted(glob.glob(os.path.join(REF, "results/json/ocr_comparison_*.json"))):
    d = json.load(open(f))
    e = d.get("easyocr") or {}
    if not e.get("text"):
        continue
    out.append({"record": os.path.basename(f), "image_path_recorded": d.get("image_path"), "preprocessing_used": d.get("preprocessing_used"),
                "text": e["text"], "text_length": e.get("text_length"), "processing_time_s": e.get("processing_time")})
json.dump({"source": "pipeline_components/img_to_json/ocr_testing/results/json (EasyOCR >= 1.7.0, default readtext args, CPU int8 path, Windows host)",
           "records": out}, open(os.path.join(HERE, "easyocr_recorded_strings.json"), "w"), indent=1, ensure_ascii=False)
print(len(out), "records")
