#!/usr/bin/env python
"""Golden `readtext` results of the CPU oracle (oracle/easyocr_restated.py, FP32) at the BENCHED geometries.

    python tests/golden/make_golden_readtext.py [--pages 64] [--threads 8]

The oracle needs ~5-15 s per page, so its results are cached here for the -m gpu end-to-end parity tests
(tests/test_gpu_e2e_parity.py), which compare the library's *bf16* (throughput) and fp32 modes against them:

    readtext_title_1920x1440.json   BASELINE config[1]: synth.title_page(2001 + i, 1920, 1440), i < pages
    readtext_cover_1280x960.json    BASELINE config[0]: synth.book_cover(1001 + i, 1280, 960),  i < pages
    readtext_maps.npz               the oracle's score maps of the first page of either set (float16 storage, plus an
                                    exact float32 copy of every 4th row/column) for the score-map tolerance at full size

Per page: [[box (8 numbers, clockwise from top-left), is_free, text, confidence], ...] in upstream order.  Weights are the
seeded synthetic ones (bbocr_b200/weights.py) -- the same the tests and bench.py load.  Resumable: finished pages are kept.
"""
import argparse
import json
import os
import sys
import time

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

import numpy as np


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--pages", type=int, default=64)
    ap.add_argument("--threads", type=int, default=os.cpu_count() or 1)
    args = ap.parse_args()
    import torch
    torch.set_num_threads(args.threads)
    from bbocr_b200 import synth, weights
    from oracle import easyocr_restated as E
    craft = E.CRAFT(); craft.load_state_dict(weights.to_torch_state(weights.calibrated_craft_state()))
    crnn = E.CRNN(); crnn.load_state_dict(weights.to_torch_state(weights.calibrated_crnn_state()))
    reader = E.Reader(craft, crnn)
    sets = [("readtext_title_1920x1440.json", "title_page", 2001, 1920, 1440),
            ("readtext_cover_1280x960.json", "book_cover", 1001, 1280, 960)]
    maps = {}
    for fname, gen, seed0, W, H in sets:
        path = os.path.join(HERE, fname)
        doc = json.load(open(path)) if os.path.exists(path) else {
            "generator": f"bbocr_b200.synth.{gen}(seed, {W}, {H})", "oracle": "oracle/easyocr_restated.py Reader.readtext (FP32, defaults)",
            "weights": "weights.calibrated_craft_state() / calibrated_crnn_state()", "pages": []}
        done = {p["seed"] for p in doc["pages"]}
        for i in range(args.pages):
            seed = seed0 + i
            page = getattr(synth, gen)(seed, W, H)
            if i == 0:
                t, l, _ = reader.score_maps(page)
                maps[gen + "_text_f16"] = t.astype(np.float16); maps[gen + "_link_f16"] = l.astype(np.float16)
                maps[gen + "_text_sub"] = t[::4, ::4].copy(); maps[gen + "_link_sub"] = l[::4, ::4].copy()
            if seed in done:
                continue
            t0 = time.time()
            res = reader.readtext(page)
            rows = []
            for box, text, conf in res:
                flat = [float(v) for pt in box for v in pt]
                is_free = any(isinstance(v, float) and not float(v).is_integer() for pt in box for v in pt) or \
                    not all(isinstance(v, (int, np.integer)) for pt in box for v in pt)
                rows.append([flat, bool(is_free), text, float(conf)])
            doc["pages"].append({"seed": seed, "results": rows})
            doc["pages"].sort(key=lambda p: p["seed"])
            with open(path + ".tmp", "w") as f:
                json.dump(doc, f, ensure_ascii=False, separators=(",", ":"))
            os.replace(path + ".tmp", path)
            print(f"{gen} {seed}: {len(rows)} regions, {time.time() - t0:.1f} s", flush=True)
    np.savez_compressed(os.path.join(HERE, "readtext_maps.npz"), **maps)


if __name__ == "__main__":
    main()
