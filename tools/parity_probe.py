#!/usr/bin/env python
"""GPU exploration: how far is the throughput (bf16) mode from the FP32 mode END TO END at the benched geometries?

    python tools/parity_probe.py [--pages 64] > gpurun_out/parity_probe.json

For the title pages (1920x1440) and covers (1280x960): readtext_batched in fp32 and in bf16; counts identical boxes and
strings; score-map max-abs error and the number of pixels on the other side of the 0.4 / 0.7 thresholds for two pages of
each kind.  Also dumps the fp32 results so they can be diffed against the CPU oracle's golden file off the box.
"""
import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import numpy as np


def key(r):
    return tuple(np.round(np.asarray(r[0], float).reshape(-1), 6).tolist())


def compare(a, b):
    """a, b: per-page result lists.  -> dict of counts (reference = a)."""
    n_ref = n_box = n_both = 0
    n_pages_same = 0
    conf_err = 0.0
    for ra, rb in zip(a, b):
        mb = {key(r): r for r in rb}
        page_same = len(ra) == len(rb)
        for r in ra:
            n_ref += 1
            o = mb.get(key(r))
            if o is None:
                page_same = False
                continue
            n_box += 1
            if o[1] == r[1]:
                n_both += 1
                conf_err = max(conf_err, abs(o[2] - r[2]))
            else:
                page_same = False
        n_pages_same += page_same
    return {"regions_ref": n_ref, "regions_other": sum(len(r) for r in b), "boxes_identical": n_box,
            "box_and_string_identical": n_both, "pages_fully_identical": n_pages_same, "max_conf_err_on_identical": conf_err}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--pages", type=int, default=64)
    ap.add_argument("--modes", default="bf16")
    args = ap.parse_args()
    import bbocr_b200
    from bbocr_b200 import synth
    reader = bbocr_b200.Reader(["en"], gpu=True, verbose=False, precision="fp32")
    out = {}
    dump = {}
    for name, gen, seed0, W, H in (("title", synth.title_page, 2001, 1920, 1440), ("cover", synth.book_cover, 1001, 1280, 960)):
        pages = [gen(seed0 + i, W, H) for i in range(args.pages)]
        reader.set_precision("fp32")
        t0 = time.time()
        ref = reader.readtext_batched(pages)
        t_fp32 = time.time() - t0
        dump[name] = [[[np.asarray(b, float).reshape(-1).tolist(), t, c] for b, t, c in r] for r in ref]
        entry = {"pages": args.pages, "fp32_s": t_fp32}
        maps_ref = [reader.score_maps(pages[i]) for i in range(2)]
        for mode in args.modes.split(","):
            reader.set_precision(mode)
            got = reader.readtext_batched(pages)
            reader.readtext_batched(pages)
            t0 = time.time()
            reader.readtext_batched(pages)
            dt = time.time() - t0
            e = compare(ref, got)
            e["pages_per_s_host_arrays"] = args.pages / dt
            errs = []
            for i in range(2):
                t, l, _ = reader.score_maps(pages[i])
                rt, rl, _ = maps_ref[i]
                errs.append({"text_maxabs": float(np.abs(t - rt).max()), "link_maxabs": float(np.abs(l - rl).max()),
                             "text_mean_abs": float(np.abs(t - rt).mean()),
                             "flips_text_0.4": int(((t > 0.4) != (rt > 0.4)).sum()), "flips_link_0.4": int(((l > 0.4) != (rl > 0.4)).sum()),
                             "flips_text_0.7": int(((t > 0.7) != (rt > 0.7)).sum()), "fg_pixels": int((rt > 0.4).sum())})
            e["score_maps"] = errs
            entry[mode] = e
        out[name] = entry
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    json.dump(dump, open(os.path.join(ROOT, "gpurun_out", "gpu_fp32_results.json"), "w"))
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
