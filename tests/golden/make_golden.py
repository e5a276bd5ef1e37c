#!/usr/bin/env python
"""Generate the committed golden fixtures from the read-only reference checkout (/root/reference).

Run in the build container only (the GPU box has no /root/reference):  python tests/golden/make_golden.py

preprocess_<book>.npz holds
  bgr            input image (reference dataset cover, or a 2x INTER_AREA reduction of it for the large ones)
  ref_current_ippoff / ref_current_ippon   reference pipeline_demo preprocess_for_book_cover output (image_preprocessor.py:147-160)
  ref_legacy_ippoff  / ref_legacy_ippon    reference legacy chain output (ocr_testing/preprocessing/image_preprocessor.py:221-252)
  golden_legacy  (book2, book4 only; unscaled inputs)  the PNG the reference author recorded:
                 pipeline_components/img_to_json/ocr_testing/results/images/<book>_preprocessed.png
"""
import importlib.util
import os
import sys
import tempfile

import cv2
import numpy as np

REF = "/root/reference"
HERE = os.path.dirname(os.path.abspath(__file__))


def _load(path, name):
    spec = importlib.util.spec_from_file_location(name, path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def main():
    cur = _load(f"{REF}/pipeline_demo/ocr_testing/preprocessing/image_preprocessor.py", "ref_pp_current")
    leg = _load(f"{REF}/pipeline_components/img_to_json/ocr_testing/preprocessing/image_preprocessor.py", "ref_pp_legacy")
    tmp = tempfile.mkdtemp()
    for book, shrink in (("book2", 1), ("book4", 1), ("book1", 2), ("book5", 2), ("book6", 2)):
        src = f"{REF}/pipeline_components/books/dataset/{book}.png"
        bgr = cv2.imread(src)
        if shrink > 1:
            bgr = cv2.resize(bgr, (bgr.shape[1] // shrink, bgr.shape[0] // shrink), interpolation=cv2.INTER_AREA)
        p = os.path.join(tmp, f"{book}.png")
        cv2.imwrite(p, bgr)
        out = {"bgr": bgr}
        for ipp in (False, True):
            cv2.ipp.setUseIPP(ipp)
            tag = "ippon" if ipp else "ippoff"
            out[f"ref_current_{tag}"] = cur.preprocess_for_book_cover(p)[0]
            out[f"ref_legacy_{tag}"] = leg.preprocess_for_book_cover(p)[0]
        cv2.ipp.setUseIPP(True)
        if shrink == 1:
            out["golden_legacy"] = cv2.imread(
                f"{REF}/pipeline_components/img_to_json/ocr_testing/results/images/{book}_preprocessed.png",
                cv2.IMREAD_UNCHANGED)
        np.savez_compressed(os.path.join(HERE, f"preprocess_{book}.npz"), **out)
        print(book, bgr.shape, {k: v.shape for k, v in out.items()})


if __name__ == "__main__":
    sys.exit(main())
