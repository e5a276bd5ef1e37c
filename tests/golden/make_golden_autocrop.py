#!/usr/bin/env python
"""Generate tests/golden/autocrop_*.npz by running the reference's OWN crop heuristics from the read-only checkout.

Run in the build container only (the GPU box has no /root/reference):  python tests/golden/make_golden_autocrop.py

pipeline_demo/extractor/enhanced_extractor.py cannot be imported (easyocr / pytesseract / jsonschema are absent), so
the two methods `_auto_crop_text_region` (:239-372) and `_central_edge_crop` (:374-397) are cut out of its source with
`ast` and executed unmodified against a stub `self`; `cv2.imwrite` is intercepted to capture the crop they write.
Each fixture holds
  bgr        the input page (a reference dataset cover reduced with INTER_AREA, or a synthetic page)
  margins    the margins tried
  rects      per margin: (x0, y0, x1, y1) of the reference's crop, or (-1,-1,-1,-1) when it returned None.  The rectangle
             is recovered from the captured crop: it is the oracle's rectangle, accepted only if bgr[y0:y1, x0:x1] is
             byte-identical to what the reference wrote (the script fails otherwise)
  edge       rows (percent, x0, y0, x1, y1) for _central_edge_crop, -1s for None
"""
import ast
import os
import sys
import tempfile
import textwrap
import types

import cv2
import numpy as np

REF = "/root/reference"
HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

from bbocr_b200 import synth                      # noqa: E402
from oracle import autocrop_np as A               # noqa: E402


def reference_methods():
    path = f"{REF}/pipeline_demo/extractor/enhanced_extractor.py"
    src = open(path).read()
    tree = ast.parse(src)
    cls = next(n for n in tree.body if isinstance(n, ast.ClassDef) and n.name == "EnhancedBookMetadataExtractor")
    captured = {}

    class Cv2Proxy:
        def __getattr__(self, k):
            return getattr(cv2, k)

        @staticmethod
        def imwrite(p, arr):
            captured["path"], captured["arr"] = p, arr.copy()
            return True

    ns = {"cv2": Cv2Proxy(), "os": os, "np": np}
    exec("from typing import *", ns)
    for fn in cls.body:
        if isinstance(fn, ast.FunctionDef) and fn.name in ("_auto_crop_text_region", "_central_edge_crop"):
            exec(textwrap.dedent(ast.get_source_segment(src, fn)), ns)
    return ns["_auto_crop_text_region"], ns["_central_edge_crop"], captured


def main():
    auto, edge, captured = reference_methods()
    tmp = tempfile.mkdtemp()
    stub = types.SimpleNamespace(_get_temp_dir=lambda: tmp, debug_autocrop=False)
    pages = {}
    for book, width in (("book1", 600), ("book2", 480), ("book4", 500), ("book5", 640), ("book6", 560)):
        bgr = cv2.imread(f"{REF}/pipeline_components/books/dataset/{book}.png")
        h = int(round(bgr.shape[0] * width / bgr.shape[1]))
        pages[book] = cv2.resize(bgr, (width, h), interpolation=cv2.INTER_AREA)
    photo = cv2.imread(f"{REF}/pipeline_components/books/2a/IMG_9684.JPG")
    if photo is not None:
        pages["img9684"] = cv2.resize(photo, (756, 567) if photo.shape[1] > photo.shape[0] else (567, 756),
                                      interpolation=cv2.INTER_AREA)
    pages["sparse"] = synth.sparse_page(11, 1100, 800)
    pages["framed"] = synth.sparse_page(12, 900, 1200, frame=True)
    for name, bgr in pages.items():
        bgr = np.ascontiguousarray(bgr[:, :, :3])
        p = os.path.join(tmp, f"{name}.png")
        cv2.imwrite(p, bgr)
        margins = [0, 16, 40]
        rects = []
        for m in margins:
            captured.clear()
            ret = auto(stub, p, m)
            mine = A.auto_crop_rect(bgr, m)
            if ret is None:
                assert mine is None, (name, m, mine)
                rects.append((-1, -1, -1, -1))
            else:
                x0, y0, x1, y1 = mine
                assert np.array_equal(captured["arr"], bgr[y0:y1, x0:x1]), (name, m, mine, captured["arr"].shape)
                rects.append(mine)
        edges = []
        for pc in (0.0, 2.0, 5.5, 12.0, 41.0):
            captured.clear()
            ret = edge(stub, p, pc)
            mine = A.central_edge_crop_rect(bgr.shape[0], bgr.shape[1], pc)
            if ret is None:
                assert mine is None, (name, pc, mine)
                edges.append((pc, -1, -1, -1, -1))
            else:
                x0, y0, x1, y1 = mine
                assert np.array_equal(captured["arr"], bgr[y0:y1, x0:x1]), (name, pc)
                edges.append((pc,) + tuple(mine))
        np.savez_compressed(os.path.join(HERE, f"autocrop_{name}.npz"), bgr=bgr, margins=np.array(margins),
                            rects=np.array(rects, np.int32), edge=np.array(edges, np.float64))
        print(name, bgr.shape, rects, [e[1:] for e in edges])


if __name__ == "__main__":
    sys.exit(main())
