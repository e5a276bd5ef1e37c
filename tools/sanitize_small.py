"""A small pass over every device path (preprocessing chain, auto-crop, JPEG decode, readtext in bf16x3 / bf16 / fp32) for
`compute-sanitizer --tool memcheck python tools/sanitize_small.py`: sizes are tiny because memcheck runs kernels 10-50x slower.
Prints one line per path; the sanitizer's own report says whether any kernel touched memory it does not own."""
import os
import sys

import cv2
import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import bbocr_b200
    from bbocr_b200 import synth, weights, decode
    from bbocr_b200.preprocess import preprocess_array, CURRENT
    from bbocr_b200.extractor import extract_text_with_ocr

    bgr = synth.phone_photo(3001, 331, 250)                     # odd sizes: the scalar / border branches
    out = preprocess_array(bgr, CURRENT, 0)
    print("preprocess chain", out.shape, int(out.sum()))

    reader = bbocr_b200.Reader(["en"], gpu=True, verbose=False, precision="bf16x3", craft_state=weights.calibrated_craft_state(),
                               crnn_state=weights.calibrated_crnn_state())
    h = reader.handle
    ok, jpg = cv2.imencode(".jpg", synth.book_cover(5, 203, 150), [cv2.IMWRITE_JPEG_QUALITY, 90, cv2.IMWRITE_JPEG_RST_INTERVAL, 4])
    img = decode.imdecode(h, jpg.tobytes())
    print("jpeg decode", None if img is None else img.shape)
    print("auto-crop rect", h.autocrop_rect(synth.sparse_page(11, 275, 200), 16))

    page = synth.title_page(2001, 333, 250)
    for prec in ("bf16x3", "bf16", "fp32"):
        reader.set_precision(prec)
        res = reader.readtext(page)
        print("readtext", prec, len(res), [s for _, s, _ in res][:3])
    reader.set_precision("bf16x3")
    res = reader.readtext_batched([synth.title_page(2002 + i, 320, 256) for i in range(3)])
    print("readtext_batched", [len(r) for r in res])
    print("extract_text_with_ocr", repr(extract_text_with_ocr(reader, synth.phone_photo(3002, 320, 240))[:40]))


if __name__ == "__main__":
    main()
