"""TEST INFRASTRUCTURE ONLY (oracle): the OCR-input cap of BB-OCR's extractor, restated with the reference's own
library calls.  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may import this.

Reference: pipeline_demo/extractor/enhanced_extractor.py:486-512
    img = Image.open(crop_image_path); max_dim = 1600 if image_index in (None, 0) else 2400
    if max(img.size) > max_dim: img = img.convert("RGB"); img.thumbnail((max_dim, max_dim)); img.save(JPEG, quality=90|95)
Pinned: this IS Pillow's Image.thumbnail (BICUBIC, reducing_gap=2.0) on the same pixels, so parity of the device path
against it is parity against the reference up to the JPEG round trip, which the in-memory path drops on purpose
(bbocr_b200/extractor.py).  The pipeline's input here is the gray preprocessing output; converting a gray image to RGB
before the resize gives three identical planes, so the gray thumbnail equals every channel of the reference's RGB one
(checked in tests/test_oracle_extractor.py).
"""
import numpy as np
from PIL import Image


def ocr_max_dim(image_index=None) -> int:                    # enhanced_extractor.py:494
    return 1600 if (image_index is None or image_index == 0) else 2400


def ocr_input_image(gray: np.ndarray, image_index=None, *, as_reference_rgb=False) -> np.ndarray:
    """enhanced_extractor.py:490-497: thumbnail cap (no JPEG)."""
    img = Image.fromarray(gray)
    m = ocr_max_dim(image_index)
    if max(img.size) > m:
        if as_reference_rgb:
            img = img.convert("RGB")
        img.thumbnail((m, m))
    return np.asarray(img)


def thumbnail_size(W: int, H: int, m: int):
    """Size rule of Image.thumbnail restated (PIL/Image.py::thumbnail): aspect-preserving, floor/ceil by aspect error."""
    import math
    x = y = m
    if x >= W and y >= H:
        return W, H
    aspect = W / H

    def round_aspect(number, key):
        return max(min(math.floor(number), math.ceil(number), key=key), 1)
    if x / y >= aspect:
        x = round_aspect(y * aspect, key=lambda n: abs(aspect - n / y))
    else:
        y = round_aspect(x / aspect, key=lambda n: 0 if n == 0 else abs(aspect - x / n))
    return x, y
