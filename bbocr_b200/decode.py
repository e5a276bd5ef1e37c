"""Image decode on the device: `imread` / `imdecode` with cv2's names, flags and results for baseline JPEG.

Reference call sites: pipeline_demo/ocr_testing/preprocessing/image_preprocessor.py:18 (`cv2.imread(image_path)`, ValueError
when it returns None) and easyocr/utils.py::reformat_input (one IMREAD_GRAYSCALE and one colour read per file).  The JPEG
path is bit-exact with cv2 (oracle/jpeg_np.py pins the arithmetic); anything libbbocr refuses (PNG, progressive JPEG, ...) is
decoded by cv2 on the host exactly as before -- decoding is the edge of the path, not the path."""
from __future__ import annotations

import os

import cv2
import numpy as np

from . import _lib

IMREAD_GRAYSCALE = cv2.IMREAD_GRAYSCALE
IMREAD_COLOR = cv2.IMREAD_COLOR
IMREAD_IGNORE_ORIENTATION = cv2.IMREAD_IGNORE_ORIENTATION


def _is_jpeg(data) -> bool:
    return len(data) > 3 and data[0] == 0xFF and data[1] == 0xD8


def imdecode_both(handle: "_lib.Handle", data: bytes, ignore_orientation: bool = False):
    """One device decode -> (BGR HxWx3, gray HxW): the two reads reformat_input makes of a file."""
    return handle.jpeg_decode(data, color=True, gray=True, ignore_orientation=ignore_orientation)


def imdecode(handle: "_lib.Handle", buf, flags: int = cv2.IMREAD_COLOR):
    """cv2.imdecode(buf, flags) for flags in {IMREAD_COLOR, IMREAD_GRAYSCALE} (| IMREAD_IGNORE_ORIENTATION)."""
    data = bytes(buf) if not isinstance(buf, (bytes, bytearray)) else buf
    base = flags & ~cv2.IMREAD_IGNORE_ORIENTATION
    if _is_jpeg(data) and base in (cv2.IMREAD_COLOR, cv2.IMREAD_GRAYSCALE):
        try:
            bgr, gray = handle.jpeg_decode(data, color=base == cv2.IMREAD_COLOR, gray=base == cv2.IMREAD_GRAYSCALE,
                                           ignore_orientation=bool(flags & cv2.IMREAD_IGNORE_ORIENTATION))
            return bgr if base == cv2.IMREAD_COLOR else gray
        except _lib.BbocrError:                                 # unsupported layout or damaged stream: cv2 decides
            pass
    return cv2.imdecode(np.frombuffer(data, np.uint8), flags)


def imread(handle: "_lib.Handle", path: str, flags: int = cv2.IMREAD_COLOR):
    """cv2.imread(path, flags): None when the file cannot be read, like cv2."""
    try:
        with open(os.path.expanduser(path), "rb") as f:
            data = f.read()
    except OSError:
        return None
    return imdecode(handle, data, flags)
