"""Shared helpers of the JPEG tests: synthetic photos encoded with cv2 / Pillow in every layout the decoder supports."""
import io

import cv2
import numpy as np

from bbocr_b200 import synth


def _photo(h, w, seed=5):
    return np.ascontiguousarray(synth.book_cover(seed, max(w, 64), max(h, 64))[:h, :w])


SAMPLINGS = {"420": cv2.IMWRITE_JPEG_SAMPLING_FACTOR_420, "422": cv2.IMWRITE_JPEG_SAMPLING_FACTOR_422,
             "444": cv2.IMWRITE_JPEG_SAMPLING_FACTOR_444, "440": cv2.IMWRITE_JPEG_SAMPLING_FACTOR_440}


def encode(img, q=90, samp="420", rst=0):
    ok, buf = cv2.imencode(".jpg", img, [cv2.IMWRITE_JPEG_QUALITY, q, cv2.IMWRITE_JPEG_SAMPLING_FACTOR, SAMPLINGS[samp],
                                         cv2.IMWRITE_JPEG_RST_INTERVAL, rst])
    assert ok
    return buf.tobytes()


def with_orientation(img_bgr, o, q=85):
    from PIL import Image
    ex = Image.Exif()
    ex[0x0112] = o
    b = io.BytesIO()
    Image.fromarray(img_bgr[:, :, ::-1]).save(b, "JPEG", quality=q, exif=ex.tobytes())
    return b.getvalue()
