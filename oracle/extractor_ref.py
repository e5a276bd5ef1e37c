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


# ---- NumPy restatement of Image.thumbnail for one u8 plane (pinned to Pillow by tests/test_oracle_extractor.py) ---------------
# PIL/Image.py::resize(reducing_gap=2.0): factor = int(src / dst / 2.0) or 1 per axis; if a factor exceeds 1 the image is first
# box-reduced (libImaging/Reduce.c) and the bicubic pass (libImaging/Resample.c) then samples the reduced image through the
# fractional box (0, 0, W / fx, H / fy), which the C entry point receives as single-precision floats.

def reduce_np(a: np.ndarray, fx: int, fy: int) -> np.ndarray:
    """ImagingReduce, mode L: every output pixel = ((amend + sum of its fx x fy block clipped to the image) * mult) >> 24 with
    amend = count // 2 and mult = (UINT32)(2^32f / (256 * count)) in float arithmetic (division_UINT32); the dedicated 2x2 / 4x4
    shift kernels and ImagingReduceCorners' partial edge blocks are the same expression."""
    H, W = a.shape
    oh, ow = (H + fy - 1) // fy, (W + fx - 1) // fx
    csum = np.zeros((H + 1, W + 1), np.int64)
    csum[1:, 1:] = a.astype(np.int64).cumsum(0).cumsum(1)
    ys = np.arange(oh) * fy
    ye = np.minimum(ys + fy, H)
    xs = np.arange(ow) * fx
    xe = np.minimum(xs + fx, W)
    s = csum[ye][:, xe] - csum[ys][:, xe] - csum[ye][:, xs] + csum[ys][:, xs]
    cnt = (ye - ys)[:, None] * (xe - xs)[None, :]
    mult = np.zeros_like(cnt)
    for c in np.unique(cnt):
        mult[cnt == c] = int(np.float32(4294967296.0) / np.float32(256 * int(c)))
    return (((s + cnt // 2) * mult) >> 24).astype(np.uint8)


def _bicubic(x: float) -> float:
    a = -0.5
    x = abs(x)
    if x < 1.0:
        return ((a + 2.0) * x - (a + 3.0)) * x * x + 1
    if x < 2.0:
        return (((x - 5) * x + 8) * x - 4) * a
    return 0.0


def resample_coeffs(in_size: int, in0: float, in1: float, out_size: int):
    """Resample.c::precompute_coeffs + normalize_coeffs_8bpc (BICUBIC): -> [(first, count)], int coefficients (22-bit)."""
    import math
    scale = (in1 - in0) / out_size
    fs = max(scale, 1.0)
    support = 2.0 * fs
    ksize = int(math.ceil(support)) * 2 + 1
    bounds, kk = [], np.zeros((out_size, ksize), np.int64)
    for xx in range(out_size):
        center = in0 + (xx + 0.5) * scale
        xmin = max(int(center - support + 0.5), 0)
        xmax = min(int(center + support + 0.5), in_size) - xmin
        k = [_bicubic((x + xmin - center + 0.5) * (1.0 / fs)) for x in range(xmax)]
        ww = 0.0
        for w in k:
            ww += w
        for x, w in enumerate(k):
            if ww != 0.0:
                w = w / ww
            kk[xx, x] = int(-0.5 + w * (1 << 22)) if w < 0 else int(0.5 + w * (1 << 22))
        bounds.append((xmin, xmax))
    return bounds, kk


def _resample_pass(a: np.ndarray, bounds, kk, axis: int) -> np.ndarray:
    if axis == 0:
        a = a.T
    out = np.zeros((a.shape[0], len(bounds)), np.uint8)
    for o, (lo, cnt) in enumerate(bounds):
        ss = (a[:, lo:lo + cnt].astype(np.int64) * kk[o, :cnt]).sum(1) + (1 << 21)
        out[:, o] = np.clip(ss >> 22, 0, 255)
    return out.T if axis == 0 else out


def thumbnail_np(a: np.ndarray, m: int) -> np.ndarray:
    """Image.thumbnail((m, m)) of a u8 plane: [reduce()] -> horizontal pass -> vertical pass."""
    H, W = a.shape
    ow, oh = thumbnail_size(W, H, m)
    if (ow, oh) == (W, H):
        return a
    fx, fy = int(W / ow / 2.0) or 1, int(H / oh / 2.0) or 1
    bx1, by1 = float(W), float(H)
    if fx > 1 or fy > 1:
        a = reduce_np(a, fx, fy)                         # _get_safe_box of the full-image box is the full image
        bx1, by1 = float(np.float32(W / fx)), float(np.float32(H / fy))
    h2, w2 = a.shape
    if ow != w2 or bx1 != ow:
        a = _resample_pass(a, *resample_coeffs(w2, 0.0, bx1, ow), 1)
    if oh != h2 or by1 != oh:
        a = _resample_pass(a, *resample_coeffs(h2, 0.0, by1, oh), 0)
    return a
