"""CPU oracle (TEST INFRASTRUCTURE ONLY) for BB-OCR's optional page crops in front of the OCR stage.

NumPy / SciPy restatement of the OpenCV arithmetic behind
  pipeline_demo/extractor/enhanced_extractor.py:239-372   _auto_crop_text_region   (the text-region heuristic)
  pipeline_demo/extractor/enhanced_extractor.py:374-397   _central_edge_crop
Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may import this module; the product path
(bbocr_b200) never does.

Parity status: PINNED.  Every step is compared with cv2 4.13 in tests/test_oracle_autocrop.py (Gaussian 3x3 sigma=0,
Otsu on both maps, the Sobel magnitude, rectangle morphology incl. the iteration/kernel fusion used here, external
contour boxes incl. components nested in holes), and the whole function is compared with the reference's own
`_auto_crop_text_region` executed from /root/reference (tests/golden/make_golden_autocrop.py -> tests/golden/autocrop_*.npz).

Facts about cv2 that the restatement (and the device path) rely on, each checked by a test:
  * GaussianBlur((3,3), 0) on u8 = fixed-point kernel [64,128,64]/256 in both directions, one rounding (v + 2^15) >> 16.
  * THRESH_OTSU: the threshold is computed from the 256-bin histogram in double precision (`otsu_threshold`);
    BINARY is `v > t`, BINARY_INV is `v <= t`.
  * Sobel(CV_16S, ksize 3, BORDER_REFLECT_101) -> convertScaleAbs -> addWeighted(1,1) -> convertScaleAbs is
    min(255, min(255,|gx|) + min(255,|gy|)).
  * Morphology with a rectangle and the default (constant, "does not contribute") border: n iterations of a kw x kh
    rectangle equal one pass with (n(kw-1)+1) x (n(kh-1)+1), consecutive erosions / dilations compose by adding extents,
    and dilation distributes over OR, so
        merged = dilate13x5( erode19x7(dilate17x5(mask)) | erode31x11(dilate29x9(mask)) ).
  * findContours(RETR_EXTERNAL) returns one contour per 8-connected foreground component that is NOT enclosed in a hole of
    another component (background is 4-connected; the image is surrounded by a virtual background frame), and
    boundingRect of it is the component's pixel bounding box.
"""
from __future__ import annotations

import numpy as np
from scipy import ndimage

from . import preprocess_np as P


def gaussian3_sigma0(img: np.ndarray) -> np.ndarray:
    """cv2.GaussianBlur(gray, (3,3), 0)   enhanced_extractor.py:254"""
    a = img.astype(np.int64)
    p = np.pad(a, 1, mode="reflect")
    hor = p[:, :-2] * 64 + p[:, 1:-1] * 128 + p[:, 2:] * 64
    ver = hor[:-2] * 64 + hor[1:-1] * 128 + hor[2:] * 64
    return ((ver + (1 << 15)) >> 16).astype(np.uint8)


def otsu_threshold(hist: np.ndarray, npix: int) -> int:
    """cv2 getThreshVal_Otsu_8u: plain double arithmetic, first maximum of the between-class variance wins."""
    eps = float(np.finfo(np.float32).eps)
    scale = 1.0 / float(npix)
    mu = 0.0
    for i in range(256):
        mu += float(i) * float(hist[i])
    mu *= scale
    mu1 = 0.0
    q1 = 0.0
    max_sigma = 0.0
    max_val = 0
    for i in range(256):
        p_i = float(hist[i]) * scale
        mu1 *= q1
        q1 += p_i
        q2 = 1.0 - q1
        if min(q1, q2) < eps or max(q1, q2) > 1.0 - eps:
            continue
        mu1 = (mu1 + i * p_i) / q1
        mu2 = (mu - q1 * mu1) / q2
        sigma = q1 * q2 * (mu1 - mu2) * (mu1 - mu2)
        if sigma > max_sigma:
            max_sigma = sigma
            max_val = i
    return max_val


def otsu(img: np.ndarray) -> int:
    return otsu_threshold(np.bincount(img.reshape(-1), minlength=256), img.size)


def sobel_magnitude(img: np.ndarray) -> np.ndarray:
    """enhanced_extractor.py:263-265"""
    p = np.pad(img.astype(np.int32), 1, mode="reflect")
    gx = (p[:-2, 2:] + 2 * p[1:-1, 2:] + p[2:, 2:]) - (p[:-2, :-2] + 2 * p[1:-1, :-2] + p[2:, :-2])
    gy = (p[2:, :-2] + 2 * p[2:, 1:-1] + p[2:, 2:]) - (p[:-2, :-2] + 2 * p[:-2, 1:-1] + p[:-2, 2:])
    return np.minimum(255, np.minimum(255, np.abs(gx)) + np.minimum(255, np.abs(gy))).astype(np.uint8)


def _rect_filter(m: np.ndarray, kw: int, kh: int, dilate: bool) -> np.ndarray:
    """kw x kh rectangle (odd sizes, centre anchor); pixels outside the image do not contribute."""
    fill = 0 if dilate else 255
    op = np.maximum if dilate else np.minimum
    rx, ry = kw // 2, kh // 2
    H, W = m.shape
    p = np.pad(m, ((0, 0), (rx, rx)), constant_values=fill)
    hor = p[:, 0:W].copy()
    for i in range(1, kw):
        hor = op(hor, p[:, i:i + W])
    p = np.pad(hor, ((ry, ry), (0, 0)), constant_values=fill)
    ver = p[0:H].copy()
    for i in range(1, kh):
        ver = op(ver, p[i:i + H])
    return ver


def text_mask(bgr: np.ndarray, return_stages: bool = False):
    """enhanced_extractor.py:252-269: the composite text mask (255 = text cue)."""
    gray = P.bgr2gray(bgr)
    eq = P.clahe(gaussian3_sigma0(gray), 2.0)
    thr_mean = P.adaptive_threshold(eq, 255, "mean", True, 35, 10)
    thr_gaus = P.adaptive_threshold(eq, 255, "gaussian", True, 31, 5)
    t_eq = otsu(eq)
    grad = sobel_magnitude(eq)
    t_grad = otsu(grad)
    mask = thr_mean | thr_gaus | np.where(eq <= t_eq, 255, 0).astype(np.uint8) | np.where(grad > t_grad, 255, 0).astype(np.uint8)
    if return_stages:
        return mask, dict(eq=eq, grad=grad, t_eq=t_eq, t_grad=t_grad, thr_mean=thr_mean, thr_gaus=thr_gaus)
    return mask


def merged_mask(mask: np.ndarray) -> np.ndarray:
    """enhanced_extractor.py:271-284 with the iterations / consecutive passes folded into single rectangles."""
    v1 = _rect_filter(_rect_filter(mask, 17, 5, True), 19, 7, False)
    v2 = _rect_filter(_rect_filter(mask, 29, 9, True), 31, 11, False)
    return _rect_filter(v1 | v2, 13, 5, True)


def external_boxes(binary: np.ndarray) -> np.ndarray:
    """boundingRect of every RETR_EXTERNAL contour: (x, y, w, h) rows sorted by (y, x, w, h)."""
    fg = np.pad(binary != 0, 1)
    lab, n = ndimage.label(fg, structure=np.ones((3, 3), int))
    if n == 0:
        return np.zeros((0, 4), np.int64)
    bg_lab, _ = ndimage.label(~fg, structure=ndimage.generate_binary_structure(2, 1))
    outer = bg_lab == bg_lab[0, 0]
    near_outer = ndimage.binary_dilation(outer, structure=ndimage.generate_binary_structure(2, 1))
    ext = np.zeros(n + 1, bool)
    ext[np.unique(lab[near_outer & fg])] = True
    out = []
    for k, sl in enumerate(ndimage.find_objects(lab), start=1):
        if ext[k]:
            y0, y1, x0, x1 = sl[0].start - 1, sl[0].stop - 1, sl[1].start - 1, sl[1].stop - 1
            out.append((x0, y0, x1 - x0, y1 - y0))
    return np.array(sorted(out, key=lambda b: (b[1], b[0], b[2], b[3])), np.int64).reshape(-1, 4)


def crop_rect_from_boxes(boxes, h: int, w: int, margin: int):
    """enhanced_extractor.py:288-333: area filter, union, inflate-if-small, margin.  -> (x0, y0, x1, y1) or None."""
    img_area = float(h * w)
    keep = [b for b in boxes if not (float(b[2] * b[3]) < 0.0001 * img_area or float(b[2] * b[3]) > 0.10 * img_area)]
    if not keep:
        return None
    x0 = min(int(b[0]) for b in keep)
    y0 = min(int(b[1]) for b in keep)
    x1 = max(int(b[0] + b[2]) for b in keep)
    y1 = max(int(b[1] + b[3]) for b in keep)
    if float((x1 - x0) * (y1 - y0)) < 0.12 * img_area:
        pad = int(0.03 * max(w, h))
        x0, y0, x1, y1 = max(0, x0 - pad), max(0, y0 - pad), min(w, x1 + pad), min(h, y1 + pad)
    x0, y0, x1, y1 = max(0, x0 - margin), max(0, y0 - margin), min(w, x1 + margin), min(h, y1 + margin)
    if x1 <= x0 or y1 <= y0:
        return None
    return x0, y0, x1, y1


def auto_crop_rect(bgr: np.ndarray, margin: int):
    """_auto_crop_text_region up to the slice it writes: the crop rectangle (x0, y0, x1, y1), or None for "no crop"."""
    h, w = bgr.shape[:2]
    boxes = external_boxes(merged_mask(text_mask(bgr)))
    return crop_rect_from_boxes(boxes, h, w, margin)


def central_edge_crop_rect(h: int, w: int, percent: float):
    """_central_edge_crop (:374-397) -> (x0, y0, x1, y1) or None."""
    if percent <= 0.0:
        return None
    mx = int(round(w * (percent / 100.0)))
    my = int(round(h * (percent / 100.0)))
    x0, y0, x1, y1 = max(0, mx), max(0, my), min(w, w - mx), min(h, h - my)
    if x1 - x0 < max(16, w * 0.2) or y1 - y0 < max(16, h * 0.2):
        return None
    return x0, y0, x1, y1
