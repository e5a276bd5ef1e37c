"""GPU: the page crops in front of the OCR stage (enhanced_extractor.py:239-397) through the C ABI
(bbocr_autocrop_rect, bbocr_external_boxes, bbocr_rect_morph), bit-exact against the oracle, against cv2 itself, and
against the fixtures recorded from the reference's own function."""
import glob
import os

import cv2
import numpy as np
import pytest

from bbocr_b200 import extractor, synth
from oracle import autocrop_np as A

pytestmark = pytest.mark.gpu

GOLD = sorted(glob.glob(os.path.join(os.path.dirname(__file__), "golden", "autocrop_*.npz")))


def cv_boxes(m):
    cs, _ = cv2.findContours(m, cv2.RETR_EXTERNAL, cv2.CHAIN_APPROX_SIMPLE)
    return np.array(sorted([cv2.boundingRect(c) for c in cs], key=lambda b: (b[1], b[0], b[2], b[3])), np.int64).reshape(-1, 4)


@pytest.mark.parametrize("shape", [(97, 130), (64, 64), (33, 31), (1, 70), (70, 1), (200, 333)])
def test_rect_morphology_vs_cv2(handle, shape):
    rng = np.random.default_rng(shape[0] * 1000 + shape[1])
    m = (rng.random(shape) < 0.08).astype(np.uint8) * 255
    for kw, kh in [(3, 3), (17, 5), (19, 7), (29, 9), (31, 11), (13, 5), (1, 9), (63, 1)]:
        k = cv2.getStructuringElement(cv2.MORPH_RECT, (kw, kh))
        assert np.array_equal(handle.rect_morph(m, kw, kh, False), cv2.dilate(m, k)), (kw, kh, "dilate")
        d = cv2.dilate(m, k)
        assert np.array_equal(handle.rect_morph(d, kw, kh, True), cv2.erode(d, k)), (kw, kh, "erode")


def test_external_boxes_vs_cv2_fuzz(handle):
    for s in range(24):
        r = np.random.default_rng(s)
        H, W = int(r.integers(1, 140)), int(r.integers(1, 200))
        m = (r.random((H, W)) < r.uniform(0.2, 0.75)).astype(np.uint8) * 255
        if s % 3 == 1:
            m = cv2.dilate(m, np.ones((2, 2), np.uint8))
        assert np.array_equal(handle.external_boxes(m), cv_boxes(m)), (s, H, W)
    assert handle.external_boxes(np.zeros((9, 40), np.uint8)).shape == (0, 4)
    assert handle.external_boxes(np.full((9, 40), 255, np.uint8)).tolist() == [[0, 0, 40, 9]]
    ring = np.zeros((50, 70), np.uint8)
    ring[2:48, 2:68] = 255
    ring[6:44, 6:64] = 0
    ring[20:30, 20:40] = 255                                  # nested in the hole of the ring: not external
    ring[0, 69] = 255                                         # diagonal contact only with nothing: its own component
    assert np.array_equal(handle.external_boxes(ring), cv_boxes(ring))
    assert handle.external_boxes(ring).tolist() == [[69, 0, 1, 1], [2, 2, 66, 46]]


def test_external_boxes_large_page_vs_cv2(handle):
    page, ink = synth.title_page(77, 1920, 1440, return_mask=True)
    m = cv2.dilate(ink, np.ones((3, 9), np.uint8))
    m[100:1300, 150] = 255                                    # a frame: the lines inside it stop being external
    m[100:1300, 1770] = 255
    m[100, 150:1771] = 255
    m[1300, 150:1771] = 255
    assert np.array_equal(handle.external_boxes(m), cv_boxes(m))
    assert np.array_equal(handle.external_boxes(ink), cv_boxes(ink))


@pytest.mark.parametrize("case", ["cover", "title", "sparse", "framed", "noise", "odd"])
def test_stages_bit_exact_vs_oracle(handle, case):
    rng = np.random.default_rng(3)
    bgr = {"cover": lambda: synth.book_cover(1, 640, 480), "title": lambda: synth.title_page(2, 803, 601),
           "sparse": lambda: synth.sparse_page(11, 1100, 800), "framed": lambda: synth.sparse_page(12, 900, 1200, frame=True),
           "noise": lambda: rng.integers(0, 256, (333, 517, 3), dtype=np.uint8),
           "odd": lambda: synth.sparse_page(13, 611, 397)}[case]()
    bgr = np.ascontiguousarray(bgr[:, :, :3])
    rect, dbg = handle.autocrop_rect(bgr, 16, debug=True)
    mask, st = A.text_mask(bgr, True)
    assert dbg["otsu"] == (st["t_eq"], st["t_grad"])
    assert np.array_equal(dbg["mask"], mask)
    merged = A.merged_mask(mask)
    assert np.array_equal(dbg["merged"], merged)
    boxes = A.external_boxes(merged)
    assert dbg["nboxes"] == len(boxes) and np.array_equal(dbg["boxes"], boxes)
    assert rect == A.auto_crop_rect(bgr, 16)
    assert np.array_equal(dbg["boxes"], cv_boxes(merged))


@pytest.mark.parametrize("path", GOLD, ids=[os.path.basename(p)[9:-4] for p in GOLD])
def test_against_the_reference_function(handle, path):
    z = np.load(path)
    bgr = z["bgr"]
    for m, want in zip(z["margins"], z["rects"]):
        got = handle.autocrop_rect(bgr, int(m))
        assert (got is None and want[0] < 0) or tuple(int(v) for v in want) == got, (m, want, got)
    for pc, *want in z["edge"]:
        got = extractor.central_edge_crop(bgr, float(pc))
        if want[0] < 0:
            assert got is None
        else:
            x0, y0, x1, y1 = (int(v) for v in want)
            assert np.array_equal(got, bgr[y0:y1, x0:x1])


def test_phone_photo_size_and_gray_input(handle):
    bgr = synth.phone_photo(3001)                             # 4032 x 3024
    rect, dbg = handle.autocrop_rect(bgr, 40, debug=True)
    merged = A.merged_mask(A.text_mask(bgr))
    assert np.array_equal(dbg["merged"], merged)
    boxes = A.external_boxes(merged)
    assert np.array_equal(dbg["boxes"], boxes)
    assert rect == A.crop_rect_from_boxes(boxes, bgr.shape[0], bgr.shape[1], 40)
    gray = cv2.cvtColor(synth.sparse_page(21, 900, 700), cv2.COLOR_BGR2GRAY)
    want = A.auto_crop_rect(cv2.cvtColor(gray, cv2.COLOR_GRAY2BGR), 8)
    assert want is not None and handle.autocrop_rect(gray, 8) == want


def test_extract_text_with_crops(gpu_reader):
    page = synth.sparse_page(31, 1100, 800)
    gpu_reader.set_precision("fp32")
    text, results = extractor.extract_text_with_ocr(gpu_reader, page, use_preprocessing=False, crop_for_ocr=True, crop_margin=16,
                                                    edge_crop_percent=2.0, return_results=True)
    g = cv2.cvtColor(page, cv2.COLOR_BGR2GRAY)
    x0, y0, x1, y1 = A.central_edge_crop_rect(g.shape[0], g.shape[1], 2.0)
    g = np.ascontiguousarray(g[y0:y1, x0:x1])
    r = A.auto_crop_rect(cv2.cvtColor(g, cv2.COLOR_GRAY2BGR), 16)
    assert r is not None
    g = np.ascontiguousarray(g[r[1]:r[3], r[0]:r[2]])
    want = gpu_reader.readtext(g, paragraph=False, batch_size=1, workers=0)
    assert results == want and text == " ".join(t for _, t, _ in want)
