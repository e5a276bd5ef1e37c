"""CPU: the auto-crop oracle (oracle/autocrop_np.py) against cv2 itself and against the reference's own function
(fixtures from tests/golden/make_golden_autocrop.py).  SURVEY.md §8f-2; enhanced_extractor.py:239-397."""
import glob
import os

import cv2
import numpy as np
import pytest

from bbocr_b200 import synth
from oracle import autocrop_np as A

GOLD = sorted(glob.glob(os.path.join(os.path.dirname(__file__), "golden", "autocrop_*.npz")))


def cv_boxes(m):
    cs, _ = cv2.findContours(m, cv2.RETR_EXTERNAL, cv2.CHAIN_APPROX_SIMPLE)
    return np.array(sorted([cv2.boundingRect(c) for c in cs], key=lambda b: (b[1], b[0], b[2], b[3])), np.int64).reshape(-1, 4)


def cv_merged(mask):
    R = lambda w, h: cv2.getStructuringElement(cv2.MORPH_RECT, (w, h))      # noqa: E731

    def variant(kclose):
        c = cv2.morphologyEx(mask, cv2.MORPH_CLOSE, kclose, iterations=2)
        return cv2.dilate(cv2.morphologyEx(c, cv2.MORPH_OPEN, R(3, 3), iterations=1), R(11, 3), iterations=1)
    return variant(R(9, 3)) | variant(R(15, 5))


@pytest.mark.parametrize("ipp", [True, False])
def test_steps_match_cv2(ipp):
    cv2.ipp.setUseIPP(ipp)
    try:
        rng = np.random.default_rng(5)
        pages = [synth.book_cover(1, 640, 480), synth.title_page(2, 803, 601), synth.sparse_page(3, 500, 333, frame=True),
                 rng.integers(0, 256, (333, 517, 3), dtype=np.uint8)]
        for im in pages:
            im = np.ascontiguousarray(im[:, :, :3])
            mask, st = A.text_mask(im, True)
            eq = cv2.createCLAHE(clipLimit=2.0, tileGridSize=(8, 8)).apply(
                cv2.GaussianBlur(cv2.cvtColor(im, cv2.COLOR_BGR2GRAY), (3, 3), 0))
            assert np.array_equal(st["eq"], eq)
            t1, c = cv2.threshold(eq, 0, 255, cv2.THRESH_BINARY_INV + cv2.THRESH_OTSU)
            gx, gy = cv2.Sobel(eq, cv2.CV_16S, 1, 0, ksize=3), cv2.Sobel(eq, cv2.CV_16S, 0, 1, ksize=3)
            grad = cv2.convertScaleAbs(cv2.addWeighted(cv2.convertScaleAbs(gx), 1.0, cv2.convertScaleAbs(gy), 1.0, 0))
            assert np.array_equal(st["grad"], grad)
            t2, d = cv2.threshold(grad, 0, 255, cv2.THRESH_BINARY + cv2.THRESH_OTSU)
            assert (t1, t2) == (st["t_eq"], st["t_grad"])
            a = cv2.adaptiveThreshold(eq, 255, cv2.ADAPTIVE_THRESH_MEAN_C, cv2.THRESH_BINARY_INV, 35, 10)
            b = cv2.adaptiveThreshold(eq, 255, cv2.ADAPTIVE_THRESH_GAUSSIAN_C, cv2.THRESH_BINARY_INV, 31, 5)
            assert np.array_equal(mask, a | b | c | d)
            merged = A.merged_mask(mask)
            assert np.array_equal(merged, cv_merged(mask))
            assert np.array_equal(A.external_boxes(merged), cv_boxes(merged))
            assert np.array_equal(A.external_boxes(mask), cv_boxes(mask))
    finally:
        cv2.ipp.setUseIPP(True)


def test_otsu_and_nested_components_fuzz():
    for s in range(30):
        r = np.random.default_rng(s)
        m = (r.random((90, 130)) < r.uniform(0.3, 0.7)).astype(np.uint8) * 255
        if s % 2:
            m = cv2.dilate(m, np.ones((2, 2), np.uint8))
        assert np.array_equal(A.external_boxes(m), cv_boxes(m)), s
        e = r.uniform(0.5, 2.0)
        g = (r.integers(0, 256, (50, 70)).astype(np.float64) ** e / 255.0 ** (e - 1)).clip(0, 255).astype(np.uint8)
        assert cv2.threshold(g, 0, 255, cv2.THRESH_OTSU)[0] == A.otsu(g), s
    assert A.external_boxes(np.zeros((8, 8), np.uint8)).shape == (0, 4)
    ring = np.zeros((20, 20), np.uint8)
    ring[2:18, 2:18] = 255
    ring[4:16, 4:16] = 0
    ring[8:12, 8:12] = 255                                    # a component inside the hole of another one: not external
    assert A.external_boxes(ring).tolist() == [[2, 2, 16, 16]]
    assert cv_boxes(ring).tolist() == [[2, 2, 16, 16]]


@pytest.mark.parametrize("path", GOLD, ids=[os.path.basename(p)[9:-4] for p in GOLD])
def test_against_the_reference_function(path):
    z = np.load(path)
    bgr = z["bgr"]
    for m, want in zip(z["margins"], z["rects"]):
        got = A.auto_crop_rect(bgr, int(m))
        assert (got is None and want[0] < 0) or tuple(int(v) for v in want) == got
    for pc, *want in z["edge"]:
        got = A.central_edge_crop_rect(bgr.shape[0], bgr.shape[1], float(pc))
        assert (got is None and want[0] < 0) or tuple(int(v) for v in want) == got


def test_fixtures_exist():
    assert len(GOLD) >= 8
