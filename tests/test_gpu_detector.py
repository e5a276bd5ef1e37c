"""GPU: CRAFT score maps (FP32 <= 1e-3, stated BF16 tolerance) and getDetBoxes / grouping (bit-exact) vs the oracle."""
import cv2
import numpy as np
import pytest

from bbocr_b200 import synth, _lib
from oracle import easyocr_restated as E

pytestmark = pytest.mark.gpu

BF16_TOL = 6e-2      # stated BF16 tolerance on score maps in [0, 1.3]: bf16 operands, FP32 accumulation, ~25 layers


@pytest.mark.parametrize("page", [("title", 320, 256), ("cover", 416, 288), ("title", 333, 250)])
def test_score_maps_fp32(gpu_reader, oracle_reader, page):
    kind, w, h = page
    img = synth.title_page(11, w, h) if kind == "title" else synth.book_cover(12, w, h)
    gpu_reader.set_precision("fp32")
    t, l, ratio = gpu_reader.score_maps(img)
    ot, ol, oratio = oracle_reader.score_maps(img)
    assert t.shape == ot.shape and ratio == oratio
    assert np.abs(t - ot).max() < 1e-3 and np.abs(l - ol).max() < 1e-3


def test_score_maps_resized_canvas(gpu_reader, oracle_reader):
    img = synth.title_page(13, 700, 500)
    gpu_reader.set_precision("fp32")
    t, l, ratio = gpu_reader.score_maps(img, canvas_size=480)      # exercises the INTER_LINEAR u8 resize + padding
    ot, ol, oratio = oracle_reader.score_maps(img, canvas_size=480)
    assert t.shape == ot.shape and ratio == oratio
    assert np.abs(t - ot).max() < 1e-3 and np.abs(l - ol).max() < 1e-3


def test_score_maps_bf16_tolerance(gpu_reader, oracle_reader):
    img = synth.title_page(11, 640, 480)
    gpu_reader.set_precision("bf16")
    try:
        t, l, _ = gpu_reader.score_maps(img)
    finally:
        gpu_reader.set_precision("fp32")
    ot, ol, _ = oracle_reader.score_maps(img)
    err = max(np.abs(t - ot).max(), np.abs(l - ol).max())
    print("bf16 score-map max-abs error", err)
    assert err < BF16_TOL


def _maps(seed, w, h, cover=False):
    page, mask = (synth.book_cover if cover else synth.title_page)(seed, w, h, True)
    return synth.score_maps_for(mask, np.random.default_rng(seed))


@pytest.mark.parametrize("case", [(21, 640, 480, False), (22, 1280, 960, True), (23, 1920, 1440, False), (24, 333, 250, False)])
def test_det_boxes_bit_exact_on_synthetic_maps(handle, case):
    seed, w, h, cover = case
    t, l = _maps(seed, w, h, cover)
    want, _, _ = E.get_det_boxes_core(t, l, 0.7, 0.4, 0.4)
    got = handle.det_boxes(t, l, 0.7, 0.4, 0.4)
    assert len(got) == len(want) and len(want) > 0
    for a, b in zip(got, want):
        assert np.array_equal(a.view(np.uint32), np.asarray(b, np.float32).view(np.uint32))


def test_det_boxes_edge_cases(handle):
    z = np.zeros((64, 96), np.float32)
    assert len(handle.det_boxes(z, z)) == 0                               # empty page
    one = np.ones((64, 96), np.float32)
    want, _, _ = E.get_det_boxes_core(one, one, 0.7, 0.4, 0.4)
    got = handle.det_boxes(one, one)                                      # one component covering everything
    assert len(got) == len(want) == 1 and np.array_equal(got[0], want[0])
    rng = np.random.default_rng(5)
    t = cv2.GaussianBlur((rng.random((200, 300)) > 0.97).astype(np.float32), (0, 0), 1.5) * 6
    l = cv2.GaussianBlur((rng.random((200, 300)) > 0.98).astype(np.float32), (0, 0), 2.5) * 8
    want, _, _ = E.get_det_boxes_core(t, l, 0.7, 0.4, 0.4)                 # many small, ragged, link-only components
    got = handle.det_boxes(t, l)
    assert len(got) == len(want)
    for a, b in zip(got, want):
        assert np.array_equal(a.view(np.uint32), np.asarray(b, np.float32).view(np.uint32))


def test_detect_matches_oracle_given_same_maps(gpu_reader, oracle_reader):
    img = synth.title_page(31, 960, 704)
    ot, ol, ratio = oracle_reader.score_maps(img)
    boxes = gpu_reader.handle.det_boxes(ot, ol)
    hl, fl = _lib.group_boxes(boxes, ratio)
    oh, of = oracle_reader.boxes_from_maps(ot, ol, ratio)
    assert [list(map(int, b)) for b in oh] == hl.tolist()
    assert len(of) == len(fl) and all(np.array_equal(np.array(a, np.float64), b) for a, b in zip(of, fl))
