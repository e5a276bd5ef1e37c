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


def _shape_maps(seed, h, w, n_shapes):
    """Score maps full of rotated rectangles, ellipses, thin strokes and touching blobs: many hull / caliper configurations."""
    rng = np.random.default_rng(seed)
    text = np.zeros((h, w), np.float32)
    link = np.zeros((h, w), np.float32)
    for _ in range(n_shapes):
        cx, cy = int(rng.integers(0, w)), int(rng.integers(0, h))
        kind = int(rng.integers(0, 4))
        val = float(rng.uniform(0.5, 0.99))
        if kind == 0:
            box = cv2.boxPoints(((cx, cy), (float(rng.uniform(3, 120)), float(rng.uniform(2, 40))), float(rng.uniform(-90, 90))))
            cv2.fillConvexPoly(text, box.astype(np.int32), val)
        elif kind == 1:
            cv2.ellipse(text, (cx, cy), (int(rng.integers(2, 60)), int(rng.integers(2, 25))), float(rng.uniform(0, 180)), 0, 360, val, -1)
        elif kind == 2:
            cv2.line(text, (cx, cy), (cx + int(rng.integers(-80, 80)), cy + int(rng.integers(-30, 30))), val, int(rng.integers(1, 4)))
        else:
            cv2.circle(text, (cx, cy), int(rng.integers(1, 12)), val, -1)
            cv2.line(link, (cx, cy), (cx + int(rng.integers(5, 60)), cy), float(rng.uniform(0.45, 0.9)), int(rng.integers(1, 5)))
    text += rng.normal(0, 0.004, text.shape).astype(np.float32)
    return text, link


@pytest.mark.parametrize("case", [(31, 480, 640, 90), (32, 720, 960, 220), (33, 333, 517, 60), (34, 1280, 960, 400), (35, 64, 96, 4)] +
                                 [(40 + i, 480, 640, 90) for i in range(24)])
def test_device_min_area_boxes_equal_host_tail_and_cv2(handle, case):
    """getDetBoxes_core with hull + rotating calipers ON THE DEVICE (k_det_boxes, geom.cuh) vs the round-1 host tail (boxes.cpp,
    same geom.cuh code on x86-64) vs cv2 itself: bitwise identical boxes, thousands of components of every shape."""
    seed, h, w, n = case
    t, l = _shape_maps(seed, h, w, n)
    dev, used = handle.det_boxes(t, l, 0.7, 0.4, 0.4, host_path=False)
    host, used_h = handle.det_boxes(t, l, 0.7, 0.4, 0.4, host_path=True)
    assert used and not used_h, "the device tail must be the one that ran"
    want, _, _ = E.get_det_boxes_core(t, l, 0.7, 0.4, 0.4)
    assert len(dev) == len(host) == len(want) and len(want) > 0
    assert np.array_equal(dev.view(np.uint32), host.view(np.uint32))
    for a, b in zip(dev, want):
        assert np.array_equal(a.view(np.uint32), np.asarray(b, np.float32).view(np.uint32))
    print(f"{len(want)} boxes identical (device == host == cv2)")


def test_device_box_tail_overflow_falls_back_to_host(handle):
    """A component taller than the per-block point budget (> 1024 score-map rows) is flagged and the page takes the host tail."""
    t = np.zeros((1300, 200), np.float32)
    l = np.zeros((1300, 200), np.float32)
    t[20:1280, 90:110] = 0.9
    boxes, used = handle.det_boxes(t, l, 0.7, 0.4, 0.4, host_path=False)
    want, _, _ = E.get_det_boxes_core(t, l, 0.7, 0.4, 0.4)
    assert not used and len(boxes) == len(want) == 1
    assert np.array_equal(boxes[0].view(np.uint32), np.asarray(want[0], np.float32).view(np.uint32))


def test_fused_stem_equals_two_kernel_path(tmp_path):
    """k_stem_norm + k_conv_stem (bf16x3: conv1_1 with its A tiles gathered in shared memory by producer warps and a TMA-store
    epilogue) against k_im2col_rgb_split + k_conv_tc on the same pages (odd sizes, a resized canvas, a batch): score maps and
    batched readtext results bitwise equal.  The switch is read once per process, hence the two sub-processes
    (tools/stem_ab.py)."""
    import os
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    files = []
    for fused in ("1", "0"):
        f = str(tmp_path / f"stem{fused}.npz")
        env = dict(os.environ, BBOCR_STEM_FUSED=fused)
        subprocess.run([sys.executable, os.path.join(root, "tools", "stem_ab.py"), f], check=True, env=env, timeout=600)
        files.append(np.load(f))
    a, b = files
    assert set(a.files) == set(b.files) and len(a.files) >= 11
    for k in a.files:
        assert np.array_equal(a[k], b[k]), k
