"""CPU: structural pins of the restated EasyOCR oracle (parity unpinned: no easyocr / checkpoints in the image)."""
import numpy as np
import pytest
import torch

from bbocr_b200 import weights
from oracle import easyocr_restated as E


def test_parameter_counts_match_published_checkpoints():
    assert sum(p.numel() for p in E.CRAFT().parameters()) == 20770466          # craft_mlt_25k.pth (83.1 MB FP32)
    assert sum(p.numel() for p in E.CRNN().parameters()) == 3781345            # english_g2.pth   (15.1 MB FP32)
    assert len(E.CHARACTERS) == 96


def test_state_dicts_load_strictly():
    E.CRAFT().load_state_dict(weights.to_torch_state(weights.calibrated_craft_state()), strict=True)
    E.CRNN().load_state_dict(weights.to_torch_state(weights.calibrated_crnn_state()), strict=True)
    a, b = weights.random_craft_state(7), weights.random_craft_state(7)
    assert all(np.array_equal(a[k], b[k]) for k in a)


def test_inplace_relu_rectifies_skip_taps():
    net = E.VGG16BN().eval()
    with torch.no_grad():
        fc7, r53, r43, r32, r22 = net(torch.randn(1, 3, 64, 64))
    assert r22.min() >= 0 and r32.min() >= 0 and r43.min() >= 0       # aliased by the next slice's in-place ReLU
    assert r53.min() < 0                                              # followed by MaxPool: raw BN output


def test_ctc_greedy_known_answers():
    # blank=0; repeats collapse unless separated by a blank
    idx = np.array([0, 11, 11, 0, 11, 12, 12, 0, 0, 13])
    assert E.decode_greedy(idx, [len(idx)]) == ["".join(E.CHARACTERS[i - 1] for i in (11, 11, 12, 13))]
    assert E.custom_mean(np.array([0.5, 0.5, 0.5, 0.5], np.float32)) == 0.0625 ** (2.0 / 2.0)
    assert E.custom_mean(np.array([0])) == 0.0


def test_group_text_box_known_answer():
    polys = [np.array([10, 10, 110, 10, 110, 40, 10, 40], np.int32), np.array([120, 12, 200, 12, 200, 42, 120, 42], np.int32),
             np.array([10, 100, 90, 130, 80, 160, 0, 130], np.int32)]
    h, f = E.group_text_box(polys, 0.1, 0.5, 0.5, 0.5, 0.1, True)
    assert [list(map(int, b)) for b in h] == [[7, 203, 7, 45]]
    assert len(f) == 1


def test_rotation_info_is_rot90():
    """utils.make_rotated_img_list calls scipy.ndimage.rotate(img, angle, reshape=True); for the eligible angles the spline
    rotation is exactly np.rot90 -- the form the oracle (and the device kernel) use."""
    ndi = pytest.importorskip("scipy.ndimage")
    rng = np.random.default_rng(0)
    for trial in range(200):
        h, w = int(rng.integers(1, 80)), int(rng.integers(1, 300))
        a = rng.integers(0, 256, (h, w), dtype=np.uint8)
        for angle in (90, 180, 270):
            assert np.array_equal(ndi.rotate(a, angle, reshape=True), np.rot90(a, angle // 90))
    box = [[0, 0], [1, 0], [1, 1], [0, 1]]
    out = E.make_rotated_img_list([90, 270], [(box, a)])
    assert len(out) == 3 and np.array_equal(out[1][1], np.rot90(a, 1)) and np.array_equal(out[2][1], np.rot90(a, 3))


def test_set_result_with_confidence_takes_the_first_maximum():
    rows = [[("b0", "a", 0.5), ("b1", "x", 0.2)], [("b0", "b", 0.5), ("b1", "y", 0.9)], [("b0", "c", 0.1), ("b1", "z", 0.9)]]
    assert E.set_result_with_confidence(rows) == [("b0", "a", 0.5), ("b1", "y", 0.9)]
