"""CPU: structural pins of the restated EasyOCR oracle (parity unpinned: no easyocr / checkpoints in the image)."""
import numpy as np
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
