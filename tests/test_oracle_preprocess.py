"""CPU: pin the NumPy preprocessing oracle against cv2 / Pillow and the reference's golden outputs."""
import glob
import os

import cv2
import numpy as np
import pytest
from PIL import Image, ImageEnhance, ImageFilter

from oracle import preprocess_np as P

GOLD = os.path.join(os.path.dirname(__file__), "golden")


def smooth(rng, h, w):
    a = rng.integers(0, 256, (h // 8 + 2, w // 8 + 2)).astype(np.float32)
    a = cv2.resize(a, (w, h), interpolation=cv2.INTER_CUBIC) + rng.normal(0, 12, (h, w))
    return np.clip(a, 0, 255).astype(np.uint8)


@pytest.mark.parametrize("hw", [(97, 131), (240, 320), (203, 517), (64, 64)])
def test_steps_match_cv2_and_pillow(hw):
    rng = np.random.default_rng(hw[0] * 1000 + hw[1])
    h, w = hw
    bgr = np.stack([smooth(rng, h, w) for _ in range(3)], -1)
    g = smooth(rng, h, w)
    pil = Image.fromarray(g)
    assert np.array_equal(P.bgr2gray(bgr), cv2.cvtColor(bgr, cv2.COLOR_BGR2GRAY))
    ipp = cv2.ipp.useIPP()
    try:
        cv2.ipp.setUseIPP(False)
        ref = cv2.resize(g, (int(w * 1.5), int(h * 1.5)), interpolation=cv2.INTER_CUBIC)
    finally:
        cv2.ipp.setUseIPP(ipp)
    assert np.array_equal(P.resize_scale(g, 1.5, "T1"), ref)
    for s in (3, 5):
        assert np.array_equal(P.gaussian_blur3(g, s), cv2.GaussianBlur(g, (3, 3), s))
    for f in (1.9, 1.3):
        assert np.array_equal(P.pil_contrast(g, f), np.array(ImageEnhance.Contrast(pil).enhance(f)))
    assert np.array_equal(P.pil_brightness(g, 1.2), np.array(ImageEnhance.Brightness(pil).enhance(1.2)))
    for cl in (2.0, 2.5):
        assert np.array_equal(P.clahe(g, cl), cv2.createCLAHE(clipLimit=cl, tileGridSize=(8, 8)).apply(g))
    for pc in (20, 30):
        assert np.array_equal(P.pil_unsharp(g, 1.0, pc, 3),
                              np.array(pil.filter(ImageFilter.UnsharpMask(radius=1.0, percent=pc, threshold=3))))
    # GAUSSIAN_C: cv2's float filter rounds its scalar tail columns differently -> pixel-count tolerance
    d = P.adaptive_threshold(g, 255, "gaussian", False, 11, 2) != \
        cv2.adaptiveThreshold(g, 255, cv2.ADAPTIVE_THRESH_GAUSSIAN_C, cv2.THRESH_BINARY, 11, 2)
    assert d.mean() < 2e-4
    assert np.array_equal(P.adaptive_threshold(g, 255, "mean", True, 35, 10),
                          cv2.adaptiveThreshold(g, 255, cv2.ADAPTIVE_THRESH_MEAN_C, cv2.THRESH_BINARY_INV, 35, 10))
    d = P.adaptive_threshold(g, 255, "gaussian", True, 31, 5) != \
        cv2.adaptiveThreshold(g, 255, cv2.ADAPTIVE_THRESH_GAUSSIAN_C, cv2.THRESH_BINARY_INV, 31, 5)
    assert d.mean() < 2e-4


@pytest.mark.parametrize("path", sorted(glob.glob(os.path.join(GOLD, "preprocess_*.npz"))))
def test_chain_matches_reference_fixtures(path):
    """Fixtures = the reference's own preprocess_for_book_cover (current + legacy constants) run by make_golden.py."""
    z = np.load(path)
    for name, params in (("current", P.CURRENT), ("legacy", P.LEGACY)):
        t1 = P.preprocess_chain(z["bgr"], params, "T1")
        assert np.array_equal(t1, z[f"ref_{name}_ippoff"]), (path, name)
        # IPP-enabled x86 wheels: real-arithmetic cubic; tolerance-checked (pixel count and amplitude)
        t2 = P.preprocess_chain(z["bgr"], params, "T2")
        ref = z[f"ref_{name}_ippon"]
        d = np.abs(t2.astype(int) - ref.astype(int))
        assert (d > 0).mean() < 2e-3 and d.max() <= 8, (path, name, (d > 0).mean(), d.max())
    if "golden_legacy" in z.files:
        # the PNG recorded by the reference author reproduces with the IPP-on reference to <= 4 px
        assert (z["golden_legacy"] != z["ref_legacy_ippon"]).sum() <= 4
        d = np.abs(P.preprocess_chain(z["bgr"], P.LEGACY, "T2").astype(int) - z["golden_legacy"].astype(int))
        assert (d > 0).mean() < 2e-3 and d.max() <= 8


def test_step_strings():
    assert P.steps_list(P.CURRENT) == ["original", "grayscale", "resize(scale_factor=1.5)", "denoise(strength=3)",
                                       "increase_contrast(factor=1.9)", "increase_brightness(factor=1.2)",
                                       "clahe(clip_limit=2.5)", "sharpen(amount=0.3)"]


def test_pillow_box_pass_closed_form():
    """Closed form behind a cheaper device box pass: for UnsharpMask(radius=1) Pillow's pass
    (c*ww + (l+r)*fw + 2^23) >> 24 with ww = 11184811 = 4 fw + 3 and 6 fw = 2^24 - 4 is round-half-up of (4c + l + r) / 6 except
    that the exact ties t = 6k + 3 round DOWN unless 3c >= 4k + 2 (the weights sit a hair below 2/3 and 1/6); and
    x // 6 == (x * 10923) >> 16 on the range, i.e. 11-bit arithmetic."""
    edge_a, ww, fw = P._pil_box_params(1.0)
    assert (edge_a, ww, fw) == (0, 11184811, 2796202) and ww == 4 * fw + 3 and 6 * fw == (1 << 24) - 4
    c = np.arange(256, dtype=np.int64)[:, None]
    s = np.arange(511, dtype=np.int64)[None, :]
    pillow = (c * ww + s * fw + (1 << 23)) >> 24
    t = 4 * c + s
    tie_down = ((t % 6) == 3) & (3 * c < 4 * ((t - 3) // 6) + 2)
    assert np.array_equal(pillow, (t + 3) // 6 - tie_down)
    assert np.array_equal((t + 3) // 6, ((t + 3) * 10923) >> 16)
    assert int(tie_down.sum()) > 0                              # the plain (t + 3) // 6 is NOT Pillow: 14.5 % of the triples differ


def test_library_call_chain_equals_the_numpy_oracle():
    """oracle/preprocess_cv.py (the reference's own cv2 + Pillow calls, used to time the CPU arm of stage 1) gives the bytes
    of the NumPy restatement when OpenCV's IPP paths are off (mode T1), and stays within the documented T1/T2 gap with IPP on."""
    from oracle import preprocess_cv as CV
    from bbocr_b200 import synth
    bgr = synth.phone_photo(3009, 640, 480)
    want = P.preprocess_chain(bgr, P.CURRENT, "T1")
    try:
        cv2.ipp.setUseIPP(False)
        assert np.array_equal(CV.preprocess_for_book_cover_cv(bgr), want)
    finally:
        cv2.ipp.setUseIPP(True)
    got = CV.preprocess_for_book_cover_cv(bgr)
    assert got.shape == want.shape and np.abs(got.astype(int) - want.astype(int)).max() <= 12


def test_equalize_hist_matches_cv2():
    import cv2
    from bbocr_b200 import synth
    rng = np.random.default_rng(9)
    imgs = [rng.integers(0, 256, (97, 131), dtype=np.uint8), rng.integers(40, 90, (300, 200), dtype=np.uint8),
            np.full((20, 30), 77, np.uint8), cv2.cvtColor(synth.phone_photo(3007, 1008, 756), cv2.COLOR_BGR2GRAY),
            (rng.random((512, 512)) ** 3 * 255).astype(np.uint8)]
    for g in imgs:
        assert np.array_equal(P.equalize_hist(g), cv2.equalizeHist(g))
