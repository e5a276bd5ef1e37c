"""GPU: stage-1 kernels through the C ABI vs the NumPy oracle (bit-exact) and the reference fixtures."""
import glob
import os

import numpy as np
import pytest

from bbocr_b200 import synth
from bbocr_b200.preprocess import ImagePreprocessor, preprocess_array, pp_params, CURRENT, LEGACY
from oracle import preprocess_np as P

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(__file__), "golden")


def rnd(seed, h, w, c=None):
    rng = np.random.default_rng(seed)
    return rng.integers(0, 256, (h, w) if c is None else (h, w, c)).astype(np.uint8)


SIZES = [(1, 1), (7, 5), (16, 16), (97, 131), (240, 320), (203, 517), (600, 801)]


@pytest.mark.parametrize("hw", SIZES)
def test_steps_bit_exact(handle, hw):
    h, w = hw
    g = rnd(h * 7 + w, h, w)
    bgr = rnd(h * 11 + w, h, w, 3)
    assert np.array_equal(handle.pp_gray(bgr), P.bgr2gray(bgr))
    for s in (3.0, 5.0):
        assert np.array_equal(handle.pp_gaussian3(g, s), P.gaussian_blur3(g, s))
    for f in (1.9, 1.3, 0.5):
        assert np.array_equal(handle.pp_contrast(g, f), P.pil_contrast(g, f))
    assert np.array_equal(handle.pp_brightness(g, 1.2), P.pil_brightness(g, 1.2))
    if h >= 8 and w >= 8:
        for cl in (2.0, 2.5):
            assert np.array_equal(handle.pp_clahe(g, cl), P.clahe(g, cl))
    for pc in (20, 30):
        assert np.array_equal(handle.pp_unsharp(g, pc, 3), P.pil_unsharp(g, 1.0, pc, 3))
    dh, dw = int(h * 1.5), int(w * 1.5)
    for mode, name in ((0, "T1"), (1, "T2")):
        assert np.array_equal(handle.pp_resize_cubic(g, dh, dw, mode), P.resize_cubic(g, dw, dh, name))
    if h >= 2 and w >= 2:
        assert np.array_equal(handle.pp_adaptive_threshold(g, 1, False, 11, 2.0), P.adaptive_threshold(g, 255, "gaussian", False, 11, 2))
        assert np.array_equal(handle.pp_adaptive_threshold(g, 0, True, 35, 10.0), P.adaptive_threshold(g, 255, "mean", True, 35, 10))
        assert np.array_equal(handle.pp_adaptive_threshold(g, 1, True, 31, 5.0), P.adaptive_threshold(g, 255, "gaussian", True, 31, 5))


@pytest.mark.parametrize("path", sorted(glob.glob(os.path.join(GOLD, "preprocess_*.npz"))))
def test_chain_matches_reference_fixtures(handle, path):
    z = np.load(path)
    for name, cfg in (("current", CURRENT), ("legacy", LEGACY)):
        got = preprocess_array(z["bgr"], cfg, 0)
        assert np.array_equal(got, z[f"ref_{name}_ippoff"]), (path, name)     # the reference's own output, cv2.ipp off
        got2 = preprocess_array(z["bgr"], cfg, 1)
        d = np.abs(got2.astype(int) - z[f"ref_{name}_ippon"].astype(int))
        assert (d > 0).mean() < 2e-3 and d.max() <= 8


def test_fluent_interface_mirrors_reference(handle):
    bgr = synth.phone_photo(3003, 403, 302)
    pp = ImagePreprocessor().load_array(bgr)
    pp.to_grayscale().resize(scale_factor=1.5).denoise(strength=3).increase_contrast(1.9).increase_brightness(1.2)
    pp.clahe(clip_limit=2.5).sharpen(amount=0.3)
    want, stages = P.preprocess_chain(bgr, P.CURRENT, "T1", return_stages=True)
    assert np.array_equal(pp.get_image(), want)
    assert pp.get_steps_applied() == P.steps_list(P.CURRENT)
    assert np.array_equal(preprocess_array(bgr, CURRENT, 0), want)            # fused chain == step-by-step
    with pytest.raises(ValueError):
        ImagePreprocessor().to_grayscale()
    with pytest.raises(ValueError):
        ImagePreprocessor().load_image("/nonexistent/file.png")


def test_full_size_phone_photo(handle):
    """BASELINE config 3 size (4032x3024): bit-exact against the oracle on one photo, plus size-independent properties."""
    bgr = synth.phone_photo(3001)
    got = preprocess_array(bgr, CURRENT, 0)
    assert got.shape == (4536, 6048)
    want = P.preprocess_chain(bgr, P.CURRENT, "T1")
    assert np.array_equal(got, want)
    # idempotence of the tone LUT on a constant image and determinism of repeated launches
    assert np.array_equal(preprocess_array(bgr, CURRENT, 0), got)
    flat = np.full((3024, 4032, 3), 127, np.uint8)
    out = preprocess_array(flat, CURRENT, 0)
    assert out.min() == out.max()


def test_deskew_matches_oracle_and_recovers_angle(handle):
    """Deskew is defined by this repository (not in the reference): CUDA == NumPy definition, and a page rotated by a
    known angle is brought back."""
    import cv2
    page = cv2.cvtColor(synth.title_page(5, 640, 480), cv2.COLOR_RGB2GRAY)
    for ang in (-3.0, 2.2, 0.7):
        M = cv2.getRotationMatrix2D((319.5, 239.5), -ang, 1.0)
        rot = cv2.warpAffine(page, M, (640, 480), flags=cv2.INTER_LINEAR, borderMode=cv2.BORDER_REPLICATE)
        want, wa = P.deskew(rot, 5.0)
        got, ga = handle.pp_deskew(rot, 5.0)
        assert abs(ga - float(wa)) < 1e-6 and abs(ga - ang) <= 0.15
        assert np.array_equal(got, want)


def test_config3_scan_chain(handle):
    """BASELINE config[2] (gray, CLAHE, adaptive threshold, deskew) as one chain: bit-exact against the composed oracles on a
    reduced photo, and at the full 4032x3024 size determinism, a binary result and agreement of the estimated angle."""
    import cv2
    bgr = synth.phone_photo(3005, 1008, 756)
    M = cv2.getRotationMatrix2D((503.5, 377.5), 1.8, 1.0)
    bgr = cv2.warpAffine(bgr, M, (1008, 756), flags=cv2.INTER_LINEAR, borderMode=cv2.BORDER_REPLICATE)
    got, ang = handle.preprocess_scan(bgr, 2.0, 11, 2.0, 5.0)
    eq = P.clahe(P.bgr2gray(bgr), 2.0)
    rot, want_ang = P.deskew(eq, 5.0)
    want = P.adaptive_threshold(rot, 255, "gaussian", False, 11, 2)
    assert abs(ang - float(want_ang)) < 1e-6
    assert np.array_equal(got, want)
    big = synth.phone_photo(3001)
    a, ang_a = handle.preprocess_scan(big)
    b, ang_b = handle.preprocess_scan(big)
    assert a.shape == (3024, 4032) and set(np.unique(a).tolist()) <= {0, 255} and np.array_equal(a, b) and ang_a == ang_b
    assert abs(ang_a) <= 5.0


@pytest.mark.parametrize("mode", [0, 1])
def test_chain_fuzz_odd_sizes(handle, mode):
    """The fused chain kernels on ragged geometries (widths not divisible by 4 or 8, tiles with one row / column, images
    smaller than a tile): bit-exact against the oracle, and the batched entry gives the same bytes as the single one."""
    rng = np.random.default_rng(90 + mode)
    sizes = [(9, 9), (11, 86), (86, 11), (22, 171), (171, 22), (86, 87), (129, 257), (341, 343)] + \
            [(int(rng.integers(8, 300)), int(rng.integers(8, 400))) for _ in range(10)]
    for (hh, ww) in sizes:
        bgr = rng.integers(0, 256, (hh, ww, 3), dtype=np.uint8)
        if hh > 40:
            bgr = synth.phone_photo(hh * 1000 + ww, ww, hh)
        for cfg, pcfg in ((CURRENT, P.CURRENT), (LEGACY, P.LEGACY)):
            got = preprocess_array(bgr, cfg, mode)
            want = P.preprocess_chain(bgr, pcfg, "T1" if mode == 0 else "T2")
            assert got.shape == want.shape and np.array_equal(got, want), (hh, ww, mode)
    same = [rng.integers(0, 256, (37, 53, 3), dtype=np.uint8) for _ in range(11)]
    outs = handle.preprocess_batch(same, pp_params(CURRENT, mode))
    for im, o in zip(same, outs):
        assert np.array_equal(o, preprocess_array(im, CURRENT, mode))


def test_deskew_histogram_kernels_agree(handle):
    """The three projection-profile kernels (per-pixel atomics = the definition, banded, run-based) give identical scores
    for every angle, including steep ranges where consecutive even columns skip bins, and equal the NumPy definition."""
    import cv2
    import math
    rng = np.random.default_rng(12)
    imgs = [cv2.cvtColor(synth.title_page(6, 801, 603), cv2.COLOR_RGB2GRAY), cv2.cvtColor(synth.phone_photo(7, 1000, 750), cv2.COLOR_BGR2GRAY),
            rng.integers(0, 256, (97, 131), dtype=np.uint8), rng.integers(0, 256, (64, 1), dtype=np.uint8)]
    for g in imgs:
        for max_deg in (5.0, 40.0, 0.0):
            ref = handle.dbg_deskew_scores(g, max_deg, 0)
            assert len(ref) == 2 * int(round(max_deg / 0.1)) + 1
            for variant in (1, 2):
                assert np.array_equal(handle.dbg_deskew_scores(g, max_deg, variant), ref), (g.shape, max_deg, variant)
    g = imgs[0]
    H, W = g.shape
    fg = P.adaptive_threshold(g, 255, "gaussian", True, 31, 5)
    ys, xs = np.nonzero(fg[:, ::2])
    dx, dy = xs.astype(np.float64) * 2 - (W - 1) * 0.5, ys.astype(np.float64) - (H - 1) * 0.5
    NR = int(math.ceil(math.sqrt(float(H) * H + float(W) * W))) + 3
    got = handle.dbg_deskew_scores(g, 5.0, 2)
    for i in (0, 17, 50, 100):
        rad = float(i - 50) * 0.1 * 3.141592653589793 / 180.0
        r = np.clip(np.floor(dy * math.cos(rad) - dx * math.sin(rad)).astype(np.int64) + NR // 2, 0, NR - 1)
        cnt = np.bincount(r, minlength=NR).astype(np.uint64)
        assert int((cnt * cnt).sum()) == int(got[i])


def test_equalize_histogram_bit_exact(handle):
    """ImagePreprocessor.equalize_histogram (image_preprocessor.py:39-46) = cv2.equalizeHist."""
    import cv2
    from bbocr_b200.preprocess import ImagePreprocessor
    rng = np.random.default_rng(9)
    big = cv2.cvtColor(synth.phone_photo(3007, 4032, 3024), cv2.COLOR_BGR2GRAY)          # 12 MP: count sums beyond 2^24 are rare but legal
    for g in [rng.integers(0, 256, (97, 131), dtype=np.uint8), rng.integers(40, 90, (300, 200), dtype=np.uint8),
              np.full((20, 30), 77, np.uint8), big]:
        assert np.array_equal(handle.pp_equalize_hist(g), cv2.equalizeHist(g))
        assert np.array_equal(handle.pp_equalize_hist(g), P.equalize_hist(g))
    pp = ImagePreprocessor().load_array(synth.phone_photo(3008, 640, 480)).equalize_histogram()
    assert pp.get_steps_applied() == ["original", "grayscale", "equalize_histogram"]
    assert np.array_equal(pp.get_image(), cv2.equalizeHist(cv2.cvtColor(synth.phone_photo(3008, 640, 480), cv2.COLOR_BGR2GRAY)))


def test_input_validation_raises_instead_of_reading_out_of_bounds(handle):
    """a 2-D array reaching bbocr_pp_gray was read as HxWx3; non-uint8 arrays were cast silently."""
    from bbocr_b200.preprocess import ImagePreprocessor
    gray = np.zeros((40, 50), np.uint8)
    with pytest.raises(ValueError):
        handle.pp_gray(gray)
    with pytest.raises(ValueError):
        ImagePreprocessor().load_array(gray).to_grayscale()
    with pytest.raises(ValueError):
        handle.pp_gaussian3(np.zeros((40, 50), np.float32), 3.0)
