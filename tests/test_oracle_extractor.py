"""CPU: the extractor-glue oracle against Pillow itself (enhanced_extractor.py:486-512)."""
import numpy as np
from PIL import Image

from oracle import extractor_ref as X


def test_thumbnail_size_rule_matches_pillow():
    rng = np.random.default_rng(0)
    cases = [(6048, 4536, 1600), (4536, 6048, 2400), (1920, 1440, 1600), (1601, 37, 1600), (3000, 2999, 1600), (2401, 2400, 2400)]
    cases += [(int(rng.integers(1601, 6000)), int(rng.integers(50, 6000)), int(rng.choice([1600, 2400]))) for _ in range(40)]
    for W, H, m in cases:
        img = Image.new("L", (W, H))
        img.thumbnail((m, m))
        assert X.thumbnail_size(W, H, m) == img.size, (W, H, m)


def test_gray_thumbnail_equals_every_channel_of_the_reference_rgb_thumbnail():
    rng = np.random.default_rng(1)
    g = rng.integers(0, 256, (1500, 2100), dtype=np.uint8)
    a = X.ocr_input_image(g, 0)
    b = X.ocr_input_image(g, 0, as_reference_rgb=True)
    assert a.shape == (1143, 1600) and b.shape == (1143, 1600, 3)
    assert all(np.array_equal(a, b[:, :, c]) for c in range(3))
    small = rng.integers(0, 256, (900, 1200), dtype=np.uint8)
    assert X.ocr_input_image(small, 0) is not None and np.array_equal(X.ocr_input_image(small, 0), small)   # below the cap: untouched
    assert X.ocr_input_image(g, 3).shape == (1500, 2100)                                                    # non-cover pages: 2400 cap


def test_reduce_restatement_matches_pillow():
    """libImaging/Reduce.c for mode L: the one-expression restatement against Image.reduce for every kernel family
    (1xN, Nx1, 2x2 .. 5x5, NxN) and for partial edge blocks (ImagingReduceCorners)."""
    rng = np.random.default_rng(2)
    for H, W, fx, fy in [(64, 80, 2, 2), (65, 81, 2, 2), (67, 83, 3, 3), (66, 85, 4, 4), (71, 93, 5, 5), (71, 93, 6, 6), (71, 93, 2, 3),
                         (70, 93, 3, 2), (71, 90, 1, 2), (71, 93, 7, 5), (50, 50, 1, 3), (50, 50, 3, 1), (51, 52, 2, 1), (53, 55, 6, 1)]:
        a = rng.integers(0, 256, (H, W), dtype=np.uint8)
        assert np.array_equal(X.reduce_np(a, fx, fy), np.asarray(Image.fromarray(a).reduce((fx, fy)))), (H, W, fx, fy)


def test_thumbnail_restatement_matches_pillow_with_and_without_reduce():
    """Image.thumbnail incl. the reduce() pre-pass of shrinks >= 4x (enhanced_extractor.py:494-497 on 5712x4284 photos:
    x1.5 = 8568x6426 -> 1600 is a 5.36x shrink): scaled-down versions of the reference's photo geometries, odd sizes
    (fractional float32 boxes), factors 2..6 and the plain < 4x regime."""
    rng = np.random.default_rng(3)
    for H, W, m in [(643, 857, 160), (642, 856, 160), (481, 643, 100), (1071, 1428, 266), (300, 1000, 120), (1000, 301, 90),
                    (611, 799, 66), (500, 700, 600), (756, 1008, 400), (1607, 1205, 300)]:
        a = rng.integers(0, 256, (H, W), dtype=np.uint8)
        img = Image.fromarray(a)
        img.thumbnail((m, m))
        assert np.array_equal(X.thumbnail_np(a, m), np.asarray(img)), (H, W, m)
