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
