"""GPU: the in-memory extractor glue (bbocr_b200/extractor.py, bbocr_thumbnail_u8) against Pillow, bit-exact."""
import numpy as np
import pytest

from bbocr_b200 import extractor, synth
from oracle import extractor_ref as X
from oracle import preprocess_np as P

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("shape_m", [((1500, 2100), 1600), ((2100, 1500), 1600), ((2268, 3024), 1600), ((3024, 2268), 2400),
                                     ((1601, 64), 1600), ((2500, 2500), 2400), ((1700, 1699), 1600)])
def test_thumbnail_bit_exact_vs_pillow(handle, shape_m):
    (H, W), m = shape_m
    rng = np.random.default_rng(H * 7 + W)
    g = synth.phone_photo(H + W, W, H)[:, :, 1].copy() if min(H, W) > 200 else rng.integers(0, 256, (H, W), dtype=np.uint8)
    from PIL import Image
    img = Image.fromarray(g)
    img.thumbnail((m, m))
    want = np.asarray(img)
    got = handle.thumbnail(g, m)
    assert got.shape == want.shape
    assert np.array_equal(got, want)


@pytest.mark.parametrize("shape_m", [((6426, 8568), 1600),      # the reference's 5712x4284 photos after x1.5: 5.36x -> reduce(2)
                                     ((8568, 6426), 1600),      # portrait
                                     ((5184, 6912), 1600),      # 4608x3456 (16 MP) after x1.5
                                     ((2143, 2857), 533),       # odd sizes: fractional float32 box after reduce(2)
                                     ((1443, 1929), 300),       # factor 3 (multiplier kernel), partial edge blocks
                                     ((400, 6500), 1600),       # factor 2 x 2 on a strip
                                     ((3001, 1001), 150)])      # factor 10 / 10 (NxN kernel)
def test_thumbnail_reduce_regime_bit_exact_vs_pillow(handle, shape_m):
    """Shrinks >= 4x: Pillow's Image.reduce() pre-pass + bicubic over the fractional box (enhanced_extractor.py:494-497 on the
    reference's own 5712x4284 cover photos).  Round 1 refused these sizes."""
    from PIL import Image
    (H, W), m = shape_m
    rng = np.random.default_rng(H + 3 * W)
    g = rng.integers(0, 256, (H, W), dtype=np.uint8)
    t = synth.title_page(H, W // 4, H // 4)[:, :, 0]                  # some structure besides the noise
    g[H // 4: H // 4 + t.shape[0], W // 4: W // 4 + t.shape[1]] = t
    img = Image.fromarray(g)
    img.thumbnail((m, m))
    want = np.asarray(img)
    got = handle.thumbnail(g, m)
    assert got.shape == want.shape
    assert np.array_equal(got, want)
    assert np.array_equal(X.thumbnail_np(g, m), want) if H * W < 4_000_000 else True     # the NumPy restatement agrees too


def test_extract_text_with_ocr_in_memory(gpu_reader, oracle_reader):
    """preprocess -> cap -> readtext -> join on the device == the same chain with the oracles for the first two steps."""
    bgr = synth.phone_photo(3100, 1400, 1050)                 # x1.5 -> 2100x1575 > 1600: the cap applies
    gpu_reader.set_precision("fp32")
    text, results = extractor.extract_text_with_ocr(gpu_reader, bgr, image_index=0, return_results=True)
    pre = P.preprocess_chain(bgr, P.CURRENT, "T1")
    cap = X.ocr_input_image(pre, 0)
    assert cap.shape[1] == 1600
    assert np.array_equal(extractor.ocr_input_image(gpu_reader, pre, 0), cap)
    want = gpu_reader.readtext(cap, paragraph=False, batch_size=1, workers=0)
    assert results == want and text == " ".join(r[1] for r in want)
    # host-buffer glue == device-resident glue, with and without the optional crops
    for kw in ({}, {"edge_crop_percent": 3.0}, {"crop_for_ocr": True, "crop_margin": 16}, {"edge_crop_percent": 2.0, "crop_for_ocr": True}):
        a = extractor.extract_text_with_ocr(gpu_reader, bgr, image_index=0, return_results=True, device_resident=False, **kw)
        b = extractor.extract_text_with_ocr(gpu_reader, bgr, image_index=0, return_results=True, device_resident=True, **kw)
        assert a == b and len(a[1]) > 0, kw
    big = synth.phone_photo(3002)                            # 4032x3024 -> x1.5 -> cap 1600: the extractor's real geometry
    a = extractor.extract_text_with_ocr(gpu_reader, big, image_index=0, return_results=True, device_resident=False)
    b = extractor.extract_text_with_ocr(gpu_reader, big, image_index=0, return_results=True, device_resident=True)
    assert a == b
    assert extractor.extract_text_with_ocr(gpu_reader, "/nonexistent.png") == ""      # errors become "" (:529-531)


def test_extract_without_preprocessing_feeds_the_colour_page(gpu_reader):
    """use_preprocessing=False (enhanced_extractor.py:446-447, :520): the ORIGINAL colour page reaches readtext -- detector
    input in RGB order (upstream decodes the file with skimage), crops from its BGR2GRAY plane; a preprocessing failure
    falls back to the same path (:441-443)."""
    import cv2
    bgr = synth.book_cover(77, 1100, 800)[:, :, ::-1].copy()          # a BGR "file"
    gpu_reader.set_precision("fp32")
    text, results = extractor.extract_text_with_ocr(gpu_reader, bgr, use_preprocessing=False, image_index=0, return_results=True)
    want = gpu_reader.readtext_pair(cv2.cvtColor(bgr, cv2.COLOR_BGR2RGB), cv2.cvtColor(bgr, cv2.COLOR_BGR2GRAY))
    assert len(want) > 0 and results == want and text == " ".join(r[1] for r in want)
    gray_only = gpu_reader.readtext(cv2.cvtColor(bgr, cv2.COLOR_BGR2GRAY))
    assert [r[0] for r in gray_only] != [r[0] for r in want] or [r[2] for r in gray_only] != [r[2] for r in want]   # colour matters
    # above the cap: every channel is thumbnailed like Pillow's RGB thumbnail
    from PIL import Image
    big = synth.book_cover(78, 2000, 1500)[:, :, ::-1].copy()
    cap = extractor.ocr_input_color(gpu_reader, big, 0)
    ref = Image.fromarray(big[:, :, ::-1].copy())
    ref.thumbnail((1600, 1600))
    assert np.array_equal(cap[:, :, ::-1], np.asarray(ref))
    # preprocessing raises (gray input) -> the original image goes on instead of ""
    g = cv2.cvtColor(bgr, cv2.COLOR_BGR2GRAY)
    t2, r2 = extractor.extract_text_with_ocr(gpu_reader, g, use_preprocessing=True, image_index=0, return_results=True)
    assert r2 == gpu_reader.readtext(g) and len(r2) > 0


def test_extract_reference_photo_geometry_with_reduce(gpu_reader):
    """A 5712x4284 cover photo (19 of the reference's 30 fixtures): x1.5 -> 8568x6426 -> cap 1600 needs reduce(); round 1
    returned "" here."""
    bgr = synth.phone_photo(3300, 5712, 4284)
    gpu_reader.set_precision("fp32")
    a = extractor.extract_text_with_ocr(gpu_reader, bgr, image_index=0, return_results=True, device_resident=False)
    b = extractor.extract_text_with_ocr(gpu_reader, bgr, image_index=0, return_results=True, device_resident=True)
    assert a == b and len(a[1]) > 0
