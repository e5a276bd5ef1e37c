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


def test_thumbnail_refuses_the_reduce_regime(handle):
    g = np.zeros((400, 6500), np.uint8)                      # 6500 -> 1600 is a > 4x shrink: Pillow would box-reduce first
    with pytest.raises(Exception):
        handle.thumbnail(g, 1600)


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
