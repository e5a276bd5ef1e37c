"""GPU: bbocr_jpeg_decode (device Huffman per restart interval / host Huffman, device IDCT + up-sampling + colour + EXIF
orientation) bit-exact against the oracle and against cv2.imdecode.  SURVEY.md §8f-4."""
import cv2
import numpy as np
import pytest

from bbocr_b200 import decode, synth
from oracle import jpeg_np as J
from _jpeg_helpers import SAMPLINGS, _photo, encode, with_orientation

pytestmark = pytest.mark.gpu


def _check(handle, data, oracle=True):
    buf = np.frombuffer(data, np.uint8)
    bgr, gray = handle.jpeg_decode(data, color=True, gray=True)
    want, wantg = cv2.imdecode(buf, cv2.IMREAD_COLOR), cv2.imdecode(buf, cv2.IMREAD_GRAYSCALE)
    assert bgr.shape == want.shape and np.array_equal(bgr, want)
    assert np.array_equal(gray, wantg)
    if oracle:
        assert np.array_equal(bgr, J.imdecode(data)) and np.array_equal(gray, J.imdecode(data, True))
    only_gray = handle.jpeg_decode(data, color=False, gray=True)[1]
    assert np.array_equal(only_gray, wantg)


@pytest.mark.parametrize("samp", sorted(SAMPLINGS))
def test_small_images_every_sampling_mode(handle, samp):
    for (h, w) in [(64, 64), (37, 53), (1, 1), (8, 8), (17, 16), (33, 100), (120, 7), (3, 2), (2, 3), (16, 5), (5, 4)]:
        for q, rst in ((95, 0), (60, 3), (20, 1)):
            _check(handle, encode(_photo(h, w), q, samp, rst))


def test_device_huffman_on_restart_intervals(handle):
    """The phone-camera layout (one restart interval per MCU row, 4:2:0, EXIF orientation 6 / 3): entropy decoding on the device."""
    img = synth.phone_photo(3001, 1008, 756)
    for samp, mcu_w in (("420", 16), ("422", 16), ("444", 8)):
        rst = -(-img.shape[1] // mcu_w)
        _check(handle, encode(img, 92, samp, rst), oracle=False)
        _check(handle, encode(img, 35, samp, 7), oracle=False)                 # intervals that straddle MCU rows
    for o in (1, 3, 6, 8, 2, 4, 5, 7):
        _check(handle, with_orientation(img[:400, :600], o), oracle=False)
    # orientation + restart intervals together, as the reference's IMG_*.JPG files have it
    from PIL import Image
    import io
    ex = Image.Exif()
    ex[0x0112] = 6
    b = io.BytesIO()
    Image.fromarray(img[:, :, ::-1]).save(b, "JPEG", quality=90, exif=ex.tobytes(), restart_marker_rows=1)
    data = b.getvalue()
    assert J.parse(data)["restart"] == -(-img.shape[1] // 16) and J.parse(data)["orientation"] == 6
    _check(handle, data, oracle=False)


def test_full_size_photo_and_batch(handle):
    import torch
    img = synth.phone_photo(3002, 4032, 3024)
    datas = [encode(img, 90, "420", 252), encode(img[:, ::-1].copy(), 80, "420", 252), encode(img[:1500, :2000], 90, "420", 0)]
    want = [cv2.imdecode(np.frombuffer(d, np.uint8), cv2.IMREAD_COLOR) for d in datas]
    outs = [torch.empty(w.shape, dtype=torch.uint8, device="cuda") for w in want]
    grays = [torch.empty(w.shape[:2], dtype=torch.uint8, device="cuda") for w in want]
    handle.jpeg_decode_batch_dev(datas, [t.data_ptr() for t in outs], [t.data_ptr() for t in grays])
    for d, w, o, g in zip(datas, want, outs, grays):
        assert np.array_equal(o.cpu().numpy(), w)
        assert np.array_equal(g.cpu().numpy(), cv2.imdecode(np.frombuffer(d, np.uint8), cv2.IMREAD_GRAYSCALE))


def test_imread_surface_and_fallbacks(handle, tmp_path):
    img = _photo(200, 300)
    p = tmp_path / "a.jpg"
    p.write_bytes(encode(img, 90, "420", 19))
    assert np.array_equal(decode.imread(handle, str(p)), cv2.imread(str(p)))
    assert np.array_equal(decode.imread(handle, str(p), cv2.IMREAD_GRAYSCALE), cv2.imread(str(p), cv2.IMREAD_GRAYSCALE))
    assert decode.imread(handle, str(tmp_path / "missing.jpg")) is None
    png = tmp_path / "b.png"
    cv2.imwrite(str(png), img)
    assert np.array_equal(decode.imread(handle, str(png)), img)                 # not a JPEG: host decode, as before
    from PIL import Image
    prog = tmp_path / "c.jpg"
    Image.fromarray(img[:, :, ::-1]).save(str(prog), "JPEG", progressive=True)
    assert np.array_equal(decode.imread(handle, str(prog)), cv2.imread(str(prog)))


def test_damaged_entropy_data_is_survivable(handle):
    """Bit flips / truncation inside the entropy-coded segment (restart intervals -> device Huffman): the call decodes something
    or raises, never faults; the device stays usable (a clean file decodes bit-exactly afterwards)."""
    rng = np.random.default_rng(1)
    img = synth.phone_photo(3003, 640, 480)
    good = encode(img, 85, "420", 40)
    sos = J.parse(good)["scan"]["start"]
    for it in range(120):
        d = bytearray(good)
        if it % 3 == 0:
            d = d[:int(rng.integers(sos + 1, len(d)))]
        else:
            for _ in range(int(rng.integers(1, 20))):
                d[int(rng.integers(sos, len(d)))] = int(rng.integers(0, 256))
        try:
            bgr, gray = handle.jpeg_decode(bytes(d), color=True, gray=True)
            assert bgr.shape == img.shape and gray.shape == img.shape[:2]
        except Exception as e:      # noqa: BLE001
            assert "libbbocr" in str(e)
    _check(handle, good, oracle=False)
