"""CPU: the JPEG-decode oracle (oracle/jpeg_np.py) pinned against cv2.imdecode of this image (OpenCV 4.13 / libjpeg-turbo
3.1.2), and the host half of the product's decoder (marker parser) -- no device needed.  SURVEY.md §8f-4."""
import io

import cv2
import numpy as np
import pytest

from oracle import jpeg_np as J
from _jpeg_helpers import SAMPLINGS, _photo, encode, with_orientation


@pytest.mark.parametrize("samp", sorted(SAMPLINGS))
def test_oracle_equals_cv2_on_every_sampling_mode(samp):
    for (h, w) in [(64, 64), (37, 53), (1, 1), (8, 8), (17, 16), (33, 100), (120, 7), (3, 2), (2, 3), (16, 5), (5, 4)]:
        for q, rst in ((95, 0), (60, 3), (20, 1)):
            data = encode(_photo(h, w), q, samp, rst)
            buf = np.frombuffer(data, np.uint8)
            assert np.array_equal(J.imdecode(data), cv2.imdecode(buf, cv2.IMREAD_COLOR)), (h, w, q, rst)
            assert np.array_equal(J.imdecode(data, grayscale=True), cv2.imdecode(buf, cv2.IMREAD_GRAYSCALE)), (h, w, q, rst)


def test_oracle_grayscale_file_and_exif_orientations():
    g = cv2.cvtColor(_photo(50, 70), cv2.COLOR_BGR2GRAY)
    ok, buf = cv2.imencode(".jpg", g, [cv2.IMWRITE_JPEG_QUALITY, 80])
    assert np.array_equal(J.imdecode(buf.tobytes()), cv2.imdecode(buf, cv2.IMREAD_COLOR))
    assert np.array_equal(J.imdecode(buf.tobytes(), True), cv2.imdecode(buf, cv2.IMREAD_GRAYSCALE))
    for o in range(1, 9):
        data = with_orientation(_photo(40, 60), o)
        buf = np.frombuffer(data, np.uint8)
        assert np.array_equal(J.imdecode(data), cv2.imdecode(buf, cv2.IMREAD_COLOR)), o
        assert np.array_equal(J.imdecode(data, True), cv2.imdecode(buf, cv2.IMREAD_GRAYSCALE)), o
        assert np.array_equal(J.imdecode(data, ignore_orientation=True),
                              cv2.imdecode(buf, cv2.IMREAD_COLOR | cv2.IMREAD_IGNORE_ORIENTATION)), o


def test_product_parser_reads_geometry_and_refuses_what_it_cannot_decode(lib_built):
    from PIL import Image
    data = with_orientation(_photo(40, 60), 6)
    assert lib_built.jpeg_info(data) == (60, 40, 3, 6)                      # oriented size, like cv2.imread
    assert lib_built.jpeg_info(encode(_photo(33, 100), 90, "422", 4)) == (33, 100, 3, 1)
    b = io.BytesIO()
    Image.fromarray(_photo(64, 64)).save(b, "JPEG", progressive=True)
    with pytest.raises(lib_built.BbocrError) as e:
        lib_built.jpeg_info(b.getvalue())
    assert e.value.code == lib_built.E_UNSUPPORTED
    with pytest.raises(lib_built.BbocrError):
        lib_built.jpeg_info(b"\x89PNG\r\n\x1a\n" + bytes(64))
    with pytest.raises(lib_built.BbocrError):
        lib_built.jpeg_info(data[:200])                                         # truncated header


def test_product_entropy_decoder_matches_the_oracle(lib_built):
    """The product's parser + restart-interval finder + Huffman routine (the code the device threads run, compiled for the
    host) against the oracle's coefficient blocks: every sampling mode, with and without restart intervals."""
    for samp in sorted(SAMPLINGS):
        for (h, w), q, rst in (((64, 64), 95, 0), ((37, 53), 60, 3), ((120, 7), 20, 1), ((33, 100), 85, 7), ((5, 4), 50, 0)):
            data = encode(_photo(h, w), q, samp, rst)
            info = J.parse(data)
            coefs, _ = J.decode_coefficients(data, info)
            want = np.concatenate([c.reshape(-1, 64) for c in coefs])
            got = lib_built.jpeg_coefficients(data)
            assert got.shape == want.shape and np.array_equal(got, want), (samp, h, w, q, rst)
    g = cv2.cvtColor(_photo(50, 70), cv2.COLOR_BGR2GRAY)
    data = cv2.imencode(".jpg", g, [cv2.IMWRITE_JPEG_QUALITY, 80, cv2.IMWRITE_JPEG_RST_INTERVAL, 2])[1].tobytes()
    coefs, _ = J.decode_coefficients(data, J.parse(data))
    assert np.array_equal(lib_built.jpeg_coefficients(data), coefs[0].reshape(-1, 64))


def test_damaged_files_never_crash_the_host_half(lib_built):
    """Truncated, bit-flipped, spliced streams: the parser / restart-interval finder / entropy decoder either decode or
    refuse with a BbocrError -- no out-of-bounds access (the same routines feed the device path their offsets)."""
    rng = np.random.default_rng(0)
    base = [encode(_photo(64, 80), 85, "420", 3), encode(_photo(33, 47), 60, "444", 0), with_orientation(_photo(40, 60), 6),
            encode(_photo(120, 17), 30, "422", 1)]
    decoded = 0
    for it in range(1600):
        d = bytearray(base[it % len(base)])
        mode = it % 4
        if mode == 0:
            d = d[:int(rng.integers(0, len(d)))]
        elif mode == 1:
            for _ in range(int(rng.integers(1, 6))):
                d[int(rng.integers(0, len(d)))] = int(rng.integers(0, 256))
        elif mode == 2:
            i = int(rng.integers(2, min(len(d), 700)))
            d[i:i + int(rng.integers(1, 8))] = bytes(rng.integers(0, 256, int(rng.integers(1, 8)), dtype=np.uint8))
        else:
            i = int(rng.integers(0, len(d)))
            d = d[:i] + bytes(rng.integers(0, 256, int(rng.integers(1, 40)), dtype=np.uint8)) + d[i:]
        try:
            lib_built.jpeg_info(bytes(d))
            lib_built.jpeg_coefficients(bytes(d))
            decoded += 1
        except lib_built.BbocrError:
            pass
    assert decoded > 100                     # many damaged streams still decode (libjpeg is equally tolerant)
