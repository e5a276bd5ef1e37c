"""CPU: host-side logic of the product (no CUDA calls): export table, box geometry vs cv2 / the oracle, sharding."""
import ctypes
import os
import re
import subprocess
import sys

import cv2
import numpy as np
import pytest

from oracle import easyocr_restated as E

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol(lib_built):
    hdr = open(os.path.join(ROOT, "include", "bbocr.h")).read()
    declared = set(re.findall(r"\b(bbocr_[a-z0-9_]+)\s*\(", hdr))
    L = lib_built.lib()
    for s in declared:
        assert hasattr(L, s), s
    assert declared == set(lib_built.SYMBOLS)
    assert b"sm_100a" in L.bbocr_version()


def test_no_cpu_fallback_without_gpu(lib_built):
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(lib_built.BbocrError):
        lib_built.Handle(0)
    import bbocr_b200
    with pytest.raises(lib_built.BbocrError):
        bbocr_b200.Reader(["en"], verbose=False)


def test_product_never_imports_oracle():
    for dirpath, _, files in os.walk(os.path.join(ROOT, "bbocr_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cpp", ".h", ".cuh")):
                src = open(os.path.join(dirpath, f), errors="ignore").read()
                assert "import oracle" not in src and "from oracle" not in src, f


def test_min_area_box_matches_cv2(lib_built):
    rng = np.random.default_rng(0)
    for it in range(1500):
        kind = it % 3
        if kind == 0:
            h, w = int(rng.integers(3, 40)), int(rng.integers(3, 120))
            m = np.zeros((h, w), np.uint8)
            for _ in range(int(rng.integers(1, 6))):
                cv2.ellipse(m, (int(rng.integers(0, w)), int(rng.integers(0, h))),
                            (int(rng.integers(1, w)), int(rng.integers(1, h))), float(rng.uniform(0, 180)), 0, 360, 255, -1)
            pts = np.roll(np.array(np.where(m != 0)), 1, axis=0).transpose().reshape(-1, 2)
            if len(pts) == 0:
                continue
        elif kind == 1:
            pts = rng.integers(0, 60, (int(rng.integers(1, 40)), 2))
        else:
            t = rng.integers(0, 30, int(rng.integers(1, 10)))
            pts = np.stack([t * int(rng.integers(0, 3)), t * int(rng.integers(0, 3))], 1)
        pts = np.ascontiguousarray(pts, np.int32)
        ref = cv2.boxPoints(cv2.minAreaRect(pts))
        got = lib_built.min_area_box(pts)
        assert np.array_equal(ref.view(np.uint32), got.view(np.uint32)), (kind, pts.tolist())


def test_group_boxes_matches_oracle(lib_built):
    rng = np.random.default_rng(1)
    for it in range(300):
        n = int(rng.integers(0, 40))
        boxes = []
        for _ in range(n):
            cx, cy = rng.uniform(20, 600), rng.uniform(20, 400)
            w, h = rng.uniform(5, 150), rng.uniform(4, 40)
            ang = rng.choice([0, 0, 0, rng.uniform(-25, 25)])
            r = cv2.boxPoints(((cx, cy), (w, h), float(ang)))
            r = np.roll(r, 4 - r.sum(1).argmin(), 0)
            if rng.random() < 0.5:
                r = np.round(r)
            boxes.append(r.astype(np.float32))
        ratio = float(rng.choice([1.0, 0.8, 2560 / 3000]))
        polys = E.boxes_to_polys_int([b.copy() for b in boxes], ratio)
        hl, fl = E.group_text_box(polys, 0.1, 0.5, 0.5, 0.5, 0.1, True)
        hl, fl = E.filter_min_size(hl, fl, 20)
        arr = np.array(boxes, np.float32).reshape(-1, 8) if n else np.zeros((0, 8), np.float32)
        mh, mf = lib_built.group_boxes(arr, ratio)
        assert [list(map(int, a)) for a in hl] == mh.tolist()
        assert len(fl) == len(mf) and all(np.array_equal(np.array(a, np.float64), b) for a, b in zip(fl, mf))


def test_sharding_is_a_partition():
    from bbocr_b200 import sharding
    for world in (1, 2, 4, 8):
        parts = [sharding.shard_round_robin(37, r, world) for r in range(world)]
        assert sorted(sum(parts, [])) == list(range(37))
        costs = [1280 * 960 if i % 2 else 1920 * 1440 for i in range(37)]
        parts = [sharding.shard_by_cost(costs, r, world) for r in range(world)]
        assert sorted(sum(parts, [])) == list(range(37))
        loads = [sum(costs[i] for i in p) for p in parts]
        assert max(loads) - min(loads) <= max(costs)


_WORKER = r"""
import os, sys
sys.path.insert(0, %r)
import torch.distributed as dist
from bbocr_b200 import sharding
dist.init_process_group("gloo", init_method="tcp://127.0.0.1:%%d" %% int(os.environ["PORT"]),
                        rank=int(os.environ["RANK"]), world_size=2)
mine = sharding.shard_by_cost([3, 1, 4, 1, 5, 9, 2, 6], dist.get_rank(), 2)
counts = sharding.gather_counts(len(mine))
assert sum(counts) == 8, counts
dist.barrier()
print("ok", dist.get_rank(), mine, counts)
"""


def test_two_rank_gloo_sharding(tmp_path):
    script = tmp_path / "w.py"
    script.write_text(_WORKER % ROOT)
    port = 29500 + os.getpid() % 2000
    procs = [subprocess.Popen([sys.executable, str(script)], env={**os.environ, "RANK": str(r), "PORT": str(port)},
                              stdout=subprocess.PIPE, stderr=subprocess.STDOUT) for r in range(2)]
    outs = [p.communicate(timeout=120)[0].decode() for p in procs]
    assert all(p.returncode == 0 for p in procs), outs
    assert all("ok" in o for o in outs)


def test_get_paragraph_matches_the_restated_upstream():
    """paragraph=True post-pass (easyocr/utils.py::get_paragraph): product vs oracle on random layouts."""
    from bbocr_b200.reader import get_paragraph
    rng = np.random.default_rng(7)
    words = ["the", "red", "men", "of", "iowa", "x", "x", "history"]
    for it in range(300):
        n = int(rng.integers(1, 25))
        res = []
        for _ in range(n):
            x, y = int(rng.integers(0, 900)), int(rng.integers(0, 12)) * int(rng.integers(20, 70))
            w, h = int(rng.integers(20, 300)), int(rng.integers(10, 60))
            if rng.random() < 0.2:                               # a free-form (float) quad
                q = [[x + rng.random(), y - 3.5], [x + w + 0.25, y + 2.0], [x + w - 0.5, y + h + 4.75], [x - 1.5, y + h]]
            else:
                q = [[x, y], [x + w, y], [x + w, y + h], [x, y + h]]
            res.append((q, str(words[int(rng.integers(len(words)))]), float(rng.random())))
        if it % 5 == 0 and n > 1:
            res[-1] = res[0]                                    # exact duplicates exercise the remove-by-equality rule
        for mode in ("ltr", "rtl"):
            x_ths, y_ths = float(rng.choice([1.0, 0.3, 2.5])), float(rng.choice([0.5, 0.1, 1.5]))
            assert get_paragraph(res, x_ths, y_ths, mode) == E.get_paragraph(res, x_ths, y_ths, mode), (it, mode)
    assert get_paragraph([]) == [] == E.get_paragraph([])


def test_format_tail_detail_and_output_formats():
    """Reader._format without a device: detail=0, dict / json, paragraph variants (easyocr/easyocr.py::readtext tail)."""
    import json
    import bbocr_b200
    r = bbocr_b200.Reader.__new__(bbocr_b200.Reader)
    r.character = bbocr_b200.reader.CHARACTERS
    idx = lambda t: np.array([r.character.index(c) + 1 for c in t], np.int32)          # noqa: E731
    raw = [(np.array([[10, 10], [90, 10], [90, 40], [10, 40]], np.float32), False, idx("Red"), 0.9),
           (np.array([[100, 12], [180, 12], [180, 41], [100, 41]], np.float32), False, idx("Men"), 0.5),
           (np.array([[10.5, 300.25], [90, 310], [88, 340], [9, 330]], np.float32), True, idx("1854"), 0.25)]
    std = r._format(raw)
    assert std[0] == ([[10, 10], [90, 10], [90, 40], [10, 40]], "Red", 0.9) and isinstance(std[2][0][0][0], float)
    assert r._format(raw, detail=0) == ["Red", "Men", "1854"]
    assert r._format(raw, output_format="dict")[1] == {"boxes": std[1][0], "text": "Men", "confident": 0.5}
    assert json.loads(r._format(raw, output_format="json")[2]) == {"boxes": [[10, 300], [90, 310], [88, 340], [9, 330]], "text": "1854", "confident": 0.25}
    para = r._format(raw, paragraph=True)
    assert para == [[[[10, 10], [180, 10], [180, 41], [10, 41]], "Red Men"], [[[9, 300], [90, 300], [90, 340], [9, 340]], "1854"]]
    assert para == E.get_paragraph(std)
    assert r._format(raw, detail=0, paragraph=True) == ["Red Men", "1854"]
    assert r._format(raw, output_format="dict", paragraph=True)[0] == {"boxes": para[0][0], "text": "Red Men"}
    with pytest.raises(NotImplementedError):
        r._format(raw, output_format="free_merge")


def test_central_edge_crop_matches_the_reference_fixtures():
    """_central_edge_crop (enhanced_extractor.py:374-397) is pure host slicing: product vs the rectangles recorded from the
    reference's own function (tests/golden/make_golden_autocrop.py)."""
    import glob
    from bbocr_b200 import extractor
    paths = sorted(glob.glob(os.path.join(ROOT, "tests", "golden", "autocrop_*.npz")))
    assert len(paths) >= 8
    for path in paths:
        z = np.load(path)
        bgr = z["bgr"]
        for pc, *want in z["edge"]:
            got = extractor.central_edge_crop(bgr, float(pc))
            if want[0] < 0:
                assert got is None
            else:
                x0, y0, x1, y1 = (int(v) for v in want)
                assert got.shape == (y1 - y0, x1 - x0, 3) and np.shares_memory(got, bgr) and np.array_equal(got, bgr[y0:y1, x0:x1])


def test_reformat_input_kinds_match_the_oracle(tmp_path):
    """easyocr/utils.py::reformat_input: every input kind gives the detector image and gray page of the restated upstream
    (the product leaves the gray of 3-/4-channel arrays to the device: same fixed-point BGR2GRAY formula)."""
    from PIL import Image
    from bbocr_b200.reader import reformat_input
    from oracle import preprocess_np as P
    rng = np.random.default_rng(3)
    rgb = rng.integers(0, 256, (37, 53, 3), dtype=np.uint8)
    png = str(tmp_path / "p.png")
    cv2.imwrite(png, rgb)
    kinds = [png, open(png, "rb").read(), rgb, rgb[:, :, 0].copy(), rgb[:, :, :1].copy(),
             np.concatenate([rgb, rgb[:, :, :1]], axis=2), Image.fromarray(rgb)]
    for k in kinds:
        img, grey = reformat_input(k)
        oimg, ogrey = E.reformat_input(k)
        assert img.dtype == np.uint8 and np.array_equal(img, oimg)
        if grey is None:
            grey = P.bgr2gray(img)                           # what the device derives
        assert np.array_equal(grey, ogrey)
    for bad in (3.5, None, np.zeros((2, 2, 2), np.uint8), np.zeros((4, 4), np.float32)):
        with pytest.raises(ValueError):
            reformat_input(bad)
    with pytest.raises(ValueError):
        reformat_input(str(tmp_path / "missing.png"))


def _peaky_probs(rng, T, C=97, space=43, sharp=4.0):
    """A probability matrix shaped like a recogniser's softmax: mostly blank, a few confident characters, some spaces."""
    logits = rng.normal(size=(T, C)).astype(np.float32)
    logits[:, 0] += 3.0
    for t in range(T):
        u = rng.random()
        if u < 0.35:
            logits[t, rng.integers(1, C)] += sharp * rng.uniform(0.5, 2.0)
        elif u < 0.45:
            logits[t, space] += sharp
    e = np.exp(logits - logits.max(1, keepdims=True))
    return (e / e.sum(1, keepdims=True)).astype(np.float32)


def test_beam_decoders_match_the_restated_upstream(lib_built):
    """decoder='beamsearch' / 'wordbeamsearch' (easyocr/utils.py::ctcBeamSearch and friends): the library's
    host decoder vs the restated upstream on random softmax-like matrices, several beam widths, with and without a
    dictionary."""
    rng = np.random.default_rng(11)
    chars = E.CHARACTERS
    space = chars.index(" ") + 1
    n = 0
    for trial in range(60):
        T = int(rng.integers(1, 60))
        probs = _peaky_probs(rng, T, sharp=float(rng.uniform(1.0, 6.0)))
        bw = int(rng.choice([1, 2, 5, 10]))
        want = E.decode_beamsearch(probs[None], chars, bw)[0]
        got = "".join(chars[i - 1] for i in lib_built.ctc_beam_decode(probs, 1, bw, space))
        assert got == want, (trial, got, want)
        # dictionary: half of the greedy words (so that some candidates hit), plus noise
        greedy_words = [w for w in E.decode_probs(probs[None], chars)[0][0].split(" ") if w]
        words = [w for k, w in enumerate(greedy_words) if k % 2 == 0] + ["the", "of", "Book"]
        want = E.decode_wordbeamsearch(probs[None], chars, bw, words)[0]
        enc = [[chars.index(ch) + 1 for ch in w] for w in words]
        got = "".join(chars[i - 1] for i in lib_built.ctc_beam_decode(probs, 2, bw, space, enc))
        assert got == want, (trial, got, want)
        want = E.decode_wordbeamsearch(probs[None], chars, bw, [])[0]
        got = "".join(chars[i - 1] for i in lib_built.ctc_beam_decode(probs, 2, bw, space, []))
        assert got == want, (trial, got, want)
        n += 3
    assert n == 180
    # a sharp matrix without repeated characters: every decoder returns the greedy string
    probs = np.full((12, 97), 1e-4, np.float32)
    for t, c in enumerate([0, 30, 0, 0, 31, 32, 0, space, 0, 50, 0, 0]):
        probs[t, c] = 1.0
    probs /= probs.sum(1, keepdims=True)
    greedy = E.decode_probs(probs[None], chars)[0][0]
    assert greedy == "".join(chars[c - 1] for c in (30, 31, 32, space, 50))
    for dec in (1, 2):
        assert "".join(chars[i - 1] for i in lib_built.ctc_beam_decode(probs, dec, 5, space)) == greedy


def test_epilogue_shared_memory_layouts_are_what_the_kernels_assume():
    """Index arithmetic the conv_tc.cu epilogues rely on, restated (no GPU needed).
    (1) 2x2 pooling through shared memory: window w (0..7) of a warp's 32 pixels has its base lane at w's bits spread around bit 0
        and bit log2(TW); {b, b^1, b^TW, b^TW^1} over the eight windows partition the 32 lanes, and b is the lane the shuffle
        path calls the writer ((lane & TW) == 0 and (lane & 1) == 0).
    (2) staging swizzles: a quarter warp's 128-bit shared stores hit 8 distinct 16-byte bank groups
        (pool tile: 64-byte rows, chunk k at k ^ ((lane >> 1) & 3); stem output tile: 128-byte rows, chunk q at q ^ (row & 7)).
    (3) k_conv_stem's A rows: SWIZZLE_64B places 16-byte chunk c of 64-byte row r at chunk c ^ ((r >> 1) & 3), i.e. address bits
        [4:5] ^= address bits [7:8]."""
    for TW in (8, 16):
        seen = set()
        for w in range(8):
            t = w << 1
            b = (t & (TW - 1)) | ((t & ~(TW - 1)) << 1)
            assert (b & TW) == 0 and (b & 1) == 0
            lanes = {b, b ^ 1, b ^ TW, b ^ TW ^ 1}
            assert len(lanes) == 4 and not (lanes & seen) and max(lanes) < 32
            seen |= lanes
        assert seen == set(range(32))
    for quarter in range(4):
        lanes = range(8 * quarter, 8 * quarter + 8)
        for k in range(4):                                   # pool tile: byte offset lane * 64 + ((k ^ sw) << 4)
            groups = {((lane * 64 + ((k ^ ((lane >> 1) & 3)) << 4)) % 128) // 16 for lane in lanes}
            assert len(groups) == 8
        for q in range(8):                                   # stem output tile: byte offset row * 128 + ((q ^ (row & 7)) << 4)
            groups = {((row * 128 + ((q ^ (row & 7)) << 4)) % 128) // 16 for row in lanes}
            assert len(groups) == 8
    for r in range(128):
        for c in range(4):
            addr = r * 64 + c * 16
            swz = addr ^ (((addr >> 7) & 3) << 4)
            assert swz == r * 64 + ((c ^ ((r >> 1) & 3)) << 4)
