"""EasyOCR-compatible `Reader` backed by libbbocr.so (hand-written sm_100a kernels).

Mirrors the surface BB-OCR uses:
    pipeline_demo/extractor/enhanced_extractor.py:153   easyocr.Reader(["en"], gpu=use_gpu)
    pipeline_demo/extractor/enhanced_extractor.py:520   reader.readtext(path, paragraph=False, batch_size=1, workers=0)
    pipeline_components/img_to_json/ocr_testing/ocr_engines/test_easyocr.py:20-23,50-53   for (bbox, text, prob) in result
Same names, argument meaning, result format `(box, text, confidence)` and error behaviour (ordinary Python
exceptions; the caller swallows them).  There is no CPU fallback: constructing a Reader without a B200 raises.
"""
from __future__ import annotations

import json
import os
import threading

import cv2
import numpy as np

from . import _lib, decode, weights

# easyocr/config.py : recognition_models['gen2']['english_g2']
SYMBOLS = "0123456789!\"#$%&'()*+,-./:;<=>?@[\\]^_`{|}~ €"
CHARACTERS = SYMBOLS + "ABCDEFGHIJKLMNOPQRSTUVWXYZabcdefghijklmnopqrstuvwxyz"


def _is_pil_image(obj) -> bool:
    try:
        from PIL import Image
        return isinstance(obj, Image.Image)
    except Exception:                                        # noqa: BLE001 -- Pillow is optional for the OCR stage
        return False


def _read_file_both(path, handle):
    """The two reads upstream makes of a file -- cv2.imread(path, IMREAD_GRAYSCALE) and a colour read -- from ONE device
    decode when the file is a baseline JPEG (bit-exact with cv2), else from cv2 on the host."""
    if handle is not None:
        try:
            with open(os.path.expanduser(path), "rb") as f:
                data = f.read()
            if decode._is_jpeg(data):
                return decode.imdecode_both(handle, data)
        except (OSError, _lib.BbocrError):
            pass
    return cv2.imread(os.path.expanduser(path), cv2.IMREAD_COLOR), cv2.imread(path, cv2.IMREAD_GRAYSCALE)


def reformat_input(image, handle=None):
    """easyocr/utils.py::reformat_input -> (img HxWx3 as fed to the detector, img_cv_grey HxW).  With a `handle`, baseline
    JPEG files / byte strings decode on the device (bbocr_jpeg_decode); everything else decodes on the host as before."""
    if isinstance(image, str):
        bgr, img_cv_grey = _read_file_both(image, handle)
        if bgr is None or img_cv_grey is None:
            raise ValueError(f"Invalid input: could not read {image}")
        img = cv2.cvtColor(bgr, cv2.COLOR_BGR2RGB)          # upstream: skimage.io.imread (RGB)
    elif isinstance(image, bytes):
        nparr = np.frombuffer(image, np.uint8)
        img = decode.imdecode(handle, image, cv2.IMREAD_COLOR) if handle is not None else cv2.imdecode(nparr, cv2.IMREAD_COLOR)
        if img is None:
            raise ValueError("Invalid input: undecodable bytes")
        img = cv2.cvtColor(img, cv2.COLOR_BGR2RGB)
        img_cv_grey = cv2.cvtColor(img, cv2.COLOR_BGR2GRAY)
    elif isinstance(image, np.ndarray):
        if image.ndim == 2:
            img_cv_grey = image
            img = cv2.cvtColor(image, cv2.COLOR_GRAY2BGR)
        elif image.ndim == 3 and image.shape[2] == 1:
            img_cv_grey = np.squeeze(image)
            img = cv2.cvtColor(img_cv_grey, cv2.COLOR_GRAY2BGR)
        elif image.ndim == 3 and image.shape[2] == 3:
            img = image
            img_cv_grey = None                                # derived on the device: BGR2GRAY fixed-point formula
        elif image.ndim == 3 and image.shape[2] == 4:
            img = cv2.cvtColor(image[:, :, :3], cv2.COLOR_RGB2BGR)
            img_cv_grey = None
        else:
            raise ValueError("Invalid input type. Supporting format = string(file path or url), bytes, numpy array")
    elif _is_pil_image(image):
        # upstream handles PIL JPEG objects: np.array(image) (RGB) -> RGB2BGR for the detector, BGR2GRAY of that for the crops
        img = cv2.cvtColor(np.array(image.convert("RGB")), cv2.COLOR_RGB2BGR)
        img_cv_grey = None                                    # derived on the device with the BGR2GRAY fixed-point formula
    else:
        raise ValueError("Invalid input type. Supporting format = string(file path or url), bytes, numpy array")
    if img.dtype != np.uint8:
        raise ValueError("Invalid input: image must be uint8")
    return img, img_cv_grey


def get_paragraph(raw_result, x_ths=1, y_ths=0.5, mode="ltr"):
    """easyocr/utils.py::get_paragraph -- the `paragraph=True` post-pass of Reader.recognize: greedy clustering of the
    result boxes into paragraphs, then reading order inside each one.  Host logic on a handful of boxes per page.
    -> [[box, text], ...] (no confidence), box = the paragraph's axis-aligned rectangle as four [x, y] int corners."""
    n = len(raw_result)
    text = [r[1] for r in raw_result]
    xs = [[int(c[0]) for c in r[0]] for r in raw_result]
    ys = [[int(c[1]) for c in r[0]] for r in raw_result]
    x_lo, x_hi = [min(v) for v in xs], [max(v) for v in xs]
    y_lo, y_hi = [min(v) for v in ys], [max(v) for v in ys]
    height = [b - a for a, b in zip(y_lo, y_hi)]
    y_mid = [0.5 * (a + b) for a, b in zip(y_lo, y_hi)]
    group = [0] * n                                            # 0 = not assigned yet
    current, members, pending = 1, [], list(range(n))
    while pending:
        if not members:                                        # the first unassigned box opens the group
            k = pending.pop(0)
            group[k] = current
            members = [k]
            continue
        mean_h = float(np.mean([height[k] for k in members]))
        gx0 = min(x_lo[k] for k in members) - x_ths * mean_h
        gx1 = max(x_hi[k] for k in members) + x_ths * mean_h
        gy0 = min(y_lo[k] for k in members) - y_ths * mean_h
        gy1 = max(y_hi[k] for k in members) + y_ths * mean_h
        for pos, k in enumerate(pending):                      # first box (input order) that touches the grown group
            if ((gx0 <= x_lo[k] <= gx1) or (gx0 <= x_hi[k] <= gx1)) and ((gy0 <= y_lo[k] <= gy1) or (gy0 <= y_hi[k] <= gy1)):
                group[k] = current
                members.append(k)
                del pending[pos]
                break
        else:                                                  # nothing joins any more: next group
            current += 1
            members = []
    out = []
    for g in sorted(set(group)):
        left = [k for k in range(n) if group[k] == g]          # input order
        mean_h = float(np.mean([height[k] for k in left]))
        rect = (min(x_lo[k] for k in left), max(x_hi[k] for k in left), min(y_lo[k] for k in left), max(y_hi[k] for k in left))
        words = []
        while left:
            top = min(y_mid[k] for k in left)
            line = [k for k in left if y_mid[k] < top + 0.4 * mean_h]
            if mode == "ltr":
                edge = min(x_lo[k] for k in line)
                pick = [k for k in line if x_lo[k] == edge][-1]
            else:                                              # 'rtl'
                edge = max(x_hi[k] for k in line)
                pick = [k for k in line if x_hi[k] == edge][-1]
            words.append(text[pick])
            # upstream removes by list equality, i.e. the first box with the same fields as the pick
            same = next(k for k in left if (text[k], x_lo[k], x_hi[k], y_lo[k], y_hi[k]) ==
                        (text[pick], x_lo[pick], x_hi[pick], y_lo[pick], y_hi[pick]))
            left.remove(same)
        gx0, gx1, gy0, gy1 = rect
        out.append([[[gx0, gy0], [gx1, gy0], [gx1, gy1], [gx0, gy1]], " ".join(words)])
    return out


class Reader:
    """easyocr.Reader(['en']) drop-in.  Unknown keyword arguments are accepted and ignored like upstream's optional ones.

    precision: "bf16x3" (default) = every layer on the tcgen05 tensor cores in split precision (bf16 hi + lo operands, three
    MMAs per product, FP32 accumulation): score maps / logits within 1e-3 of the FP32 oracle, >= 99.5 % identical (box, string)
    results end to end (tests/test_gpu_e2e_parity.py).  "bf16" = plain-bf16 detector, 2.6x faster, boxes may move by a pixel
    (stated score-map tolerance 6e-2).  "fp32" = CUDA-core parity mode."""

    def __init__(self, lang_list=("en",), gpu=True, model_storage_directory=None, user_network_directory=None,
                 detect_network="craft", recog_network="standard", download_enabled=True, detector=True,
                 recognizer=True, verbose=True, quantize=True, cudnn_benchmark=False, *, precision="bf16x3",
                 craft_state=None, crnn_state=None, **_ignored):
        if list(lang_list) != ["en"]:
            raise ValueError(f"{list(lang_list)} is not supported (only ['en'] / english_g2)")
        if detect_network != "craft":
            raise ValueError("only detect_network='craft' is supported")
        device = 0
        if isinstance(gpu, str) and gpu.startswith("cuda:"):
            device = int(gpu.split(":")[1])
        elif isinstance(gpu, int) and not isinstance(gpu, bool):
            device = gpu
        elif os.environ.get("LOCAL_RANK") is not None:
            device = int(os.environ["LOCAL_RANK"])
        self.device = f"cuda:{device}"
        self._h = _lib.Handle(device)                         # raises without a B200: no CPU path
        self._lock = threading.Lock()
        self.character = CHARACTERS
        self.lang_char = CHARACTERS                            # en_char.txt + symbols cover the whole english_g2 alphabet
        self.model_lang = "english"
        self.dict_list = []                                    # wordbeamsearch dictionary (set_dictionary)
        craft_path, crnn_path = weights.find_checkpoints(model_storage_directory)
        if craft_state is None:
            craft_state = weights.load_pth(craft_path) if craft_path else weights.calibrated_craft_state()
        if crnn_state is None:
            crnn_state = weights.load_pth(crnn_path) if crnn_path else weights.calibrated_crnn_state()
        self.weights_source = {"craft": craft_path or "seeded-random+probe", "crnn": crnn_path or "seeded-random+probe"}
        if verbose and not (craft_path and crnn_path):
            print("bbocr_b200: genuine EasyOCR checkpoints not found; using the seeded synthetic weights "
                  "(bbocr_b200/weights.py)")
        self._h.load_craft(craft_state)
        self._h.load_crnn(crnn_state)
        self.set_precision(precision)
        # CTCLabelConverter.__init__ reads easyocr/dict/en.txt for decoder='wordbeamsearch'; use it when the package is around
        try:
            import importlib.util
            spec = importlib.util.find_spec("easyocr")
            path = os.path.join(os.path.dirname(spec.origin), "dict", "en.txt") if spec and spec.origin else None
            if path and os.path.exists(path):
                with open(path, "r", encoding="utf-8-sig") as f:
                    self.set_dictionary(f.read().splitlines())
        except Exception:                                      # noqa: BLE001 -- optional, like upstream's try/except around the file
            pass

    # ------------------------------------------------------------------------------------------------------------
    def set_precision(self, precision: str):
        self._h.set_precision({"fp32": _lib.PREC_FP32, "bf16": _lib.PREC_BF16, "bf16x3": _lib.PREC_BF16X3}[precision])
        self.precision = precision

    def _params(self, kw, allowlist=None, blocklist=None):
        p = self._h.default_params()
        for k in ("min_size", "canvas_size", "contrast_ths", "adjust_contrast", "text_threshold", "low_text",
                  "link_threshold", "mag_ratio", "slope_ths", "ycenter_ths", "height_ths", "width_ths", "add_margin"):
            if k in kw and kw[k] is not None:
                setattr(p, k, kw[k])
        ignore = None
        if allowlist:
            ign = set(self.character) - set(allowlist)
        elif blocklist:
            ign = set(blocklist)
        else:
            ign = set(self.character) - set(self.lang_char)
        if ign:
            ignore = np.zeros(len(self.character) + 1, np.uint8)
            for ch in ign:
                if ch in self.character:
                    ignore[self.character.index(ch) + 1] = 1
            p.ignore = ignore.ctypes.data
        return p, ignore

    def _format(self, raw, detail=1, output_format="standard", paragraph=False, x_ths=1.0, y_ths=0.5):
        """Tail of Reader.recognize / readtext (easyocr/easyocr.py): paragraph merge, then detail / output_format."""
        out = []
        chars = self.character
        for box, is_free, idx, conf in raw:
            text = "".join([chars[i - 1] for i in idx])
            b = box.tolist() if is_free else box.astype(np.int64).tolist()      # upstream: floats for free boxes, ints otherwise
            out.append((b, text, conf))
        if paragraph:
            out = get_paragraph(out, x_ths=x_ths, y_ths=y_ths, mode="ltr")      # 'rtl' only for the arabic model
        if detail == 0:
            return [item[1] for item in out]
        if output_format == "dict":
            if paragraph:
                return [{"boxes": item[0], "text": item[1]} for item in out]
            return [{"boxes": item[0], "text": item[1], "confident": item[2]} for item in out]
        if output_format == "json":
            rows = [{"boxes": [list(map(int, c)) for c in item[0]], "text": item[1]} for item in out]
            if not paragraph:
                for row, item in zip(rows, out):
                    row["confident"] = item[2]
            return [json.dumps(row, ensure_ascii=False) for row in rows]
        if output_format != "standard":
            raise NotImplementedError(f"output_format={output_format!r} is not implemented (SURVEY.md §8f-3)")
        return out

    def _decode_options(self, p, decoder="greedy", beamWidth=5, batch_size=1, rotation_info=None):
        """decoder / beamWidth / batch_size / rotation_info of Reader.recognize (easyocr/easyocr.py) -> bbocr_params."""
        try:
            p.decoder = {"greedy": 0, "beamsearch": 1, "wordbeamsearch": 2}[decoder]
        except KeyError:
            raise ValueError(f"unknown decoder {decoder!r} (greedy, beamsearch, wordbeamsearch)") from None
        p.beam_width = int(beamWidth)
        p.space_idx = self.character.index(" ") + 1
        rot = [int(a) for a in (rotation_info or [])]
        if any(a not in (90, 180, 270) for a in rot) or len(rot) > 3:
            raise ValueError("rotation_info: eligible values are 90, 180 and 270")
        p.n_rotations = len(rot)
        for i, a in enumerate(rot):
            p.rotation[i] = a
        # upstream takes its per-box branch only for batch_size == 1 (or a CPU reader) without rotation_info; anything else
        # goes through get_image_list over the whole page: one max_width, results ordered by y
        p.batch_mode = 1 if (int(batch_size) > 1 or rot) else 0

    def set_dictionary(self, words):
        """Dictionary of decoder='wordbeamsearch' (upstream reads easyocr/dict/en.txt at construction; that file is part of
        the easyocr package, which this image does not have).  Words with characters outside the alphabet cannot match."""
        idx = []
        for w in words:
            if w and all(ch in self.character for ch in w):
                idx.append([self.character.index(ch) + 1 for ch in w])
        self.dict_list = list(words)
        with self._lock:
            self._h.set_dictionary(idx)

    def readtext(self, image, decoder="greedy", beamWidth=5, batch_size=1, workers=0, allowlist=None, blocklist=None,
                 detail=1, rotation_info=None, paragraph=False, min_size=20, contrast_ths=0.1, adjust_contrast=0.5,
                 filter_ths=0.003, text_threshold=0.7, low_text=0.4, link_threshold=0.4, canvas_size=2560, mag_ratio=1.0,
                 slope_ths=0.1, ycenter_ths=0.5, height_ths=0.5, width_ths=0.5, y_ths=0.5, x_ths=1.0, add_margin=0.1,
                 threshold=0.2, bbox_min_score=0.2, bbox_min_size=3, max_candidates=0, output_format="standard"):
        """Reader.readtext (easyocr/easyocr.py).  Returns [(box, text, confidence), ...] in upstream order."""
        return self.readtext_batched([image], decoder=decoder, beamWidth=beamWidth, batch_size=batch_size, allowlist=allowlist,
                                     blocklist=blocklist, detail=detail,
                                     rotation_info=rotation_info, paragraph=paragraph, min_size=min_size,
                                     contrast_ths=contrast_ths, adjust_contrast=adjust_contrast,
                                     text_threshold=text_threshold, low_text=low_text, link_threshold=link_threshold,
                                     canvas_size=canvas_size, mag_ratio=mag_ratio, slope_ths=slope_ths,
                                     ycenter_ths=ycenter_ths, height_ths=height_ths, width_ths=width_ths,
                                     add_margin=add_margin, y_ths=y_ths, x_ths=x_ths, output_format=output_format)[0]

    def readtext_batched(self, images, decoder="greedy", allowlist=None, blocklist=None, detail=1, rotation_info=None,
                         paragraph=False, output_format="standard", return_stats=False, y_ths=0.5, x_ths=1.0, beamWidth=5,
                         batch_size=1, **kw):
        """Batched extension: independent pages pipelined over the handle's CUDA streams; per-page semantics are exactly
        those of readtext (batch_size=1 by default, i.e. what BB-OCR passes)."""
        if output_format == "free_merge":
            raise NotImplementedError("output_format='free_merge' (utils.merge_to_free) is not implemented (SURVEY.md §8f-3)")
        p, keep = self._params(kw, allowlist, blocklist)
        self._decode_options(p, decoder, beamWidth, batch_size, rotation_info)
        pages = []
        for im in images:
            img, grey = reformat_input(im, self._h)
            pages.append((np.ascontiguousarray(img), None if grey is None else np.ascontiguousarray(grey),
                          img.shape[0], img.shape[1]))
        with self._lock:
            raw = self._h.readtext_raw(pages, p)
        res = [self._format(r, detail, output_format, paragraph, x_ths, y_ths) for r, _ in raw]
        if return_stats:
            return res, [s for _, s in raw]
        return res

    def readtext_pair(self, img, img_cv_grey, **kw):
        """readtext on the two arrays upstream's reformat_input derives from a FILE: `img` (HxWx3, RGB as skimage reads it)
        feeds the detector, `img_cv_grey` (HxW, cv2.imread(IMREAD_GRAYSCALE)) feeds the crops.  Used by the in-memory
        extractor glue when preprocessing is off (enhanced_extractor.py:446-447, :520)."""
        if img.ndim != 3 or img.shape[2] != 3 or img_cv_grey.shape != img.shape[:2] or img.dtype != np.uint8 or img_cv_grey.dtype != np.uint8:
            raise ValueError("readtext_pair expects an HxWx3 uint8 image and its HxW uint8 gray plane")
        p, keep = self._params(kw)
        with self._lock:
            raw = self._h.readtext_raw([(np.ascontiguousarray(img), np.ascontiguousarray(img_cv_grey), img.shape[0], img.shape[1])], p)
        return self._format(raw[0][0])

    def readtext_device(self, color_ptrs, H, W, **kw):
        """Pages already resident in HBM: `color_ptrs` are raw device pointers to HxWx3 u8 images."""
        p, keep = self._params(kw)
        with self._lock:
            raw = self._h.readtext_raw([(ptr, None, H, W) for ptr in color_ptrs], p, on_device=True)
        return [self._format(r) for r, _ in raw], [s for _, s in raw]

    def readtext_device_pages(self, pages, **kw):
        """Pages of any sizes already resident in HBM: `pages` = [(color_ptr, gray_ptr | None, H, W), ...] raw device pointers."""
        p, keep = self._params(kw)
        with self._lock:
            raw = self._h.readtext_raw(list(pages), p, on_device=True)
        return [self._format(r) for r, _ in raw], [s for _, s in raw]

    # stage boundaries (upstream signatures, reduced to the arguments that change arithmetic)
    def detect(self, img, min_size=20, text_threshold=0.7, low_text=0.4, link_threshold=0.4, canvas_size=2560,
               mag_ratio=1.0, slope_ths=0.1, ycenter_ths=0.5, height_ths=0.5, width_ths=0.5, add_margin=0.1,
               reformat=True, **_ignored):
        """Reader.detect -> ([horizontal_list], [free_list])"""
        if reformat:
            img, _ = reformat_input(img, self._h)
        with self._lock:
            text, link, ratio = self._h.craft_forward(img, canvas_size, mag_ratio)
            boxes = self._h.det_boxes(text, link, text_threshold, link_threshold, low_text)
        hl, fl = _lib.group_boxes(boxes, ratio, slope_ths, ycenter_ths, height_ths, width_ths, add_margin, min_size)
        return [[list(map(int, b)) for b in hl]], [[[[float(x), float(y)] for x, y in q] for q in fl]]

    def recognize(self, img_cv_grey, horizontal_list=None, free_list=None, decoder="greedy", beamWidth=5, batch_size=1,
                  workers=0, allowlist=None, blocklist=None, detail=1, rotation_info=None, paragraph=False, contrast_ths=0.1,
                  adjust_contrast=0.5, filter_ths=0.003, y_ths=0.5, x_ths=1.0, reformat=True, output_format="standard",
                  **_ignored):
        """Reader.recognize -> [(box, text, confidence)] for the given boxes (upstream order: horizontal, then free).
        With both lists None the whole image is one horizontal box, like upstream."""
        if reformat:
            if not (isinstance(img_cv_grey, np.ndarray) and img_cv_grey.ndim == 2):
                img, img_cv_grey = reformat_input(img_cv_grey, self._h)
                if img_cv_grey is None:                        # HxWx3 / HxWx4 / PIL input: upstream derives BGR2GRAY of `img`
                    img_cv_grey = cv2.cvtColor(img, cv2.COLOR_BGR2GRAY)
        if horizontal_list is None and free_list is None:
            y_max, x_max = img_cv_grey.shape
            horizontal_list, free_list = [[0, x_max, 0, y_max]], []
        p, keep = self._params({"contrast_ths": contrast_ths, "adjust_contrast": adjust_contrast}, allowlist, blocklist)
        self._decode_options(p, decoder, beamWidth, batch_size, rotation_info)
        with self._lock:
            raw, _ = self._h.recognize_raw(np.ascontiguousarray(img_cv_grey), horizontal_list or [], free_list or [], p)
        return self._format(raw, detail, output_format, paragraph, x_ths, y_ths)

    def score_maps(self, img, canvas_size=2560, mag_ratio=1.0):
        with self._lock:
            return self._h.craft_forward(img, canvas_size, mag_ratio)

    @property
    def handle(self):
        return self._h
