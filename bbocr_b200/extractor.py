"""In-memory version of the OCR-stage glue of BB-OCR's extractor.

Reference: pipeline_demo/extractor/enhanced_extractor.py::extract_text_with_ocr (:413-561).  There the stage is
    preprocess_for_book_cover(path) -> PNG on disk (:431) -> PIL reopen -> thumbnail cap 1600 / 2400 px -> JPEG q90/95 on disk
    (:486-512) -> easyocr Reader.readtext(jpeg_path, paragraph=False, batch_size=1, workers=0) -> " ".join(texts) (:520-521)
i.e. about a second of PNG/JPEG encode/decode per page around the OCR itself.  Here the same steps run on the device and
the image never leaves memory.  The one deliberate deviation: the lossy JPEG round trip between the cap and readtext is
dropped (the reference only uses it as a transport format); everything else is bit-exact against Pillow / the
preprocessing oracle (tests/test_gpu_extractor.py).
"""
from __future__ import annotations

import cv2
import numpy as np

from . import decode
from .preprocess import CURRENT, pp_params


def ocr_max_dim(image_index=None) -> int:
    """enhanced_extractor.py:494 -- covers (index None / 0) are capped harder than the other pages."""
    return 1600 if (image_index is None or image_index == 0) else 2400


def ocr_input_image(reader, gray: np.ndarray, image_index=None) -> np.ndarray:
    """enhanced_extractor.py:486-512 without the JPEG: PIL thumbnail((m, m)) (BICUBIC) when max(size) > m, on the device."""
    if gray.ndim != 2 or gray.dtype != np.uint8:
        raise ValueError("ocr_input_image expects a gray uint8 image (the preprocessing output)")
    m = ocr_max_dim(image_index)
    if max(gray.shape) <= m:
        return gray
    return reader.handle.thumbnail(gray, m)


def _torch_cuda_available() -> bool:
    try:
        import torch
        return bool(torch.cuda.is_available())
    except Exception:                                        # noqa: BLE001 -- torch is optional: the host-buffer path needs none
        return False


def central_edge_crop(img: np.ndarray, percent: float):
    """_central_edge_crop (:374-397) in memory: the centred view with `percent` removed from each edge, or None when the
    reference returns None (percent <= 0, or the rest would be under max(16 px, 20 %) of a side)."""
    if percent <= 0.0:
        return None
    h, w = img.shape[:2]
    mx = int(round(w * (percent / 100.0)))
    my = int(round(h * (percent / 100.0)))
    x0, y0, x1, y1 = max(0, mx), max(0, my), min(w, w - mx), min(h, h - my)
    if x1 - x0 < max(16, w * 0.2) or y1 - y0 < max(16, h * 0.2):
        return None
    return img[y0:y1, x0:x1]


def auto_crop_text_region(reader, img: np.ndarray, margin: int):
    """_auto_crop_text_region (:239-372) in memory: the view img[y0:y1, x0:x1] of the dominant text region, or None for
    "no crop".  The rectangle comes from the device (bbocr_autocrop_rect: text-cue mask, rectangle morphology, external
    components), identical to the cv2 result; `img` is an HxWx3 BGR page or the gray preprocessing output."""
    rect = reader.handle.autocrop_rect(img, int(margin))
    if rect is None:
        return None
    x0, y0, x1, y1 = rect
    return img[y0:y1, x0:x1]


def _decode_to_device(reader, path: str):
    """A JPEG file straight into HBM (bbocr_jpeg_decode, bit-exact with cv2.imread): the photo never exists as a host array.
    None when the file is not something the device decoder takes (the caller then reads it with cv2)."""
    import torch
    from . import _lib
    try:
        with open(path, "rb") as f:
            data = f.read()
        H, W, _ch, _o = _lib.jpeg_info(data)
    except Exception:                                        # noqa: BLE001 -- not a baseline JPEG: host decode
        return None
    out = torch.empty((H, W, 3), dtype=torch.uint8, device=torch.device(reader.device))
    torch.cuda.synchronize(out.device)
    with reader._lock:
        reader.handle.jpeg_decode_batch_dev([data], [out.data_ptr()], None)
    return out


def _extract_device_resident(reader, bgr, image_index, edge_crop_percent, crop_for_ocr, crop_margin):
    """The same steps with the planes kept in HBM between them (one upload of the photo, results down): torch only owns the
    device buffers (allocation, slicing, the gray -> 3-channel copy readtext's detector input needs); every pixel operation
    is a libbbocr call on raw device pointers, so the bytes are those of the host-buffer path."""
    import torch
    h = reader.handle
    dev = torch.device(reader.device)
    src = bgr if isinstance(bgr, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(bgr)).to(dev)
    H, W = bgr.shape[:2]
    p = pp_params(CURRENT, 0)
    gray = torch.empty((int(H * p.scale), int(W * p.scale)), dtype=torch.uint8, device=dev)
    torch.cuda.synchronize(dev)
    with reader._lock:
        h.preprocess_dev(src.data_ptr(), H, W, p, gray.data_ptr())
        if edge_crop_percent > 0.0:
            central = central_edge_crop(gray, edge_crop_percent)          # pure slicing: works on the tensor as on an array
            if central is not None:
                gray = central.contiguous()
                torch.cuda.synchronize(dev)
        if crop_for_ocr:
            try:
                rect = h.autocrop_rect_dev(gray.data_ptr(), gray.shape[0], gray.shape[1], 1, int(crop_margin))
                if rect is not None:
                    gray = gray[rect[1]:rect[3], rect[0]:rect[2]].contiguous()
                    torch.cuda.synchronize(dev)
            except Exception as e:                           # noqa: BLE001 -- :483-484
                print(f"    Auto-cropping failed: {e}")
        m = ocr_max_dim(image_index)
        if max(gray.shape) > m:
            oh, ow = h.thumbnail_dev(gray.data_ptr(), gray.shape[0], gray.shape[1], m)
            small = torch.empty((oh, ow), dtype=torch.uint8, device=dev)
            torch.cuda.synchronize(dev)
            h.thumbnail_dev(gray.data_ptr(), gray.shape[0], gray.shape[1], m, small.data_ptr())
            gray = small
        color = gray.unsqueeze(-1).expand(-1, -1, 3).contiguous()         # reformat_input: GRAY2BGR of a 2-D input
        torch.cuda.synchronize(dev)
        params, _keep = reader._params({})
        raw = h.readtext_raw([(color.data_ptr(), gray.data_ptr(), gray.shape[0], gray.shape[1])], params, on_device=True)
    return reader._format(raw[0][0])


def ocr_input_color(reader, bgr: np.ndarray, image_index=None) -> np.ndarray:
    """The cap for a colour page (preprocessing off, :486-497): `img.convert("RGB"); img.thumbnail((m, m))` resamples every
    channel independently, so the device thumbnail runs once per plane."""
    m = ocr_max_dim(image_index)
    if max(bgr.shape[:2]) <= m:
        return bgr
    planes = [reader.handle.thumbnail(np.ascontiguousarray(bgr[:, :, c]), m) for c in range(3)]
    return np.ascontiguousarray(np.stack(planes, axis=-1))


def extract_text_with_ocr(reader, image, *, use_preprocessing=True, image_index=None, return_results=False,
                          edge_crop_percent=0.0, crop_for_ocr=False, crop_margin=16, device_resident=None):
    """extract_text_with_ocr (:413-561) in memory: [preprocess_for_book_cover] -> [edge crop] -> [auto crop] -> OCR-input
    cap -> readtext -> joined text.  `image`: path or BGR / gray uint8 array.  edge_crop_percent / crop_for_ocr /
    crop_margin are the extractor's attributes of the same names (:452-484; a failing crop keeps the current image, like
    there).  With use_preprocessing=False -- or when preprocessing raises (:441-443) -- the ORIGINAL COLOUR page goes on
    (:418, :446-447): the detector sees it in RGB order (upstream reads the file with skimage) and the crops come from its
    BGR2GRAY plane (cv2.imread(IMREAD_GRAYSCALE)).  device_resident (default: whenever torch sees the GPU) keeps the planes
    in HBM between the steps instead of bouncing them through host arrays; the results are identical.  OCR errors are
    swallowed into "" exactly like :529-531."""
    try:
        if device_resident is None:
            device_resident = _torch_cuda_available()
        if isinstance(image, str):
            bgr = None
            if use_preprocessing and device_resident:        # file -> HBM -> text: decode on the device as well (§8f-4)
                dev_bgr = _decode_to_device(reader, image)
                if dev_bgr is not None:
                    try:
                        results = _extract_device_resident(reader, dev_bgr, image_index, edge_crop_percent, crop_for_ocr, crop_margin)
                        text = " ".join([r[1] for r in results])
                        return (text, results) if return_results else text
                    except Exception as e:                   # noqa: BLE001 -- :441-443: fall back to the original image
                        print(f"    Preprocessing failed: {e}")
                        bgr = dev_bgr.cpu().numpy()
                        use_preprocessing = False
            if bgr is None:
                bgr = decode.imread(reader.handle, image)
            if bgr is None:
                raise ValueError(f"Could not load image from {image}")
        else:
            bgr = image
        if bgr.dtype != np.uint8 or bgr.ndim not in (2, 3) or (bgr.ndim == 3 and bgr.shape[2] != 3):
            raise ValueError("expected an HxWx3 BGR or HxW gray uint8 image")
        page = None                                          # gray plane (preprocessed) or the colour page
        if use_preprocessing:
            try:
                if bgr.ndim != 3:
                    raise ValueError("preprocessing expects a BGR image")
                if device_resident:
                    results = _extract_device_resident(reader, bgr, image_index, edge_crop_percent, crop_for_ocr, crop_margin)
                    text = " ".join([r[1] for r in results])
                    return (text, results) if return_results else text
                page = reader.handle.preprocess(bgr, pp_params(CURRENT, 0))    # same handle (and device) as the reader
            except Exception as e:                           # noqa: BLE001 -- :441-443: fall back to the original image
                print(f"    Preprocessing failed: {e}")
                page = None
        if page is None:
            page = bgr
        if edge_crop_percent > 0.0:
            central = central_edge_crop(page, edge_crop_percent)
            if central is not None:
                page = np.ascontiguousarray(central)
        if crop_for_ocr:
            try:
                cropped = auto_crop_text_region(reader, page, crop_margin)
                if cropped is not None:
                    page = np.ascontiguousarray(cropped)
            except Exception as e:                           # noqa: BLE001 -- :483-484
                print(f"    Auto-cropping failed: {e}")
        if page.ndim == 2:
            ocr_in = ocr_input_image(reader, page, image_index)
            results = reader.readtext(ocr_in, paragraph=False, batch_size=1, workers=0)
        else:
            ocr_in = ocr_input_color(reader, page, image_index)
            results = reader.readtext_pair(cv2.cvtColor(ocr_in, cv2.COLOR_BGR2RGB), cv2.cvtColor(ocr_in, cv2.COLOR_BGR2GRAY))
        text = " ".join([r[1] for r in results])
    except Exception as e:                                   # noqa: BLE001 -- mirrors the reference's blanket handler
        print(f"    OCR failed: {e}")
        results, text = [], ""
    return (text, results) if return_results else text
