"""Host-side mirror of BB-OCR's preprocessing interface, executed by libbbocr.so on the GPU.

Reference interface (same names, argument meaning, step strings and error behaviour):
    pipeline_demo/ocr_testing/preprocessing/image_preprocessor.py:7-145   class ImagePreprocessor (fluent steps)
    pipeline_demo/ocr_testing/preprocessing/image_preprocessor.py:147-160 preprocess_for_book_cover(image_path, output_path)
Baseline JPEG files decode on the device (decode.py, bit-exact with cv2.imread); other formats and cv2.imwrite stay on the
host; every pixel operation is a CUDA kernel.
"""
from __future__ import annotations

import os
import threading

import cv2
import numpy as np

from . import _lib, decode

_handles = {}
_hlock = threading.Lock()


def _handle(device: int = 0) -> _lib.Handle:
    with _hlock:
        if device not in _handles:
            _handles[device] = _lib.Handle(device)
        return _handles[device]


CURRENT = dict(scale=1.5, sigma=3, contrast=1.9, brightness=1.2, clahe=2.5, sharpen=0.3)     # pipeline_demo :147-160
LEGACY = dict(scale=1.5, sigma=5, contrast=1.3, brightness=None, clahe=2.0, sharpen=0.2)     # ocr_testing legacy :236-242


def pp_params(cfg=CURRENT, resize_mode: int = 0) -> _lib.PPParams:
    return _lib.PPParams(cfg["scale"], cfg["sigma"], cfg["contrast"], cfg["brightness"] or 0.0, cfg["clahe"],
                         int(cfg["sharpen"] * 100), resize_mode)


class ImagePreprocessor:
    """Fluent preprocessing steps; each call runs the corresponding kernel(s) and records the reference's step string."""

    def __init__(self, device: int = 0, resize_mode: int = 0):
        self.preprocessed_image = None
        self.original_image = None
        self.steps_applied = []
        self._device = device
        self._resize_mode = resize_mode      # 0: OpenCV generic fixed-point cubic (bit-exact), 1: real-arithmetic cubic

    def _need(self):
        if self.preprocessed_image is None:
            raise ValueError("No image loaded")
        return _handle(self._device)

    def _gray_first(self, h):
        if self.preprocessed_image.ndim == 3:
            self.to_grayscale()

    def load_image(self, image_path):
        img = decode.imread(_handle(self._device), image_path)        # cv2.imread; baseline JPEG decodes on the device (§8f-4)
        if img is None:
            raise ValueError(f"Could not load image from {image_path}")
        return self.load_array(img)

    def load_array(self, bgr):
        """In-memory entry (extension): start from a BGR (or gray) uint8 array instead of a file."""
        self.original_image = bgr
        self.preprocessed_image = bgr.copy()
        self.steps_applied = ["original"]
        return self

    def to_grayscale(self):
        h = self._need()
        self.preprocessed_image = h.pp_gray(self.preprocessed_image)
        self.steps_applied.append("grayscale")
        return self

    def _require_gray(self, what):
        if self.preprocessed_image.ndim != 2:
            raise NotImplementedError(f"{what}: only single-channel images are implemented on the GPU path "
                                      "(the reference chain converts to gray first)")

    def denoise(self, strength=7):
        h = self._need()
        self._require_gray("denoise")
        self.preprocessed_image = h.pp_gaussian3(self.preprocessed_image, float(strength))
        self.steps_applied.append(f"denoise(strength={strength})")
        return self

    def equalize_histogram(self):
        h = self._need()
        self._gray_first(h)
        self.preprocessed_image = h.pp_equalize_hist(self.preprocessed_image)
        self.steps_applied.append("equalize_histogram")
        return self

    def clahe(self, clip_limit=2.0, tile_grid_size=(8, 8)):
        h = self._need()
        self._gray_first(h)
        if tuple(tile_grid_size) != (8, 8):
            raise NotImplementedError("clahe: only the 8x8 tile grid the reference uses is implemented")
        self.preprocessed_image = h.pp_clahe(self.preprocessed_image, float(clip_limit))
        self.steps_applied.append(f"clahe(clip_limit={clip_limit})")
        return self

    def gentle_threshold(self, block_size=11, constant=2):
        h = self._need()
        self._gray_first(h)
        self.preprocessed_image = h.pp_adaptive_threshold(self.preprocessed_image, 1, False, block_size, float(constant))
        self.steps_applied.append(f"gentle_threshold(block_size={block_size}, constant={constant})")
        return self

    def increase_contrast(self, factor=1.5):
        h = self._need()
        self._require_gray("increase_contrast")
        self.preprocessed_image = h.pp_contrast(self.preprocessed_image, float(factor))
        self.steps_applied.append(f"increase_contrast(factor={factor})")
        return self

    def increase_brightness(self, factor=1.2):
        h = self._need()
        self._require_gray("increase_brightness")
        self.preprocessed_image = h.pp_brightness(self.preprocessed_image, float(factor))
        self.steps_applied.append(f"increase_brightness(factor={factor})")
        return self

    def sharpen(self, amount=0.3):
        h = self._need()
        self._require_gray("sharpen")
        self.preprocessed_image = h.pp_unsharp(self.preprocessed_image, int(amount * 100), 3)
        self.steps_applied.append(f"sharpen(amount={amount})")
        return self

    def remove_borders(self, border_size=5):
        self._need()
        hh, ww = self.preprocessed_image.shape[:2]
        self.preprocessed_image = np.ascontiguousarray(
            self.preprocessed_image[border_size:hh - border_size, border_size:ww - border_size])
        self.steps_applied.append(f"remove_borders(size={border_size})")
        return self

    def resize(self, scale_factor=2.0):
        h = self._need()
        self._require_gray("resize")
        hh, ww = self.preprocessed_image.shape[:2]
        new_h, new_w = int(hh * scale_factor), int(ww * scale_factor)
        self.preprocessed_image = h.pp_resize_cubic(self.preprocessed_image, new_h, new_w, self._resize_mode)
        self.steps_applied.append(f"resize(scale_factor={scale_factor})")
        return self

    def deskew(self, max_degrees=5.0):
        """Not in the reference; see include/bbocr.h::bbocr_pp_deskew for the definition."""
        h = self._need()
        self._gray_first(h)
        self.preprocessed_image, angle = h.pp_deskew(self.preprocessed_image, float(max_degrees))
        self.steps_applied.append(f"deskew(angle={angle:.1f})")
        return self

    def save_image(self, output_path):
        if self.preprocessed_image is None:
            raise ValueError("No image to save")
        os.makedirs(os.path.dirname(output_path), exist_ok=True)
        cv2.imwrite(output_path, self.preprocessed_image)
        return output_path

    def get_image(self):
        return self.preprocessed_image

    def get_steps_applied(self):
        return self.steps_applied


def _steps(cfg):
    s = ["original", "grayscale", f"resize(scale_factor={cfg['scale']})", f"denoise(strength={cfg['sigma']})",
         f"increase_contrast(factor={cfg['contrast']})"]
    if cfg.get("brightness"):
        s.append(f"increase_brightness(factor={cfg['brightness']})")
    return s + [f"clahe(clip_limit={cfg['clahe']})", f"sharpen(amount={cfg['sharpen']})"]


def preprocess_array(bgr: np.ndarray, cfg=CURRENT, resize_mode: int = 0, device: int = 0) -> np.ndarray:
    """The whole chain on an in-memory BGR image through the fused device pipeline (one upload, one download)."""
    if bgr is None or bgr.ndim != 3 or bgr.shape[2] != 3 or bgr.dtype != np.uint8:
        raise ValueError("expected an HxWx3 uint8 BGR image")
    return _handle(device).preprocess(bgr, pp_params(cfg, resize_mode))


def preprocess_for_book_cover(image_path, output_path=None, *, resize_mode: int = 0, device: int = 0):
    """-> (preprocessed gray image, output_path, steps) exactly like the reference function."""
    bgr = decode.imread(_handle(device), image_path)                  # cv2.imread; baseline JPEG decodes on the device (§8f-4)
    if bgr is None:
        raise ValueError(f"Could not load image from {image_path}")
    out = preprocess_array(bgr, CURRENT, resize_mode, device)
    if output_path:
        os.makedirs(os.path.dirname(output_path), exist_ok=True)
        cv2.imwrite(output_path, out)
    return (out, output_path, _steps(CURRENT))
