"""bbocr_b200 -- B200-native OCR stage of BB-OCR (preprocessing + EasyOCR-compatible readtext) behind libbbocr.so.

Importing the package never touches CUDA; constructing a `Reader` or running a preprocessing step does, and fails loudly
when the in-tree library or a B200 is missing (there is no CPU fallback).
"""
from .reader import Reader, reformat_input, CHARACTERS                      # noqa: F401
from .preprocess import ImagePreprocessor, preprocess_for_book_cover, preprocess_array   # noqa: F401
from .extractor import extract_text_with_ocr, ocr_input_image              # noqa: F401
from . import _lib, decode, extractor, sharding, synth, weights                     # noqa: F401

__all__ = ["Reader", "ImagePreprocessor", "preprocess_for_book_cover", "preprocess_array", "reformat_input",
           "extract_text_with_ocr", "ocr_input_image"]
