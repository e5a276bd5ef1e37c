"""TEST / BENCH INFRASTRUCTURE ONLY: the reference's preprocessing chain restated with the reference's OWN library calls
(OpenCV + Pillow), for timing the CPU arm of stage 1 next to the device chain and as a second pin of oracle/preprocess_np.py.

Follows pipeline_demo/ocr_testing/preprocessing/image_preprocessor.py:147-160 (preprocess_for_book_cover) step by step:
to_grayscale (:25-30), resize(1.5) (:125-132), denoise(3) (:32-37), increase_contrast(1.9) (:70-84),
increase_brightness(1.2) (:86-100), clahe(2.5) (:48-56), sharpen(0.3) (:102-115); file I/O left out.
Only tests/ and bench.py may import this module; the product path (bbocr_b200) never does."""
import cv2
import numpy as np
from PIL import Image, ImageEnhance, ImageFilter


def preprocess_for_book_cover_cv(bgr: np.ndarray) -> np.ndarray:
    img = cv2.cvtColor(bgr, cv2.COLOR_BGR2GRAY)
    h, w = img.shape[:2]
    img = cv2.resize(img, (int(w * 1.5), int(h * 1.5)), interpolation=cv2.INTER_CUBIC)
    img = cv2.GaussianBlur(img, (3, 3), 3)
    img = np.array(ImageEnhance.Contrast(Image.fromarray(img)).enhance(1.9))
    img = np.array(ImageEnhance.Brightness(Image.fromarray(img)).enhance(1.2))
    img = cv2.createCLAHE(clipLimit=2.5, tileGridSize=(8, 8)).apply(img)
    return np.array(Image.fromarray(img).filter(ImageFilter.UnsharpMask(radius=1.0, percent=int(0.3 * 100), threshold=3)))
