"""ctypes binding of libbbocr.so (include/bbocr.h).  The library is built in-tree (bbocr_b200/libbbocr.so); there is
no CPU fallback: every compute entry point needs a B200, and a missing library is a hard ImportError-like failure."""
from __future__ import annotations

import ctypes as C
import os
import subprocess
import threading

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libbbocr.so")

E_UNSUPPORTED = -5
PREC_FP32, PREC_BF16, PREC_BF16X3 = 0, 1, 2


class BbocrError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"libbbocr error {code}: {msg}")
        self.code = code


class Tensor(C.Structure):
    _fields_ = [("name", C.c_char_p), ("data", C.POINTER(C.c_float)), ("ndim", C.c_int32), ("shape", C.c_int64 * 4)]


class PPParams(C.Structure):
    _fields_ = [("scale", C.c_float), ("sigma", C.c_float), ("contrast", C.c_float), ("brightness", C.c_float),
                ("clahe_clip", C.c_float), ("sharpen_percent", C.c_int32), ("resize_mode", C.c_int32)]


class GroupParams(C.Structure):
    _fields_ = [("slope_ths", C.c_double), ("ycenter_ths", C.c_double), ("height_ths", C.c_double),
                ("width_ths", C.c_double), ("add_margin", C.c_double), ("min_size", C.c_int32)]


class Image(C.Structure):
    _fields_ = [("color", C.c_void_p), ("gray", C.c_void_p), ("H", C.c_int32), ("W", C.c_int32),
                ("on_device", C.c_int32)]


class Params(C.Structure):
    _fields_ = [("min_size", C.c_int32), ("canvas_size", C.c_int32), ("contrast_ths", C.c_double),
                ("adjust_contrast", C.c_double), ("text_threshold", C.c_double), ("low_text", C.c_double),
                ("link_threshold", C.c_double), ("mag_ratio", C.c_double), ("slope_ths", C.c_double),
                ("ycenter_ths", C.c_double), ("height_ths", C.c_double), ("width_ths", C.c_double),
                ("add_margin", C.c_double), ("ignore", C.c_void_p), ("decoder", C.c_int32), ("beam_width", C.c_int32),
                ("batch_mode", C.c_int32), ("n_rotations", C.c_int32), ("rotation", C.c_int32 * 3), ("space_idx", C.c_int32)]


class Results(C.Structure):
    _fields_ = [("n", C.c_int32), ("box", C.POINTER(C.c_double)), ("is_free", C.POINTER(C.c_uint8)),
                ("text_off", C.POINTER(C.c_int32)), ("text_idx", C.POINTER(C.c_int32)), ("conf", C.POINTER(C.c_double)),
                ("n_crops", C.c_int32), ("n_components", C.c_int32)]


_lib = None
_lock = threading.Lock()

# every symbol include/bbocr.h declares (tests check the export table against this list)
SYMBOLS = [
    "bbocr_create", "bbocr_destroy", "bbocr_last_error", "bbocr_version", "bbocr_load_craft", "bbocr_load_crnn",
    "bbocr_set_precision", "bbocr_get_precision", "bbocr_preprocess_u8", "bbocr_preprocess_batch_u8", "bbocr_preprocess_scan_u8", "bbocr_preprocess_launches_per_image",
    "bbocr_pp_gray", "bbocr_pp_resize_cubic", "bbocr_pp_gaussian3", "bbocr_pp_contrast", "bbocr_pp_brightness",
    "bbocr_pp_clahe", "bbocr_pp_equalize_hist", "bbocr_pp_unsharp", "bbocr_pp_adaptive_threshold", "bbocr_pp_deskew", "bbocr_craft_forward",
    "bbocr_det_boxes", "bbocr_min_area_box", "bbocr_group_boxes", "bbocr_crop_horizontal", "bbocr_crop_free",
    "bbocr_crnn_forward", "bbocr_ctc_decode", "bbocr_default_params", "bbocr_readtext", "bbocr_readtext_batch",
    "bbocr_recognize", "bbocr_thumbnail_u8", "bbocr_autocrop_rect", "bbocr_external_boxes", "bbocr_rect_morph", "bbocr_results_free", "bbocr_launch_count", "bbocr_reset_launch_count", "bbocr_conv_stats",
    "bbocr_enable_conv_timing", "bbocr_set_dictionary", "bbocr_ctc_beam_decode",
    "bbocr_jpeg_info", "bbocr_jpeg_coefficients", "bbocr_jpeg_decode", "bbocr_jpeg_decode_batch",
]


def build(force: bool = False) -> str:
    """Compile libbbocr.so in-tree with nvcc for sm_100a (bbocr_b200/csrc/Makefile)."""
    if force or not os.path.exists(LIB_PATH):
        subprocess.run(["make", "-C", os.path.join(_HERE, "csrc"), "-j8"], check=True, stdout=subprocess.DEVNULL)
    return LIB_PATH


def lib():
    global _lib
    with _lock:
        if _lib is None:
            if not os.path.exists(LIB_PATH):
                raise BbocrError(-3, f"{LIB_PATH} is not built (run `make -C bbocr_b200/csrc` or __graft_entry__.build()); "
                                     "bbocr_b200 has no CPU fallback")
            L = C.CDLL(LIB_PATH)
            L.bbocr_last_error.restype = C.c_char_p
            L.bbocr_last_error.argtypes = [C.c_void_p]
            L.bbocr_version.restype = C.c_char_p
            L.bbocr_launch_count.restype = C.c_int64
            L.bbocr_launch_count.argtypes = [C.c_void_p]
            L.bbocr_reset_launch_count.argtypes = [C.c_void_p]
            L.bbocr_destroy.argtypes = [C.c_void_p]
            L.bbocr_results_free.argtypes = [C.c_void_p]
            _lib = L
        return _lib


def _u8(a):
    a = np.asarray(a)
    if a.dtype != np.uint8:                                  # never cast silently: float / int16 pixels would wrap
        raise ValueError(f"expected a uint8 image, got dtype {a.dtype}")
    a = np.ascontiguousarray(a)
    return a, a.ctypes.data_as(C.c_void_p)


class Handle:
    """Owns one bbocr_handle (one CUDA device)."""

    def __init__(self, device: int = 0):
        self.L = lib()
        self._h = C.c_void_p()
        rc = self.L.bbocr_create(C.c_int(device), C.byref(self._h))
        if rc != 0:
            raise BbocrError(rc, (self.L.bbocr_last_error(None) or b"").decode())
        self.device = device

    def close(self):
        if getattr(self, "_h", None) and self._h.value:
            self.L.bbocr_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check(self, rc):
        if rc != 0:
            raise BbocrError(rc, (self.L.bbocr_last_error(self._h) or b"").decode())

    # ---- weights -------------------------------------------------------------------------------------------------
    def _tensors(self, sd):
        keep, arr = [], (Tensor * len(sd))()
        for i, (k, v) in enumerate(sd.items()):
            a = np.ascontiguousarray(v, dtype=np.float32)
            keep.append(a)
            arr[i].name = k.encode()
            arr[i].data = a.ctypes.data_as(C.POINTER(C.c_float))
            arr[i].ndim = a.ndim
            for d in range(a.ndim):
                arr[i].shape[d] = a.shape[d]
        return arr, keep

    def load_craft(self, sd):
        arr, keep = self._tensors(sd)
        self._check(self.L.bbocr_load_craft(self._h, arr, len(sd)))

    def load_crnn(self, sd):
        arr, keep = self._tensors(sd)
        self._check(self.L.bbocr_load_crnn(self._h, arr, len(sd)))

    def set_precision(self, prec: int):
        self._check(self.L.bbocr_set_precision(self._h, C.c_int(prec)))

    def get_precision(self) -> int:
        return self.L.bbocr_get_precision(self._h)

    # ---- preprocessing ---------------------------------------------------------------------------------------------
    def _step(self, fn, src, out_shape, *args):
        s, sp = _u8(src)
        out = np.empty(out_shape, np.uint8)
        self._check(fn(self._h, sp, *args, out.ctypes.data_as(C.c_void_p)))
        return out

    def pp_gray(self, bgr):
        if getattr(bgr, "ndim", 0) != 3 or bgr.shape[2] != 3:
            raise ValueError("to_grayscale expects an HxWx3 BGR image (cv2.cvtColor raises for other layouts as well)")
        H, W = bgr.shape[:2]
        return self._step(self.L.bbocr_pp_gray, bgr, (H, W), C.c_int(H), C.c_int(W))

    def pp_resize_cubic(self, src, dH, dW, mode=0):
        H, W = src.shape
        return self._step(self.L.bbocr_pp_resize_cubic, src, (dH, dW), C.c_int(H), C.c_int(W), C.c_int(dH), C.c_int(dW),
                          C.c_int(mode))

    def pp_equalize_hist(self, src):
        H, W = src.shape
        return self._step(self.L.bbocr_pp_equalize_hist, src, (H, W), C.c_int(H), C.c_int(W))

    def pp_gaussian3(self, src, sigma):
        H, W = src.shape
        return self._step(self.L.bbocr_pp_gaussian3, src, (H, W), C.c_int(H), C.c_int(W), C.c_float(sigma))

    def pp_contrast(self, src, factor):
        H, W = src.shape
        return self._step(self.L.bbocr_pp_contrast, src, (H, W), C.c_int(H), C.c_int(W), C.c_float(factor))

    def pp_brightness(self, src, factor):
        H, W = src.shape
        return self._step(self.L.bbocr_pp_brightness, src, (H, W), C.c_int(H), C.c_int(W), C.c_float(factor))

    def pp_clahe(self, src, clip):
        H, W = src.shape
        return self._step(self.L.bbocr_pp_clahe, src, (H, W), C.c_int(H), C.c_int(W), C.c_float(clip))

    def pp_unsharp(self, src, percent, threshold=3):
        H, W = src.shape
        return self._step(self.L.bbocr_pp_unsharp, src, (H, W), C.c_int(H), C.c_int(W), C.c_int(percent), C.c_int(threshold))

    def pp_adaptive_threshold(self, src, method, inv, block, delta):
        H, W = src.shape
        return self._step(self.L.bbocr_pp_adaptive_threshold, src, (H, W), C.c_int(H), C.c_int(W), C.c_int(method),
                          C.c_int(int(inv)), C.c_int(block), C.c_float(delta))

    def pp_deskew(self, src, max_deg=5.0):
        H, W = src.shape
        s, sp = _u8(src)
        out = np.empty((H, W), np.uint8)
        ang = C.c_float()
        self._check(self.L.bbocr_pp_deskew(self._h, sp, C.c_int(H), C.c_int(W), C.c_float(max_deg),
                                           out.ctypes.data_as(C.c_void_p), C.byref(ang)))
        return out, float(ang.value)

    def preprocess(self, bgr, params: PPParams):
        """preprocess_for_book_cover arithmetic on a host BGR image -> host gray image."""
        H, W = bgr.shape[:2]
        s, sp = _u8(bgr)
        dH, dW = int(H * params.scale), int(W * params.scale)
        out = np.empty((dH, dW), np.uint8)
        oh, ow = C.c_int(), C.c_int()
        self._check(self.L.bbocr_preprocess_u8(self._h, sp, C.c_int(H), C.c_int(W), C.c_int(W * 3), C.c_int(0),
                                               C.byref(params), out.ctypes.data_as(C.c_void_p), C.c_int(0),
                                               C.byref(oh), C.byref(ow)))
        assert (oh.value, ow.value) == (dH, dW)
        return out

    def dbg_deskew_scores(self, gray, max_deg: float, variant: int) -> np.ndarray:
        """Test hook: per-angle projection scores from one of the three deskew histogram kernels (0 / 1 / 2)."""
        g, gp = _u8(gray.copy())
        H, W = g.shape
        sc = np.zeros(1024, np.uint64)
        n = C.c_int()
        self._check(self.L.bbocr_dbg_deskew_scores(self._h, gp, C.c_int(H), C.c_int(W), C.c_float(max_deg), C.c_int(variant),
                                                   sc.ctypes.data_as(C.c_void_p), C.c_int(1024), C.byref(n)))
        return sc[:n.value].copy()

    def preprocess_scan(self, bgr, clahe_clip=2.0, block=11, delta=2.0, max_deg=5.0):
        """BASELINE config[2] chain on a host BGR image: gray -> CLAHE -> deskew -> gentle_threshold.  -> (binary image, angle)."""
        s, sp = _u8(bgr)
        H, W = s.shape[:2]
        out = np.empty((H, W), np.uint8)
        ang = C.c_float()
        self._check(self.L.bbocr_preprocess_scan_u8(self._h, sp, C.c_int(H), C.c_int(W), C.c_int(W * 3), C.c_int(0),
                                                    C.c_float(clahe_clip), C.c_int(block), C.c_float(delta), C.c_float(max_deg),
                                                    out.ctypes.data_as(C.c_void_p), C.c_int(0), C.byref(ang)))
        return out, float(ang.value)

    def preprocess_scan_dev(self, bgr_ptr: int, H: int, W: int, out_ptr: int, clahe_clip=2.0, block=11, delta=2.0, max_deg=5.0):
        ang = C.c_float()
        self._check(self.L.bbocr_preprocess_scan_u8(self._h, C.c_void_p(bgr_ptr), C.c_int(H), C.c_int(W), C.c_int(W * 3), C.c_int(1),
                                                    C.c_float(clahe_clip), C.c_int(block), C.c_float(delta), C.c_float(max_deg),
                                                    C.c_void_p(out_ptr), C.c_int(1), C.byref(ang)))
        return float(ang.value)

    def preprocess_batch(self, images, params: PPParams):
        """The chain over a list of same-size host BGR images (bbocr_preprocess_batch_u8) -> list of host gray images."""
        if not images:
            return []
        arrs = [np.ascontiguousarray(im, dtype=np.uint8) for im in images]
        H, W = arrs[0].shape[:2]
        if any(a.shape != (H, W, 3) for a in arrs):
            raise ValueError("preprocess_batch expects same-size HxWx3 uint8 images")
        dH, dW = int(H * params.scale), int(W * params.scale)
        outs = [np.empty((dH, dW), np.uint8) for _ in arrs]
        n = len(arrs)
        ins = (C.c_void_p * n)(*[a.ctypes.data for a in arrs])
        ous = (C.c_void_p * n)(*[o.ctypes.data for o in outs])
        oh, ow = C.c_int(), C.c_int()
        self._check(self.L.bbocr_preprocess_batch_u8(self._h, C.c_int(n), ins, C.c_int(H), C.c_int(W), C.c_int(W * 3), C.c_int(0),
                                                     C.byref(params), ous, C.c_int(0), C.byref(oh), C.byref(ow)))
        return outs

    def preprocess_batch_dev(self, bgr_ptrs, H: int, W: int, params: PPParams, out_ptrs):
        """Device-resident batch (raw CUDA pointers): one call, photos pipelined over the handle's streams."""
        n = len(bgr_ptrs)
        ins = (C.c_void_p * n)(*bgr_ptrs)
        ous = (C.c_void_p * n)(*out_ptrs)
        oh, ow = C.c_int(), C.c_int()
        self._check(self.L.bbocr_preprocess_batch_u8(self._h, C.c_int(n), ins, C.c_int(H), C.c_int(W), C.c_int(W * 3), C.c_int(1),
                                                     C.byref(params), ous, C.c_int(1), C.byref(oh), C.byref(ow)))
        return oh.value, ow.value

    def preprocess_dev(self, bgr_ptr: int, H: int, W: int, params: PPParams, out_ptr: int):
        """Device-resident variant (raw CUDA pointers, e.g. torch.Tensor.data_ptr())."""
        oh, ow = C.c_int(), C.c_int()
        self._check(self.L.bbocr_preprocess_u8(self._h, C.c_void_p(bgr_ptr), C.c_int(H), C.c_int(W), C.c_int(W * 3),
                                               C.c_int(1), C.byref(params), C.c_void_p(out_ptr), C.c_int(1),
                                               C.byref(oh), C.byref(ow)))
        return oh.value, ow.value

    # ---- detector --------------------------------------------------------------------------------------------------
    def craft_forward(self, img, canvas_size=2560, mag_ratio=1.0):
        H, W = img.shape[:2]
        mh, mw, ratio = C.c_int(), C.c_int(), C.c_double()
        self._check(self.L.bbocr_craft_forward(self._h, None, C.c_int(H), C.c_int(W), C.c_int(0), C.c_int(canvas_size),
                                               C.c_double(mag_ratio), None, None, C.byref(mh), C.byref(mw),
                                               C.byref(ratio)))
        s, sp = _u8(img)
        text = np.empty((mh.value, mw.value), np.float32)
        link = np.empty((mh.value, mw.value), np.float32)
        self._check(self.L.bbocr_craft_forward(self._h, sp, C.c_int(H), C.c_int(W), C.c_int(0), C.c_int(canvas_size),
                                               C.c_double(mag_ratio), text.ctypes.data_as(C.c_void_p),
                                               link.ctypes.data_as(C.c_void_p), C.byref(mh), C.byref(mw),
                                               C.byref(ratio)))
        return text, link, ratio.value

    def det_boxes(self, text, link, text_threshold=0.7, link_threshold=0.4, low_text=0.4, cap=65536, host_path=None):
        """getDetBoxes_core on given maps.  host_path: None = the product path (bbocr_det_boxes: everything on the device);
        True / False = the test hook, which also reports whether the device tail was used: -> (boxes, used_device)."""
        t = np.ascontiguousarray(text, np.float32)
        l = np.ascontiguousarray(link, np.float32)
        boxes = np.empty((cap, 4, 2), np.float32)
        n = C.c_int()
        if host_path is None:
            self._check(self.L.bbocr_det_boxes(self._h, t.ctypes.data_as(C.c_void_p), l.ctypes.data_as(C.c_void_p),
                                               C.c_int(t.shape[0]), C.c_int(t.shape[1]), C.c_double(text_threshold),
                                               C.c_double(link_threshold), C.c_double(low_text),
                                               boxes.ctypes.data_as(C.c_void_p), C.c_int(cap), C.byref(n)))
            return boxes[:n.value].copy()
        used = C.c_int()
        self._check(self.L.bbocr_dbg_det_boxes(self._h, t.ctypes.data_as(C.c_void_p), l.ctypes.data_as(C.c_void_p),
                                               C.c_int(t.shape[0]), C.c_int(t.shape[1]), C.c_double(text_threshold),
                                               C.c_double(link_threshold), C.c_double(low_text), C.c_int(1 if host_path else 0),
                                               boxes.ctypes.data_as(C.c_void_p), C.c_int(cap), C.byref(n), C.byref(used)))
        return boxes[:n.value].copy(), bool(used.value)

    # ---- recogniser ------------------------------------------------------------------------------------------------
    def crop_horizontal(self, gray, box):
        g, gp = _u8(gray)
        H, W = g.shape
        b = (C.c_int32 * 4)(*[int(v) for v in box])
        cap = 64 * 64 * 64
        out = np.empty(cap, np.uint8)
        oh, ow, mw = C.c_int(), C.c_int(), C.c_int()
        self._check(self.L.bbocr_crop_horizontal(self._h, gp, C.c_int(H), C.c_int(W), b, out.ctypes.data_as(C.c_void_p),
                                                 C.c_int(cap), C.byref(oh), C.byref(ow), C.byref(mw)))
        return out[:oh.value * ow.value].reshape(oh.value, ow.value).copy(), mw.value

    def crop_free(self, gray, quad):
        g, gp = _u8(gray)
        H, W = g.shape
        q = (C.c_double * 8)(*[float(v) for v in np.asarray(quad, np.float64).reshape(-1)])
        cap = 64 * 64 * 64
        out = np.empty(cap, np.uint8)
        oh, ow, mw = C.c_int(), C.c_int(), C.c_int()
        self._check(self.L.bbocr_crop_free(self._h, gp, C.c_int(H), C.c_int(W), q, out.ctypes.data_as(C.c_void_p),
                                           C.c_int(cap), C.byref(oh), C.byref(ow), C.byref(mw)))
        return out[:oh.value * ow.value].reshape(oh.value, ow.value).copy(), mw.value

    def crnn_forward(self, x):
        x = np.ascontiguousarray(x, np.float32)
        N, Hh, Wm = x.shape
        assert Hh == 64
        T = Wm // 4 - 1
        out = np.empty((N, T, 97), np.float32)
        self._check(self.L.bbocr_crnn_forward(self._h, x.ctypes.data_as(C.c_void_p), C.c_int(N), C.c_int(Wm),
                                              out.ctypes.data_as(C.c_void_p)))
        return out

    def ctc_decode(self, logits, ignore=None):
        lg = np.ascontiguousarray(logits, np.float32)
        N, T, Cc = lg.shape
        idx = np.zeros((N, T), np.int32)
        ln = np.zeros(N, np.int32)
        conf = np.zeros(N, np.float64)
        ig = None
        if ignore is not None:
            ig = np.ascontiguousarray(ignore, np.uint8)
        self._check(self.L.bbocr_ctc_decode(self._h, lg.ctypes.data_as(C.c_void_p), C.c_int(N), C.c_int(T), C.c_int(Cc),
                                            ig.ctypes.data_as(C.c_void_p) if ig is not None else None,
                                            idx.ctypes.data_as(C.c_void_p), ln.ctypes.data_as(C.c_void_p),
                                            conf.ctypes.data_as(C.c_void_p)))
        return [idx[i, :ln[i]].copy() for i in range(N)], conf

    def thumbnail(self, gray, max_dim: int) -> np.ndarray:
        """PIL Image.thumbnail((max_dim, max_dim)) (BICUBIC) of a gray u8 image (bbocr_thumbnail_u8)."""
        g, gp = _u8(gray)
        H, W = g.shape
        oh, ow = C.c_int(), C.c_int()
        self._check(self.L.bbocr_thumbnail_u8(self._h, gp, C.c_int(H), C.c_int(W), C.c_int(0), C.c_int(int(max_dim)), None,
                                              C.c_int(0), C.byref(oh), C.byref(ow)))
        out = np.empty((oh.value, ow.value), np.uint8)
        self._check(self.L.bbocr_thumbnail_u8(self._h, gp, C.c_int(H), C.c_int(W), C.c_int(0), C.c_int(int(max_dim)),
                                              out.ctypes.data_as(C.c_void_p), C.c_int(0), C.byref(oh), C.byref(ow)))
        return out

    def thumbnail_dev(self, gray_ptr: int, H: int, W: int, max_dim: int, out_ptr: int | None = None):
        """Device-resident bbocr_thumbnail_u8: with out_ptr None only the output size (oh, ow) is computed."""
        oh, ow = C.c_int(), C.c_int()
        self._check(self.L.bbocr_thumbnail_u8(self._h, C.c_void_p(gray_ptr), C.c_int(H), C.c_int(W), C.c_int(1), C.c_int(int(max_dim)),
                                              C.c_void_p(out_ptr) if out_ptr else None, C.c_int(1), C.byref(oh), C.byref(ow)))
        return oh.value, ow.value

    def autocrop_rect_dev(self, ptr: int, H: int, W: int, channels: int, margin: int = 0):
        """bbocr_autocrop_rect on a device-resident packed image (channels 1 or 3) -> (x0, y0, x1, y1) or None."""
        rect = (C.c_int32 * 4)()
        found = C.c_int()
        self._check(self.L.bbocr_autocrop_rect(self._h, C.c_void_p(ptr), C.c_int(H), C.c_int(W), C.c_int(channels), C.c_int(W * channels),
                                               C.c_int(1), C.c_int(int(margin)), rect, C.byref(found), None, None, None, C.c_int(0),
                                               None, None))
        return tuple(int(v) for v in rect) if found.value else None

    def autocrop_rect(self, bgr, margin: int = 0, debug: bool = False):
        """_auto_crop_text_region up to the slice (bbocr_autocrop_rect): (x0, y0, x1, y1) or None.  With debug=True also a
        dict with the reference's `mask`, `merged`, the external-contour boxes and the two Otsu thresholds."""
        a = np.ascontiguousarray(bgr, dtype=np.uint8)
        if not (a.ndim == 2 or (a.ndim == 3 and a.shape[2] == 3)):
            raise ValueError("autocrop_rect expects an HxWx3 BGR or HxW gray uint8 image")
        H, W = a.shape[:2]
        ch = 3 if a.ndim == 3 else 1
        rect = (C.c_int32 * 4)()
        found, nb = C.c_int(), C.c_int()
        mask = merged = boxes = None
        otsu = (C.c_int32 * 2)()
        cap = 0
        if debug:
            mask, merged = np.empty((H, W), np.uint8), np.empty((H, W), np.uint8)
            cap = 1 << 16
            boxes = np.zeros((cap, 4), np.int32)
        vp = lambda x: x.ctypes.data_as(C.c_void_p) if x is not None else None      # noqa: E731
        self._check(self.L.bbocr_autocrop_rect(self._h, vp(a), C.c_int(H), C.c_int(W), C.c_int(ch), C.c_int(W * ch), C.c_int(0),
                                               C.c_int(int(margin)), rect, C.byref(found), vp(mask), vp(merged), vp(boxes),
                                               C.c_int(cap), C.byref(nb) if debug else None, otsu if debug else None))
        r = tuple(int(v) for v in rect) if found.value else None
        if not debug:
            return r
        return r, {"mask": mask, "merged": merged, "boxes": boxes[:min(nb.value, cap)].astype(np.int64), "nboxes": nb.value,
                   "otsu": (int(otsu[0]), int(otsu[1]))}

    def external_boxes(self, binary) -> np.ndarray:
        """cv2.findContours(RETR_EXTERNAL) + boundingRect on the device: rows (x, y, w, h) sorted by (y, x, w, h)."""
        b, bp = _u8(binary)
        H, W = b.shape
        nb = C.c_int()
        self._check(self.L.bbocr_external_boxes(self._h, bp, C.c_int(H), C.c_int(W), None, C.c_int(0), C.byref(nb)))
        out = np.zeros((max(nb.value, 1), 4), np.int32)
        self._check(self.L.bbocr_external_boxes(self._h, bp, C.c_int(H), C.c_int(W), out.ctypes.data_as(C.c_void_p),
                                                C.c_int(out.shape[0]), C.byref(nb)))
        return out[:nb.value].astype(np.int64)

    def rect_morph(self, binary, kw: int, kh: int, erode: bool) -> np.ndarray:
        """cv2.dilate / cv2.erode of a binary u8 image with a kw x kh rectangle (default border)."""
        b, bp = _u8(binary)
        H, W = b.shape
        out = np.empty((H, W), np.uint8)
        self._check(self.L.bbocr_rect_morph(self._h, bp, C.c_int(H), C.c_int(W), C.c_int(kw), C.c_int(kh), C.c_int(int(erode)),
                                            out.ctypes.data_as(C.c_void_p)))
        return out

    # ---- whole stage -----------------------------------------------------------------------------------------------
    def default_params(self) -> Params:
        p = Params()
        self.L.bbocr_default_params(C.byref(p))
        return p

    def _unpack(self, rp):
        r = C.cast(rp, C.POINTER(Results)).contents
        n = r.n
        out = []
        if n > 0:                                   # bulk views over the result arrays (no per-element ctypes access)
            boxes = np.ctypeslib.as_array(r.box, (n, 4, 2)).copy()
            free = np.ctypeslib.as_array(r.is_free, (n,)).astype(bool).tolist()
            offs = np.ctypeslib.as_array(r.text_off, (n + 1,)).tolist()
            conf = np.ctypeslib.as_array(r.conf, (n,)).tolist()
            idx = np.ctypeslib.as_array(r.text_idx, (max(offs[n], 1),)).tolist()
            for i in range(n):
                out.append((boxes[i], free[i], idx[offs[i]:offs[i + 1]], conf[i]))
        stats = {"n_crops": r.n_crops, "n_components": r.n_components}
        self.L.bbocr_results_free(rp)
        return out, stats

    def readtext_raw(self, images, params: Params | None = None, on_device: bool = False):
        """images: list of (color, gray|None, H, W) with numpy arrays (host) or int device pointers (on_device)."""
        n = len(images)
        arr = (Image * n)()
        keep = []
        for i, (color, gray, H, W) in enumerate(images):
            if on_device:
                arr[i].color = C.c_void_p(int(color))
                arr[i].gray = C.c_void_p(int(gray)) if gray else None
            else:
                c, cp = _u8(color)
                keep.append(c)
                arr[i].color = cp
                if gray is not None:
                    g, gp = _u8(gray)
                    keep.append(g)
                    arr[i].gray = gp
            arr[i].H, arr[i].W, arr[i].on_device = H, W, int(on_device)
        outs = (C.c_void_p * n)()
        p = C.byref(params) if params is not None else None
        self._check(self.L.bbocr_readtext_batch(self._h, C.c_int(n), arr, p, outs))
        return [self._unpack(outs[i]) for i in range(n)]

    def recognize_raw(self, gray, horizontal_list, free_list, params: Params | None = None):
        """Reader.recognize on a host gray page: boxes -> [(box, is_free, class indices, confidence)] (bbocr_recognize)."""
        g, gp = _u8(gray)
        hl = np.ascontiguousarray(np.asarray(horizontal_list, np.int32).reshape(-1, 4))
        fl = np.ascontiguousarray(np.asarray(free_list, np.float64).reshape(-1, 8))
        out = C.c_void_p()
        p = C.byref(params) if params is not None else None
        self._check(self.L.bbocr_recognize(self._h, gp, C.c_int(g.shape[0]), C.c_int(g.shape[1]), C.c_int(0),
                                           hl.ctypes.data_as(C.c_void_p), C.c_int(len(hl)), fl.ctypes.data_as(C.c_void_p),
                                           C.c_int(len(fl)), p, C.byref(out)))
        return self._unpack(out)

    # ---- image decode -----------------------------------------------------------------------------
    def jpeg_decode(self, data: bytes, color: bool = True, gray: bool = False, ignore_orientation: bool = False):
        """cv2.imdecode of a baseline JPEG on the device -> (bgr HxWx3 | None, gray HxW | None) host arrays."""
        buf = np.frombuffer(data, np.uint8)
        H, W, ch, o = jpeg_info(data)
        if ignore_orientation and o >= 5:
            H, W = W, H
        bgr = np.empty((H, W, 3), np.uint8) if color else None
        g = np.empty((H, W), np.uint8) if gray else None
        oh, ow = C.c_int(), C.c_int()
        self._check(self.L.bbocr_jpeg_decode(self._h, buf.ctypes.data_as(C.c_void_p), C.c_size_t(buf.size), C.c_int(int(ignore_orientation)),
                                             bgr.ctypes.data_as(C.c_void_p) if color else None,
                                             g.ctypes.data_as(C.c_void_p) if gray else None, C.c_int(0), C.byref(oh), C.byref(ow)))
        assert (oh.value, ow.value) == (H, W)
        return bgr, g

    def jpeg_decode_batch_dev(self, datas, bgr_ptrs=None, gray_ptrs=None, ignore_orientation: bool = False):
        """n JPEG byte strings -> n device images (raw device pointers sized from jpeg_info), pipelined over the lanes."""
        n = len(datas)
        bufs = [np.frombuffer(d, np.uint8) for d in datas]
        dp = (C.c_void_p * n)(*[b.ctypes.data for b in bufs])
        sz = (C.c_size_t * n)(*[b.size for b in bufs])
        bp = (C.c_void_p * n)(*[int(p) for p in bgr_ptrs]) if bgr_ptrs is not None else None
        gp = (C.c_void_p * n)(*[int(p) for p in gray_ptrs]) if gray_ptrs is not None else None
        self._check(self.L.bbocr_jpeg_decode_batch(self._h, C.c_int(n), dp, sz, C.c_int(int(ignore_orientation)), bp, gp))

    # ---- instrumentation -------------------------------------------------------------------------------------------
    def set_dictionary(self, words):
        """wordbeamsearch dictionary: `words` = iterable of class-index sequences (CTCLabelConverter.dict_list)."""
        words = [list(w) for w in words]
        off = np.zeros(len(words) + 1, np.int32)
        if words:
            off[1:] = np.cumsum([len(w) for w in words])
        idx = np.array([c for w in words for c in w], np.int32) if words else np.zeros(1, np.int32)
        self._check(self.L.bbocr_set_dictionary(self._h, idx.ctypes.data_as(C.c_void_p), off.ctypes.data_as(C.c_void_p),
                                                C.c_int(len(words))))

    def launch_count(self) -> int:
        return int(self.L.bbocr_launch_count(self._h))

    def reset_launch_count(self):
        self.L.bbocr_reset_launch_count(self._h)

    def enable_conv_timing(self, on: bool):
        self._check(self.L.bbocr_enable_conv_timing(self._h, C.c_int(int(on))))

    def conv_stats(self):
        ms, n, fl = C.c_double(), C.c_int64(), C.c_double()
        self._check(self.L.bbocr_conv_stats(self._h, C.byref(ms), C.byref(n), C.byref(fl)))
        return ms.value, n.value, fl.value


# ---- pure-host entry points (usable without a GPU) -----------------------------------------------------------------------

def jpeg_info(data: bytes):
    """(H, W, channels, exif orientation) of a baseline JPEG as cv2.imread would return it (oriented); host only."""
    buf = np.frombuffer(data, np.uint8)
    H, W, ch, o = C.c_int(), C.c_int(), C.c_int(), C.c_int()
    rc = lib().bbocr_jpeg_info(buf.ctypes.data_as(C.c_void_p), C.c_size_t(buf.size), C.byref(H), C.byref(W), C.byref(ch), C.byref(o))
    if rc != 0:
        raise BbocrError(rc, "unsupported or malformed JPEG")
    return H.value, W.value, ch.value, o.value


def jpeg_coefficients(data: bytes) -> np.ndarray:
    """Quantised DCT coefficients [n_blocks][64] int16 (all components, padded to whole MCUs) via the product's host code."""
    buf = np.frombuffer(data, np.uint8)
    nb = C.c_int64()
    rc = lib().bbocr_jpeg_coefficients(buf.ctypes.data_as(C.c_void_p), C.c_size_t(buf.size), None, C.c_int64(0), C.byref(nb))
    if rc != 0:
        raise BbocrError(rc, "unsupported or malformed JPEG")
    out = np.zeros((nb.value, 64), np.int16)
    rc = lib().bbocr_jpeg_coefficients(buf.ctypes.data_as(C.c_void_p), C.c_size_t(buf.size), out.ctypes.data_as(C.c_void_p),
                                       C.c_int64(nb.value), C.byref(nb))
    if rc != 0:
        raise BbocrError(rc, "unsupported or malformed JPEG")
    return out


def ctc_beam_decode(probs, decoder: int, beam_width: int = 5, space_idx: int = 43, dict_words=()):
    """CTCLabelConverter.decode_beamsearch (decoder 1) / decode_wordbeamsearch (2) of one crop: probs T x C float32 -> class
    indices.  Host code of the library (upstream runs it on the CPU as well); needs no device."""
    probs = np.ascontiguousarray(probs, np.float32)
    T, Cn = probs.shape
    words = [list(w) for w in dict_words]
    off = np.zeros(len(words) + 1, np.int32)
    if words:
        off[1:] = np.cumsum([len(w) for w in words])
    idx = np.array([c for w in words for c in w], np.int32) if words else np.zeros(1, np.int32)
    out = np.zeros(max(2 * T + 1, 1), np.int32)
    n = C.c_int()
    rc = lib().bbocr_ctc_beam_decode(probs.ctypes.data_as(C.c_void_p), C.c_int(T), C.c_int(Cn), C.c_int(decoder),
                                     C.c_int(beam_width), C.c_int(space_idx), idx.ctypes.data_as(C.c_void_p),
                                     off.ctypes.data_as(C.c_void_p), C.c_int(len(words)), out.ctypes.data_as(C.c_void_p),
                                     C.c_int(out.size), C.byref(n))
    if rc != 0:
        raise ValueError(f"bbocr_ctc_beam_decode failed ({rc})")
    return out[:n.value].tolist()


def min_area_box(points_xy) -> np.ndarray:
    """cv2.boxPoints(cv2.minAreaRect(points)) restated (bbocr_min_area_box)."""
    p = np.ascontiguousarray(points_xy, np.int32).reshape(-1, 2)
    out = np.empty(8, np.float32)
    rc = lib().bbocr_min_area_box(p.ctypes.data_as(C.c_void_p), C.c_int(len(p)), out.ctypes.data_as(C.c_void_p))
    if rc != 0:
        raise BbocrError(rc, "bbocr_min_area_box")
    return out.reshape(4, 2)


def group_boxes(boxes, ratio, slope_ths=0.1, ycenter_ths=0.5, height_ths=0.5, width_ths=0.5, add_margin=0.1, min_size=20):
    """adjustResultCoordinates + get_textbox + group_text_box + min_size filter (bbocr_group_boxes)."""
    b = np.ascontiguousarray(boxes, np.float32).reshape(-1, 8)
    n = len(b)
    cap = max(n, 1)
    gp = GroupParams(slope_ths, ycenter_ths, height_ths, width_ths, add_margin, min_size)
    hl = np.empty((cap, 4), np.int32)
    fl = np.empty((cap, 4, 2), np.float64)
    nh, nf = C.c_int(), C.c_int()
    rc = lib().bbocr_group_boxes(b.ctypes.data_as(C.c_void_p), C.c_int(n), C.c_double(ratio), C.byref(gp),
                                 hl.ctypes.data_as(C.c_void_p), C.byref(nh), fl.ctypes.data_as(C.c_void_p), C.byref(nf),
                                 C.c_int(cap))
    if rc != 0:
        raise BbocrError(rc, "bbocr_group_boxes")
    return hl[:nh.value].copy(), fl[:nf.value].copy()
