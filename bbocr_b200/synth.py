"""Seeded synthetic inputs for BASELINE.json's configs.

No fonts ship in the image except Pillow's built-in scalable default and OpenCV's Hershey strokes, so those are what
the pages are drawn with.  Everything is deterministic in (seed, size).
"""
from __future__ import annotations

import numpy as np
import cv2
from PIL import Image, ImageDraw, ImageFont

ALPHABET = "0123456789!\"#$%&'()*+,-./:;<=>?@[\\]^_`{|}~ €ABCDEFGHIJKLMNOPQRSTUVWXYZabcdefghijklmnopqrstuvwxyz"
_WORDS = ("the red men of iowa history annals society volume press university edition chapter printed bound "
          "copyright reserved library congress catalog number first published new york london boston company "
          "illustrated by with introduction notes author title collected works essays memoir river prairie").split()


def _font(px: int):
    return ImageFont.load_default(size=int(px))


def _phrase(rng, nwords, upper_p=0.3):
    ws = [str(_WORDS[int(rng.integers(len(_WORDS)))]) for _ in range(nwords)]
    if rng.random() < upper_p:
        ws = [w.upper() for w in ws]
    else:
        ws = [w.capitalize() if rng.random() < 0.5 else w for w in ws]
    return " ".join(ws)


def _shading(rng, h, w, lo, hi):
    g = rng.uniform(lo, hi, (4, 4)).astype(np.float32)
    return cv2.resize(g, (w, h), interpolation=cv2.INTER_CUBIC)


def title_page(seed: int, width: int = 1920, height: int = 1440, return_mask: bool = False):
    """Config 2: off-white paper + low-frequency shading, 8-25 centred lines 24-70 px, imprint lines.  RGB u8 HxWx3.
    With return_mask also the u8 ink mask (255 on glyph pixels) the page was drawn with."""
    rng = np.random.default_rng(seed)
    base = _shading(rng, height, width, 215, 245)
    img = Image.fromarray(np.clip(base, 0, 255).astype(np.uint8)).convert("L")
    d = ImageDraw.Draw(img)
    mask = Image.new("L", (width, height), 0)
    dm = ImageDraw.Draw(mask)
    y = int(height * 0.06)
    nlines = int(rng.integers(8, 26))
    for i in range(nlines):
        px = int(rng.integers(24, 71)) if i < nlines - 3 else int(rng.integers(16, 23))
        f = _font(px)
        text = _phrase(rng, int(rng.integers(2, 7)))
        tw = d.textlength(text, font=f)
        while tw > width * 0.9 and " " in text:
            text = text.rsplit(" ", 1)[0]
            tw = d.textlength(text, font=f)
        x = int((width - tw) / 2)
        d.text((x, y), text, fill=int(rng.integers(10, 60)), font=f)
        dm.text((x, y), text, fill=255, font=f)
        y += int(px * rng.uniform(1.3, 2.0))
        if y > height * 0.93:
            break
    a = np.asarray(img).astype(np.float32) + rng.normal(0, 3, (height, width)).astype(np.float32)
    g = np.clip(a, 0, 255).astype(np.uint8)
    tint = np.array([1.0, 0.985, 0.94], np.float32)
    page = np.clip(g[..., None].astype(np.float32) * tint, 0, 255).astype(np.uint8)
    return (page, np.asarray(mask)) if return_mask else page


def book_cover(seed: int, width: int = 1280, height: int = 960, return_mask: bool = False):
    """Config 1: colour-gradient background, large title words, author lines, some lines rotated +-2 deg.  RGB u8."""
    rng = np.random.default_rng(seed)
    chans = [_shading(rng, height, width, 20, 120) for _ in range(3)]
    bg = np.stack(chans, -1)
    bg += rng.normal(0, 4, bg.shape).astype(np.float32)
    img = Image.fromarray(np.clip(bg, 0, 255).astype(np.uint8))
    mask = Image.new("L", (width, height), 0)
    y = int(height * 0.08)
    specs = [(int(rng.integers(60, 141)), 1 + int(rng.integers(1, 3))) for _ in range(int(rng.integers(2, 5)))]
    specs += [(int(rng.integers(30, 51)), int(rng.integers(2, 4))) for _ in range(int(rng.integers(1, 3)))]
    specs += [(int(rng.integers(22, 30)), int(rng.integers(2, 5)))]
    for px, nw in specs:
        f = _font(px)
        text = _phrase(rng, nw, upper_p=0.6)
        layer = Image.new("RGBA", (width, int(px * 1.6)), (0, 0, 0, 0))
        ld = ImageDraw.Draw(layer)
        tw = ld.textlength(text, font=f)
        while tw > width * 0.92 and " " in text:
            text = text.rsplit(" ", 1)[0]
            tw = ld.textlength(text, font=f)
        col = tuple(int(c) for c in rng.integers(200, 256, 3)) + (255,)
        ld.text((int((width - tw) / 2), int(px * 0.2)), text, fill=col, font=f)
        if rng.random() < 0.35:
            layer = layer.rotate(float(rng.uniform(-2, 2)), resample=Image.BICUBIC, expand=False)
        img.paste(layer, (0, y), layer)
        alpha = layer.getchannel("A")
        mask.paste(alpha, (0, y), alpha)
        y += int(px * rng.uniform(1.4, 1.9))
        if y > height * 0.9:
            break
    a = np.asarray(img.convert("RGB"))
    if rng.random() < 0.7:
        a = cv2.GaussianBlur(a, (0, 0), float(rng.uniform(0.3, 0.8)))
    a = np.ascontiguousarray(a)
    return (a, np.asarray(mask)) if return_mask else a


def phone_photo(seed: int, width: int = 4032, height: int = 3024) -> np.ndarray:
    """Config 3: a title page warped by a small perspective + rotation, vignetting, noise, desk border.  BGR u8."""
    rng = np.random.default_rng(seed)
    pw, ph = int(width * 0.8) // 2, int(height * 0.84) // 2
    page = title_page(seed + 100000, pw, ph)[..., ::-1]
    page = cv2.resize(page, (pw * 2, ph * 2), interpolation=cv2.INTER_LINEAR)
    desk = np.empty((height, width, 3), np.float32)
    tex = _shading(rng, height, width, 60, 110)
    desk[..., 0] = tex * 0.6
    desk[..., 1] = tex * 0.8
    desk[..., 2] = tex
    src = np.float32([[0, 0], [pw * 2, 0], [pw * 2, ph * 2], [0, ph * 2]])
    ang = np.deg2rad(rng.uniform(-5, 5))
    c, s = np.cos(ang), np.sin(ang)
    ctr = np.float32([width / 2, height / 2])
    dst = []
    for (x, y) in src:
        v = np.float32([x - pw, y - ph])
        v = np.float32([c * v[0] - s * v[1], s * v[0] + c * v[1]]) + ctr
        v += rng.uniform(-0.012, 0.012, 2).astype(np.float32) * np.float32([width, height])
        dst.append(v)
    M = cv2.getPerspectiveTransform(src, np.float32(dst))
    warped = cv2.warpPerspective(page, M, (width, height), flags=cv2.INTER_LINEAR, borderValue=(0, 0, 0))
    mask = cv2.warpPerspective(np.full((ph * 2, pw * 2), 255, np.uint8), M, (width, height))
    out = np.where(mask[..., None] > 0, warped.astype(np.float32), desk)
    yy, xx = np.mgrid[0:height:8, 0:width:8].astype(np.float32)
    vig = 1.0 - 0.35 * (((xx - width / 2) / (width / 2)) ** 2 + ((yy - height / 2) / (height / 2)) ** 2)
    vig = cv2.resize(vig, (width, height), interpolation=cv2.INTER_LINEAR)
    out *= vig[..., None]
    out += rng.normal(0, 2.5, ((height + 1) // 2, (width + 1) // 2, 1)).astype(np.float32).repeat(2, 0).repeat(2, 1)[:height, :width]
    return np.clip(out, 0, 255).astype(np.uint8)


def text_line_crop(rng, width_px: int | None = None) -> np.ndarray:
    """Config 4: one gray crop, H=64, black-on-light text from the 96-char alphabet.  u8 64xW (W <= 800)."""
    n = int(rng.integers(1, 31))
    text = "".join(ALPHABET[int(i)] for i in rng.integers(0, len(ALPHABET), n)).replace("€", "E")
    f = _font(44)
    probe = Image.new("L", (8, 8))
    tw = int(ImageDraw.Draw(probe).textlength(text, font=f)) + 16
    w = min(max(tw, 64), 800) if width_px is None else width_px
    img = Image.new("L", (w, 64), int(rng.integers(190, 246)))
    ImageDraw.Draw(img).text((8, 8), text, fill=int(rng.integers(0, 70)), font=f)
    a = np.asarray(img).astype(np.float32) + rng.normal(0, 3, (64, w)).astype(np.float32)
    return np.clip(a, 0, 255).astype(np.uint8)


def score_maps_for(mask: np.ndarray, rng=None):
    """Synthetic CRAFT-like (text, link) maps (H/2 x W/2, f32) from a page's ink mask: blurred glyph strokes (peaks
    ~0.9) as the region score and a horizontally smeared, weaker copy as the affinity score, so that words form
    connected components with realistic statistics when only random CRAFT weights exist."""
    rng = rng or np.random.default_rng(0)
    m = cv2.resize(mask, (mask.shape[1] // 2, mask.shape[0] // 2), interpolation=cv2.INTER_AREA).astype(np.float32) / 255.0
    text = np.clip(cv2.GaussianBlur(m, (0, 0), 1.6) * 1.7, 0, 0.97)
    link = np.clip(cv2.GaussianBlur(m, (0, 0), 4.0, sigmaY=1.2) * 1.5, 0, 0.9)
    text = text + rng.normal(0, 0.01, text.shape).astype(np.float32)
    link = link + rng.normal(0, 0.01, link.shape).astype(np.float32)
    return np.ascontiguousarray(text.astype(np.float32)), np.ascontiguousarray(link.astype(np.float32))


def sparse_page(seed: int, width: int = 1100, height: int = 800, frame: bool = False) -> np.ndarray:
    """A clean scan for the auto-crop heuristic: flat paper without sensor noise, a few separated text
    blocks, and optionally a ruled frame around one block, which makes the components inside it non-external.  BGR u8."""
    rng = np.random.default_rng(seed)
    img = Image.new("L", (width, height), int(rng.integers(225, 250)))
    d = ImageDraw.Draw(img)
    blocks = [(0.18, 0.15, 3), (0.30, 0.52, 2), (0.12, 0.80, 2)]
    for bx, by, nl in blocks:
        y = int(height * by)
        for _ in range(nl):
            px = int(rng.integers(18, 34))
            d.text((int(width * bx), y), _phrase(rng, int(rng.integers(2, 5))), fill=int(rng.integers(10, 70)), font=_font(px))
            y += int(px * 1.5)
    if frame:
        x0, y0, x1, y1 = int(width * 0.22), int(height * 0.46), int(width * 0.9), int(height * 0.66)
        d.rectangle((x0, y0, x1, y1), outline=30, width=3)
    g = np.asarray(img)
    return np.ascontiguousarray(np.stack([g, g, g], axis=-1))
