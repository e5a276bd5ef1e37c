"""CPU oracle (TEST INFRASTRUCTURE ONLY) for the image decode in front of the OCR path:
    pipeline_demo/ocr_testing/preprocessing/image_preprocessor.py:18      cv2.imread(image_path)
    easyocr/utils.py::reformat_input                                        cv2.imread(path, IMREAD_GRAYSCALE), imdecode(bytes)
i.e. libjpeg-turbo's baseline decoder with its default settings, as OpenCV drives it (grfmt_jpeg.cpp): Huffman decoding
(jdhuff.c), dequantisation + the accurate integer IDCT (jidctint.c::jpeg_idct_islow, CONST_BITS 13 / PASS1_BITS 2), "fancy"
triangle up-sampling of sub-sampled chroma (jdsample.c::h2v1_fancy_upsample / h2v2_fancy_upsample / h1v2_fancy_upsample, with
jdmainct.c's replicated context rows at the top and bottom edge), the table-driven YCbCr -> RGB conversion (jdcolor.c) and
OpenCV's EXIF orientation step (imgcodecs/src/loadsave.cpp::ExifTransform).  Restated in NumPy (pure-Python Huffman loop:
small images only).

PINNED: tests/test_oracle_jpeg.py compares this file with cv2.imdecode of this image's OpenCV 4.13 / libjpeg-turbo 3.1.2 on
every sub-sampling mode, odd sizes, restart intervals, grayscale files and all eight EXIF orientations: bit-exact.
Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline legs may import this module."""
from __future__ import annotations

import numpy as np

ZIGZAG = np.array([0, 1, 8, 16, 9, 2, 3, 10, 17, 24, 32, 25, 18, 11, 4, 5, 12, 19, 26, 33, 40, 48, 41, 34, 27, 20, 13, 6, 7, 14, 21,
                   28, 35, 42, 49, 56, 57, 50, 43, 36, 29, 22, 15, 23, 30, 37, 44, 51, 58, 59, 52, 45, 38, 31, 39, 46, 53, 60, 61,
                   54, 47, 55, 62, 63])


class Unsupported(ValueError):
    pass


def parse(data: bytes):
    """Marker segments of a baseline / extended-sequential Huffman JPEG -> dict (frame, tables, scan bytes)."""
    if data[:2] != b"\xff\xd8":
        raise Unsupported("not a JPEG")
    pos = 2
    qt, dc_tabs, ac_tabs = {}, {}, {}
    frame, scan, restart, orientation = None, None, 0, 1
    while pos < len(data):
        if data[pos] != 0xFF:
            raise Unsupported("marker expected")
        while data[pos] == 0xFF:
            pos += 1
        m = data[pos]
        pos += 1
        if m == 0xD9:
            break
        if m == 0x01 or 0xD0 <= m <= 0xD7:
            continue
        L = (data[pos] << 8) | data[pos + 1]
        seg = data[pos + 2:pos + L]
        pos += L
        if m == 0xDB:
            i = 0
            while i < len(seg):
                pq, tq = seg[i] >> 4, seg[i] & 15
                i += 1
                if pq:
                    vals = [(seg[i + 2 * k] << 8) | seg[i + 2 * k + 1] for k in range(64)]
                    i += 128
                else:
                    vals = list(seg[i:i + 64])
                    i += 64
                q = np.zeros(64, np.int32)
                q[ZIGZAG] = vals                        # tables are stored in zig-zag order
                qt[tq] = q
        elif m in (0xC0, 0xC1):
            if seg[0] != 8:
                raise Unsupported("only 8-bit samples")
            H, W, n = (seg[1] << 8) | seg[2], (seg[3] << 8) | seg[4], seg[5]
            comps = [{"id": seg[6 + 3 * k], "h": seg[7 + 3 * k] >> 4, "v": seg[7 + 3 * k] & 15, "tq": seg[8 + 3 * k]} for k in range(n)]
            frame = {"H": H, "W": W, "comps": comps}
        elif 0xC2 <= m <= 0xCF and m not in (0xC4, 0xC8, 0xCC):
            raise Unsupported("progressive / lossless / arithmetic JPEG")
        elif m == 0xC4:
            i = 0
            while i < len(seg):
                tc, th = seg[i] >> 4, seg[i] & 15
                counts = list(seg[i + 1:i + 17])
                nv = sum(counts)
                vals = list(seg[i + 17:i + 17 + nv])
                i += 17 + nv
                (ac_tabs if tc else dc_tabs)[th] = (counts, vals)
        elif m == 0xDD:
            restart = (seg[0] << 8) | seg[1]
        elif m == 0xE1 and seg[:6] == b"Exif\x00\x00":
            orientation = _exif_orientation(seg[6:]) or orientation
        elif m == 0xDA:
            ns = seg[0]
            sc = [{"id": seg[1 + 2 * k], "td": seg[2 + 2 * k] >> 4, "ta": seg[2 + 2 * k] & 15} for k in range(ns)]
            if frame is None or ns != len(frame["comps"]):
                raise Unsupported("multi-scan (non-interleaved) JPEG")
            scan = {"comps": sc, "start": pos}
            break
    if frame is None or scan is None:
        raise Unsupported("no frame / scan")
    return {"frame": frame, "scan": scan, "qt": qt, "dc": dc_tabs, "ac": ac_tabs, "restart": restart, "orientation": orientation}


def _exif_orientation(tiff: bytes):
    if len(tiff) < 8:
        return None
    le = tiff[:2] == b"II"
    rd = (lambda b: int.from_bytes(b, "little")) if le else (lambda b: int.from_bytes(b, "big"))
    off = rd(tiff[4:8])
    n = rd(tiff[off:off + 2])
    for k in range(n):
        e = tiff[off + 2 + 12 * k: off + 14 + 12 * k]
        if rd(e[0:2]) == 0x0112:
            return rd(e[8:10])
    return None


def _huff_lookup(counts, vals):
    """code -> (length, value) as a dict keyed by (length, code)."""
    table, code, k = {}, 0, 0
    for length in range(1, 17):
        for _ in range(counts[length - 1]):
            table[(length, code)] = vals[k]
            code += 1
            k += 1
        code <<= 1
    return table


class _Bits:
    def __init__(self, data, pos):
        self.d, self.p, self.acc, self.n = data, pos, 0, 0

    def bit(self):
        if self.n == 0:
            b = self.d[self.p] if self.p < len(self.d) else 0
            self.p += 1
            if b == 0xFF:
                nxt = self.d[self.p] if self.p < len(self.d) else 0xD9
                if nxt == 0:
                    self.p += 1
                else:                                    # a marker: feed zeros (libjpeg's behaviour at a premature end)
                    self.p -= 1
                    b = 0
            self.acc, self.n = b, 8
        self.n -= 1
        return (self.acc >> self.n) & 1

    def bits(self, k):
        v = 0
        for _ in range(k):
            v = (v << 1) | self.bit()
        return v

    def symbol(self, table):
        code = 0
        for length in range(1, 17):
            code = (code << 1) | self.bit()
            if (length, code) in table:
                return table[(length, code)]
        raise Unsupported("bad Huffman code")

    def restart(self):
        self.n = 0
        while not (self.d[self.p] == 0xFF and 0xD0 <= self.d[self.p + 1] <= 0xD7):
            self.p += 1
        self.p += 2


def _extend(v, s):
    return v if s == 0 or v >= (1 << (s - 1)) else v - (1 << s) + 1


def decode_coefficients(data: bytes, info):
    """jdhuff.c::decode_mcu over the single interleaved scan -> per component an int16 array [blocks_y][blocks_x][64] in
    natural (row-major) coefficient order, block counts padded to whole MCUs."""
    f, sc = info["frame"], info["scan"]
    comps = f["comps"]
    hmax, vmax = max(c["h"] for c in comps), max(c["v"] for c in comps)
    mcux = -(-f["W"] // (8 * hmax))
    mcuy = -(-f["H"] // (8 * vmax))
    coefs = [np.zeros((mcuy * c["v"], mcux * c["h"], 64), np.int16) for c in comps]
    dct = [_huff_lookup(*info["dc"][s["td"]]) for s in sc["comps"]]
    act = [_huff_lookup(*info["ac"][s["ta"]]) for s in sc["comps"]]
    br = _Bits(data, sc["start"])
    pred = [0] * len(comps)
    ri = info["restart"]
    for m in range(mcux * mcuy):
        if ri and m and m % ri == 0:
            br.restart()
            pred = [0] * len(comps)
        my, mx = divmod(m, mcux)
        for ci, c in enumerate(comps):
            for by in range(c["v"]):
                for bx in range(c["h"]):
                    blk = coefs[ci][my * c["v"] + by, mx * c["h"] + bx]
                    s = br.symbol(dct[ci])
                    pred[ci] += _extend(br.bits(s), s) if s else 0
                    blk[0] = pred[ci]
                    k = 1
                    while k < 64:
                        rs = br.symbol(act[ci])
                        r, s = rs >> 4, rs & 15
                        if s == 0:
                            if r != 15:
                                break
                            k += 16
                            continue
                        k += r
                        blk[ZIGZAG[k & 63]] = _extend(br.bits(s), s)
                        k += 1
    return coefs, (mcux, mcuy, hmax, vmax)


def _descale(x, n):
    return (x + (1 << (n - 1))) >> n


def idct_islow(coef, q):
    """jidctint.c::jpeg_idct_islow on [..., 64] int blocks -> [..., 8, 8] uint8 samples (range-limit table semantics:
    index & 1023 into the post-IDCT table, i.e. clamp(x + 128) for |x| < 512 and libjpeg's wrap beyond)."""
    x = (coef.astype(np.int64) * q.astype(np.int64)).reshape(coef.shape[:-1] + (8, 8))

    def one_d(v0, v1, v2, v3, v4, v5, v6, v7, shift):
        z2, z3 = v2, v6
        z1 = (z2 + z3) * 4433
        tmp2 = z1 + z3 * (-15137)
        tmp3 = z1 + z2 * 6270
        tmp0 = (v0 + v4) << 13
        tmp1 = (v0 - v4) << 13
        tmp10, tmp13, tmp11, tmp12 = tmp0 + tmp3, tmp0 - tmp3, tmp1 + tmp2, tmp1 - tmp2
        t0, t1, t2, t3 = v7, v5, v3, v1
        z1, z2, z3, z4 = t0 + t3, t1 + t2, t0 + t2, t1 + t3
        z5 = (z3 + z4) * 9633
        t0, t1, t2, t3 = t0 * 2446, t1 * 16819, t2 * 25172, t3 * 12299
        z1, z2, z3, z4 = z1 * -7373, z2 * -20995, z3 * -16069 + z5, z4 * -3196 + z5
        t0, t1, t2, t3 = t0 + z1 + z3, t1 + z2 + z4, t2 + z2 + z3, t3 + z1 + z4
        return [_descale(a, shift) for a in (tmp10 + t3, tmp11 + t2, tmp12 + t1, tmp13 + t0, tmp13 - t0, tmp12 - t1, tmp11 - t2,
                                             tmp10 - t3)]

    cols = one_d(*[x[..., r, :] for r in range(8)], 13 - 2)                    # pass 1: columns
    ws = np.stack(cols, axis=-2)                                               # [..., row, col]
    rows = one_d(*[ws[..., :, c] for c in range(8)], 13 + 2 + 3)               # pass 2: rows
    out = np.stack(rows, axis=-1)
    idx = out & 1023
    res = np.where(idx < 128, idx + 128, np.where(idx < 512, 255, np.where(idx < 896, 0, idx - 896)))
    return res.astype(np.uint8)


def component_planes(coefs, info, geom):
    """IDCT of every block -> one uint8 plane per component (padded to whole MCUs)."""
    planes = []
    for ci, c in enumerate(info["frame"]["comps"]):
        q = info["qt"][c["tq"]]
        s = idct_islow(coefs[ci], q)                                            # [by, bx, 8, 8]
        by, bx = s.shape[:2]
        planes.append(s.transpose(0, 2, 1, 3).reshape(by * 8, bx * 8))
    return planes


def _ceil_div(a, b):
    return -(-a // b)


def upsample(plane, c, info, geom):
    """jdsample.c for one component -> full-resolution plane [H][>= W]."""
    f = info["frame"]
    _, _, hmax, vmax = geom
    H, W = f["H"], f["W"]
    dw, dh = _ceil_div(W * c["h"], hmax), _ceil_div(H * c["v"], vmax)          # downsampled_width / height
    p = plane[:dh, :dw].astype(np.int32)
    hx, vx = hmax // c["h"], vmax // c["v"]
    if hx == 1 and vx == 1:
        return p.astype(np.uint8)
    fancy = dw > 2                                                              # jinit_upsampler: do_fancy && downsampled_width > 2
    if hx == 2 and vx == 1:
        if not fancy:
            return np.repeat(p, 2, axis=1).astype(np.uint8)
        out = np.zeros((dh, 2 * dw), np.int32)
        left = np.concatenate([p[:, :1], p[:, :-1]], axis=1)
        right = np.concatenate([p[:, 1:], p[:, -1:]], axis=1)
        out[:, 0::2] = (3 * p + left + 1) >> 2
        out[:, 1::2] = (3 * p + right + 2) >> 2
        out[:, 0] = p[:, 0]
        out[:, -1] = p[:, -1]
        return out.astype(np.uint8)
    if hx == 1 and vx == 2:                                                     # h1v2_fancy_upsample (4:4:0)
        if not fancy:
            return np.repeat(p, 2, axis=0).astype(np.uint8)
        up = np.concatenate([p[:1], p[:-1]], axis=0)
        dn = np.concatenate([p[1:], p[-1:]], axis=0)
        out = np.zeros((2 * dh, dw), np.int32)
        out[0::2] = (3 * p + up + 1) >> 2
        out[1::2] = (3 * p + dn + 2) >> 2
        return out.astype(np.uint8)
    if hx == 2 and vx == 2:
        if not fancy:
            return np.repeat(np.repeat(p, 2, axis=0), 2, axis=1).astype(np.uint8)
        up = np.concatenate([p[:1], p[:-1]], axis=0)                            # jdmainct.c: context rows replicate the edge rows
        dn = np.concatenate([p[1:], p[-1:]], axis=0)
        out = np.zeros((2 * dh, 2 * dw), np.int32)
        for v, nb in ((0, up), (1, dn)):
            cs = 3 * p + nb                                                     # thiscolsum
            last = np.concatenate([cs[:, :1], cs[:, :-1]], axis=1)
            nxt = np.concatenate([cs[:, 1:], cs[:, -1:]], axis=1)
            even = (3 * cs + last + 8) >> 4
            odd = (3 * cs + nxt + 7) >> 4
            even[:, 0] = (4 * cs[:, 0] + 8) >> 4
            odd[:, -1] = (4 * cs[:, -1] + 7) >> 4
            out[v::2, 0::2] = even
            out[v::2, 1::2] = odd
        return out.astype(np.uint8)
    raise Unsupported(f"sampling factors {c['h']}x{c['v']} of {hmax}x{vmax}")


def ycc_to_bgr(y, cb, cr):
    """jdcolor.c::build_ycc_rgb_table + ycc_rgb_convert."""
    i = np.arange(256, dtype=np.int64) - 128
    fix = lambda v: int(v * 65536 + 0.5)
    cr_r = (fix(1.40200) * i + 32768) >> 16
    cb_b = (fix(1.77200) * i + 32768) >> 16
    cr_g = -fix(0.71414) * i
    cb_g = -fix(0.34414) * i + 32768
    yy = y.astype(np.int64)
    r = np.clip(yy + cr_r[cr], 0, 255)
    g = np.clip(yy + ((cb_g[cb] + cr_g[cr]) >> 16), 0, 255)
    b = np.clip(yy + cb_b[cb], 0, 255)
    return np.stack([b, g, r], axis=-1).astype(np.uint8)


def exif_transform(img, orientation):
    """OpenCV loadsave.cpp::ExifTransform."""
    t = lambda a: np.swapaxes(a, 0, 1)
    if orientation == 2:
        img = img[:, ::-1]
    elif orientation == 3:
        img = img[::-1, ::-1]
    elif orientation == 4:
        img = img[::-1]
    elif orientation == 5:
        img = t(img)
    elif orientation == 6:
        img = t(img)[:, ::-1]
    elif orientation == 7:
        img = t(img)[::-1, ::-1]
    elif orientation == 8:
        img = t(img)[::-1]
    return np.ascontiguousarray(img)


def imdecode(data: bytes, grayscale: bool = False, ignore_orientation: bool = False):
    """cv2.imdecode(data, IMREAD_COLOR | IMREAD_GRAYSCALE) for a baseline JPEG -> HxWx3 BGR or HxW."""
    info = parse(data)
    coefs, geom = decode_coefficients(data, info)
    f = info["frame"]
    H, W = f["H"], f["W"]
    comps = f["comps"]
    if grayscale or len(comps) == 1:
        planes = component_planes(coefs[:1], {**info, "frame": {**f, "comps": comps[:1]}}, geom)
        y = upsample(planes[0], comps[0], info, geom)[:H, :W]
        out = y if grayscale else np.stack([y, y, y], axis=-1)
    else:
        if len(comps) != 3:
            raise Unsupported("CMYK / YCCK JPEG")
        planes = component_planes(coefs, info, geom)
        full = [upsample(p, c, info, geom)[:H, :W] for p, c in zip(planes, comps)]
        out = ycc_to_bgr(*full)
    if not ignore_orientation:
        out = exif_transform(out, info["orientation"])
    return np.ascontiguousarray(out)
