"""CPU restatement of the reference's per-frame loading step: JPEG decode + resize.

TEST INFRASTRUCTURE -- imported only by tests/, oracle/make_golden_frames.py and bench.py's cpu legs; the product path
(tuber_b200.frames / csrc/jpeg_host.cpp + csrc/frames.cu) never touches it.

The reference loads a clip as (datasets/ava_frame.py:146-150)

    tmp = Image.open(video_frame_list[frame_idx])            # JPEG decode: Pillow -> libjpeg(-turbo), all defaults
    tmp = tmp.resize((target['orig_size'][1], target['orig_size'][0]))   # Pillow's default filter for RGB: BICUBIC

so the arithmetic lives in two third-party libraries that are not under /root/reference (nothing is pinned; this container has
Pillow 12.2.0 on libjpeg-turbo, API level 6.2).  Both are integer algorithms with published, stable definitions, restated here:

* baseline sequential JPEG (ITU T.81): marker parsing, Huffman entropy decoding, restart intervals      -> decode_coefficients
* libjpeg's default decompression path: dequantisation + the "islow" integer inverse DCT (jidctint.c), "fancy" (triangle
  filter) chroma upsampling h2v1 / h2v2 (jdsample.c), fixed-point YCbCr -> RGB (jdcolor.c)              -> decode_jpeg
* Pillow's ImagingResample for 8-bit images (src/libImaging/Resample.c): double-precision filter weights normalised and
  converted to 22-bit fixed point, horizontal then vertical pass, each rounded and clipped to 8 bits     -> resize_bicubic

Pin: oracle/make_golden_frames.py encodes seeded synthetic images with Pillow at several sizes / qualities / chroma
subsamplings / restart intervals, decodes and resizes them with Pillow itself and stores Pillow's pixels in
tests/golden/frames.npz; tests/test_frames_cpu.py holds this file to those pixels BIT FOR BIT.
"""
from __future__ import annotations

import math
import struct
from typing import Dict, List, Tuple

import numpy as np

ZIGZAG = np.array([0, 1, 8, 16, 9, 2, 3, 10, 17, 24, 32, 25, 18, 11, 4, 5, 12, 19, 26, 33, 40, 48, 41, 34, 27, 20, 13, 6, 7, 14, 21, 28,
                   35, 42, 49, 56, 57, 50, 43, 36, 29, 22, 15, 23, 30, 37, 44, 51, 58, 59, 52, 45, 38, 31, 39, 46, 53, 60, 61, 54, 47, 55, 62,
                   63], dtype=np.int64)                    # zigzag position -> natural (row-major) index, T.81 figure A.6


# ------------------------------------------------------------------------------------------------------------------
# T.81: markers and Huffman decoding
# ------------------------------------------------------------------------------------------------------------------
class _Bits:
    """MSB-first bit reader over entropy-coded data (0xFF00 byte stuffing removed by the caller)."""

    def __init__(self, data: bytes):
        self.data, self.pos, self.acc, self.n = data, 0, 0, 0

    def get(self, k: int) -> int:
        while self.n < k:
            b = self.data[self.pos] if self.pos < len(self.data) else 0
            self.pos += 1
            self.acc = (self.acc << 8) | b
            self.n += 8
        self.n -= k
        v = (self.acc >> self.n) & ((1 << k) - 1)
        self.acc &= (1 << self.n) - 1
        return v


def _build_huff(counts: List[int], symbols: List[int]) -> Dict[Tuple[int, int], int]:
    table, code, k = {}, 0, 0
    for length in range(1, 17):
        for _ in range(counts[length - 1]):
            table[(length, code)] = symbols[k]
            code += 1
            k += 1
        code <<= 1
    return table


def _decode_symbol(bits: _Bits, table) -> int:
    code = 0
    for length in range(1, 17):
        code = (code << 1) | bits.get(1)
        s = table.get((length, code))
        if s is not None:
            return s
    raise ValueError("bad Huffman code")


def _extend(v: int, t: int) -> int:
    return v if v >= (1 << (t - 1)) else v - (1 << t) + 1    # T.81 F.2.2.1 EXTEND


def parse_jpeg(data: bytes) -> dict:
    """Markers of a baseline (SOF0) JPEG: frame header, quantisation and Huffman tables, restart interval, the scan."""
    assert data[0:2] == b"\xff\xd8", "not a JPEG"
    pos, info = 2, {"qt": {}, "dc": {}, "ac": {}, "ri": 0}
    while True:
        assert data[pos] == 0xFF, "marker expected"
        m = data[pos + 1]
        pos += 2
        if m == 0xD8 or (0xD0 <= m <= 0xD7) or m == 0x01 or m == 0xFF:
            if m == 0xFF:
                pos -= 1
            continue
        seglen = struct.unpack(">H", data[pos:pos + 2])[0]
        seg = data[pos + 2:pos + seglen]
        if m == 0xDB:                                        # DQT
            i = 0
            while i < len(seg):
                pq, tq = seg[i] >> 4, seg[i] & 15
                i += 1
                if pq:
                    vals = struct.unpack(">64H", seg[i:i + 128]); i += 128
                else:
                    vals = list(seg[i:i + 64]); i += 64
                q = np.zeros(64, dtype=np.int64)
                q[ZIGZAG] = np.array(vals, dtype=np.int64)   # stored in zigzag order
                info["qt"][tq] = q
        elif m == 0xC0:                                      # SOF0
            p, h, w, nc = seg[0], struct.unpack(">H", seg[1:3])[0], struct.unpack(">H", seg[3:5])[0], seg[5]
            assert p == 8, "8-bit samples only"
            info["height"], info["width"] = h, w
            info["comps"] = [{"id": seg[6 + 3 * c], "h": seg[7 + 3 * c] >> 4, "v": seg[7 + 3 * c] & 15, "tq": seg[8 + 3 * c]} for c in range(nc)]
        elif m in (0xC1, 0xC2, 0xC3, 0xC5, 0xC6, 0xC7, 0xC9, 0xCA, 0xCB, 0xCD, 0xCE, 0xCF):
            raise ValueError("only baseline sequential (SOF0) JPEGs are supported")
        elif m == 0xC4:                                      # DHT
            i = 0
            while i < len(seg):
                tc, th = seg[i] >> 4, seg[i] & 15
                counts = list(seg[i + 1:i + 17])
                n = sum(counts)
                symbols = list(seg[i + 17:i + 17 + n])
                i += 17 + n
                info["dc" if tc == 0 else "ac"][th] = _build_huff(counts, symbols)
        elif m == 0xDD:                                      # DRI
            info["ri"] = struct.unpack(">H", seg[0:2])[0]
        elif m == 0xDA:                                      # SOS
            ns = seg[0]
            info["scan"] = [{"id": seg[1 + 2 * c], "td": seg[2 + 2 * c] >> 4, "ta": seg[2 + 2 * c] & 15} for c in range(ns)]
            info["scan_start"] = pos + seglen
            return info
        pos += seglen


def decode_coefficients(data: bytes):
    """-> (info, coefs): coefs[c] int64 [blocks_y, blocks_x, 64] quantised coefficients in natural order (padded to whole MCUs)."""
    info = parse_jpeg(data)
    comps = info["comps"]
    assert len(info["scan"]) == len(comps), "single interleaved scan only"
    hmax, vmax = max(c["h"] for c in comps), max(c["v"] for c in comps)
    mcux, mcuy = -(-info["width"] // (8 * hmax)), -(-info["height"] // (8 * vmax))
    info["hmax"], info["vmax"], info["mcux"], info["mcuy"] = hmax, vmax, mcux, mcuy
    coefs = [np.zeros((mcuy * c["v"], mcux * c["h"], 64), dtype=np.int64) for c in comps]
    # entropy-coded segments: split at RSTn markers, remove byte stuffing
    raw, pos, segs, cur = data, info["scan_start"], [], bytearray()
    while pos < len(raw):
        b = raw[pos]
        if b == 0xFF:
            nb = raw[pos + 1]
            if nb == 0x00:
                cur.append(0xFF); pos += 2; continue
            if 0xD0 <= nb <= 0xD7:
                segs.append(bytes(cur)); cur = bytearray(); pos += 2; continue
            if nb == 0xFF:
                pos += 1; continue
            break                                            # EOI or another marker: end of the scan
        cur.append(b); pos += 1
    segs.append(bytes(cur))
    ri = info["ri"] if info["ri"] else mcux * mcuy
    mcu = 0
    for seg in segs:
        bits, pred = _Bits(seg), [0] * len(comps)
        for _ in range(ri):
            if mcu >= mcux * mcuy:
                break
            my, mx = divmod(mcu, mcux)
            for ci, c in enumerate(comps):
                sc = info["scan"][ci]
                dct, act = info["dc"][sc["td"]], info["ac"][sc["ta"]]
                for by in range(c["v"]):
                    for bx in range(c["h"]):
                        blk = coefs[ci][my * c["v"] + by, mx * c["h"] + bx]
                        t = _decode_symbol(bits, dct)
                        diff = _extend(bits.get(t), t) if t else 0
                        pred[ci] += diff
                        blk[0] = pred[ci]
                        k = 1
                        while k < 64:
                            rs = _decode_symbol(bits, act)
                            r, s = rs >> 4, rs & 15
                            if s == 0:
                                if r == 15:
                                    k += 16; continue
                                break
                            k += r
                            blk[ZIGZAG[k]] = _extend(bits.get(s), s)
                            k += 1
            mcu += 1
    return info, coefs


# ------------------------------------------------------------------------------------------------------------------
# libjpeg: islow inverse DCT (jidctint.c), fancy upsampling (jdsample.c), colour conversion (jdcolor.c)
# ------------------------------------------------------------------------------------------------------------------
CONST_BITS, PASS1_BITS = 13, 2
F_0_298631336, F_0_390180644, F_0_541196100, F_0_765366865 = 2446, 3196, 4433, 6270
F_0_899976223, F_1_175875602, F_1_501321110, F_1_847759065 = 7373, 9633, 12299, 15137
F_1_961570560, F_2_053119869, F_2_562915447, F_3_072711026 = 16069, 16819, 20995, 25172


def _descale(x, n):
    return (x + (1 << (n - 1))) >> n


def _idct_1d(d, shift_in_even):
    """one pass of jpeg_idct_islow over the last axis of d[..., 8] (int64); returns the 8 un-descaled outputs"""
    z2, z3 = d[..., 2], d[..., 6]
    z1 = (z2 + z3) * F_0_541196100
    tmp2 = z1 + z3 * (-F_1_847759065)
    tmp3 = z1 + z2 * F_0_765366865
    z2, z3 = d[..., 0], d[..., 4]
    tmp0 = (z2 + z3) << shift_in_even
    tmp1 = (z2 - z3) << shift_in_even
    tmp10, tmp13, tmp11, tmp12 = tmp0 + tmp3, tmp0 - tmp3, tmp1 + tmp2, tmp1 - tmp2
    tmp0, tmp1, tmp2, tmp3 = d[..., 7], d[..., 5], d[..., 3], d[..., 1]
    z1, z2, z3, z4 = tmp0 + tmp3, tmp1 + tmp2, tmp0 + tmp2, tmp1 + tmp3
    z5 = (z3 + z4) * F_1_175875602
    tmp0, tmp1, tmp2, tmp3 = tmp0 * F_0_298631336, tmp1 * F_2_053119869, tmp2 * F_3_072711026, tmp3 * F_1_501321110
    z1, z2, z3, z4 = z1 * (-F_0_899976223), z2 * (-F_2_562915447), z3 * (-F_1_961570560), z4 * (-F_0_390180644)
    z3, z4 = z3 + z5, z4 + z5
    tmp0, tmp1, tmp2, tmp3 = tmp0 + z1 + z3, tmp1 + z2 + z4, tmp2 + z2 + z3, tmp3 + z1 + z4
    return np.stack([tmp10 + tmp3, tmp11 + tmp2, tmp12 + tmp1, tmp13 + tmp0, tmp13 - tmp0, tmp12 - tmp1, tmp11 - tmp2, tmp10 - tmp3], axis=-1)


def range_limit(x):
    """libjpeg's sample range-limit table applied to x + CENTERJSAMPLE, indexed modulo 1024 like `& RANGE_MASK`"""
    i = np.asarray(x) & 1023
    return np.where(i < 128, i + 128, np.where(i < 512, 255, np.where(i < 896, 0, i - 896))).astype(np.uint8)


def idct_islow(coef: np.ndarray, quant: np.ndarray) -> np.ndarray:
    """coef [..., 64] quantised coefficients (natural order), quant [64] -> samples uint8 [..., 8, 8] (jidctint.c:jpeg_idct_islow)"""
    d = (coef * quant).reshape(coef.shape[:-1] + (8, 8)).astype(np.int64)
    # pass 1: columns (the transform runs along the row index), results scaled up by 2^PASS1_BITS
    ws = _descale(_idct_1d(np.swapaxes(d, -1, -2), CONST_BITS), CONST_BITS - PASS1_BITS)      # [..., col, row]
    ws = np.swapaxes(ws, -1, -2)                                                              # [..., row, col]
    # pass 2: rows
    out = _descale(_idct_1d(ws, CONST_BITS), CONST_BITS + PASS1_BITS + 3)
    return range_limit(out)


def _plane(coefs: np.ndarray, quant: np.ndarray) -> np.ndarray:
    s = idct_islow(coefs, quant)                             # [by, bx, 8, 8]
    by, bx = s.shape[:2]
    return s.transpose(0, 2, 1, 3).reshape(by * 8, bx * 8)


def _h2v1_fancy(p: np.ndarray) -> np.ndarray:
    """jdsample.c:h2v1_fancy_upsample on the real columns of every row: p [rows, w] -> [rows, 2w]"""
    p = p.astype(np.int64)
    w = p.shape[1]
    out = np.empty((p.shape[0], 2 * w), dtype=np.int64)
    prev = np.concatenate([p[:, :1], p[:, :-1]], axis=1)
    nxt = np.concatenate([p[:, 1:], p[:, -1:]], axis=1)
    out[:, 0::2] = (p * 3 + prev + 1) >> 2
    out[:, 1::2] = (p * 3 + nxt + 2) >> 2
    out[:, 0] = p[:, 0]
    out[:, -1] = p[:, -1]
    return out.astype(np.uint8)


def _h2v2_fancy(p: np.ndarray) -> np.ndarray:
    """jdsample.c:h2v2_fancy_upsample: p [h, w] (real rows / columns; the rows above the first and below the last are their
    copies, jdmainct.c) -> [2h, 2w]"""
    p = p.astype(np.int64)
    h, w = p.shape
    up = np.concatenate([p[:1], p[:-1]], axis=0)
    dn = np.concatenate([p[1:], p[-1:]], axis=0)
    out = np.empty((2 * h, 2 * w), dtype=np.int64)
    for v, other in ((0, up), (1, dn)):
        cs = p * 3 + other                                   # column sums: 3 * nearer row + further row
        last = np.concatenate([cs[:, :1], cs[:, :-1]], axis=1)
        nxt = np.concatenate([cs[:, 1:], cs[:, -1:]], axis=1)
        a = (cs * 3 + last + 8) >> 4
        b = (cs * 3 + nxt + 7) >> 4
        a[:, 0] = (cs[:, 0] * 4 + 8) >> 4
        b[:, -1] = (cs[:, -1] * 4 + 7) >> 4
        out[v::2, 0::2] = a
        out[v::2, 1::2] = b
    return out.astype(np.uint8)


_ONE_HALF = 1 << 15


def _fix(x: float) -> int:
    return int(x * 65536 + 0.5)


_X = np.arange(256, dtype=np.int64) - 128
CR_R = (_fix(1.40200) * _X + _ONE_HALF) >> 16
CB_B = (_fix(1.77200) * _X + _ONE_HALF) >> 16
CR_G = (-_fix(0.71414)) * _X
CB_G = (-_fix(0.34414)) * _X + _ONE_HALF


def ycc_to_rgb(y: np.ndarray, cb: np.ndarray, cr: np.ndarray) -> np.ndarray:
    """jdcolor.c:ycc_rgb_convert"""
    y = y.astype(np.int64)
    r = y + CR_R[cr]
    g = y + ((CB_G[cb] + CR_G[cr]) >> 16)
    b = y + CB_B[cb]
    return np.stack([range_limit(r - 128), range_limit(g - 128), range_limit(b - 128)], axis=-1)


def decode_jpeg(data: bytes) -> np.ndarray:
    """What `np.asarray(Image.open(io.BytesIO(data)))` gives for a baseline YCbCr JPEG: RGB uint8 [H, W, 3]."""
    info, coefs = decode_coefficients(data)
    comps = info["comps"]
    assert len(comps) == 3, "three-component (YCbCr) JPEGs only"
    H, W, hmax, vmax = info["height"], info["width"], info["hmax"], info["vmax"]
    planes = []
    for c, cf in zip(comps, coefs):
        p = _plane(cf, info["qt"][c["tq"]])
        dh, dw = -(-H * c["v"] // vmax), -(-W * c["h"] // hmax)    # downsampled_height / width: the real samples
        p = p[:dh, :dw]
        hx, vx = hmax // c["h"], vmax // c["v"]
        if (hx, vx) == (1, 1):
            pass
        elif (hx, vx) == (2, 1):
            p = _h2v1_fancy(p) if dw > 2 else np.repeat(p, 2, axis=1)
        elif (hx, vx) == (2, 2):
            p = _h2v2_fancy(p) if dw > 2 else np.repeat(np.repeat(p, 2, axis=0), 2, axis=1)
        else:
            raise ValueError("unsupported chroma subsampling %dx%d" % (hx, vx))
        planes.append(p[:H, :W])
    return ycc_to_rgb(*planes)


# ------------------------------------------------------------------------------------------------------------------
# Pillow: ImagingResample, 8 bits per channel, BICUBIC (src/libImaging/Resample.c)
# ------------------------------------------------------------------------------------------------------------------
PRECISION_BITS = 32 - 8 - 2


def _bicubic(x: float) -> float:
    a = -0.5
    if x < 0.0:
        x = -x
    if x < 1.0:
        return ((a + 2.0) * x - (a + 3.0)) * x * x + 1
    if x < 2.0:
        return (((x - 5) * x + 8) * x - 4) * a
    return 0.0


def resample_coeffs(in_size: int, out_size: int):
    """precompute_coeffs + normalize_coeffs_8bpc: -> (bounds [out, 2] (first tap, taps), coefficients int64 [out, ksize])"""
    scale = filterscale = in_size / out_size
    if filterscale < 1.0:
        filterscale = 1.0
    support = 2.0 * filterscale
    ksize = int(math.ceil(support)) * 2 + 1
    bounds = np.zeros((out_size, 2), dtype=np.int64)
    kk = np.zeros((out_size, ksize), dtype=np.int64)
    ss = 1.0 / filterscale
    for xx in range(out_size):
        center = 0 + (xx + 0.5) * scale
        xmin = int(center - support + 0.5)
        if xmin < 0:
            xmin = 0
        xmax = int(center + support + 0.5)
        if xmax > in_size:
            xmax = in_size
        xmax -= xmin
        k = [_bicubic((x + xmin - center + 0.5) * ss) for x in range(xmax)]
        ww = 0.0
        for w in k:
            ww += w
        if ww != 0.0:
            k = [w / ww for w in k]
        for x, w in enumerate(k):
            kk[xx, x] = int(-0.5 + w * (1 << PRECISION_BITS)) if w < 0 else int(0.5 + w * (1 << PRECISION_BITS))
        bounds[xx] = (xmin, xmax)
    return bounds, kk


def _clip8(v: np.ndarray) -> np.ndarray:
    return np.clip(v >> PRECISION_BITS, 0, 255).astype(np.uint8)


def _resample_axis0(img: np.ndarray, out_size: int) -> np.ndarray:
    """one pass along axis 0 of img [n, m, C] uint8"""
    bounds, kk = resample_coeffs(img.shape[0], out_size)
    src = img.astype(np.int64)
    out = np.empty((out_size,) + img.shape[1:], dtype=np.uint8)
    for i in range(out_size):
        x0, n = bounds[i]
        acc = np.tensordot(kk[i, :n], src[x0:x0 + n], axes=(0, 0)) + (1 << (PRECISION_BITS - 1))
        out[i] = _clip8(acc)
    return out


def resize_bicubic(img: np.ndarray, out_h: int, out_w: int) -> np.ndarray:
    """What `np.asarray(Image.fromarray(img).resize((out_w, out_h)))` gives (Pillow >= 7: BICUBIC by default for RGB):
    horizontal pass, then vertical pass, each rounded and clipped to 8 bits; a pass is skipped when the size does not change."""
    h, w = img.shape[:2]
    if (h, w) == (out_h, out_w):
        return img.copy()
    if w != out_w:
        img = _resample_axis0(img.transpose(1, 0, 2), out_w).transpose(1, 0, 2)
    if h != out_h:
        img = _resample_axis0(img, out_h)
    return np.ascontiguousarray(img)


def load_frame(data: bytes, out_h: int, out_w: int) -> np.ndarray:
    """datasets/ava_frame.py:148-149 for one frame: decode, resize to (out_h, out_w); RGB uint8 [out_h, out_w, 3]"""
    return resize_bicubic(decode_jpeg(data), out_h, out_w)
