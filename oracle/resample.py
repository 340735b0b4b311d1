"""Oracle: Pillow's 8-bit two-pass bicubic resample, restated in NumPy.  Test infrastructure only.

Restates the third-party routine reached from reference ``src/eval/infer.py:84``
(``image.resize(..., Image.BICUBIC)``) and HF ``image_transforms.py:368``:
Pillow ``ImagingResample`` for 8 bpc images (``precompute_coeffs``,
``normalize_coeffs_8bpc``, horizontal pass, vertical pass).  Pillow is not
vendored by the reference (``requirements.txt:26``, unpinned); 12.2.0 is what is
installed and what ``tests/test_oracle_pinning.py`` checks this file against.
Also ``crop_u8`` = Pillow ``Image.crop`` (zero fill outside the image),
reached from ``infer.py:72,75``.
"""
import math

import numpy as np

PRECISION_BITS = 32 - 8 - 2          # 22


def _bicubic(x):
    a = -0.5
    if x < 0.0:
        x = -x
    if x < 1.0:
        return ((a + 2.0) * x - (a + 3.0)) * x * x + 1
    if x < 2.0:
        return (((x - 5) * x + 8) * x - 4) * a
    return 0.0


def precompute_coeffs(in_size, in0, in1, out_size):
    """Per-axis taps: returns (ksize, bounds[out,2] int32 (xmin,count), kk[out,ksize] int32 fixed point)."""
    scale = (in1 - in0) / out_size
    filterscale = scale if scale >= 1.0 else 1.0
    support = 2.0 * filterscale
    ksize = int(math.ceil(support)) * 2 + 1
    bounds = np.zeros((out_size, 2), np.int32)
    kk = np.zeros((out_size, ksize), np.int32)
    ss = 1.0 / filterscale
    for xx in range(out_size):
        center = in0 + (xx + 0.5) * scale
        xmin = int(center - support + 0.5)
        if xmin < 0:
            xmin = 0
        xmax = int(center + support + 0.5)
        if xmax > in_size:
            xmax = in_size
        xmax -= xmin
        w = [0.0] * xmax
        ww = 0.0
        for x in range(xmax):
            v = _bicubic((x + xmin - center + 0.5) * ss)
            w[x] = v
            ww += v                                   # sequential, left to right
        for x in range(xmax):
            if ww != 0.0:
                w[x] /= ww
            v = w[x]
            kk[xx, x] = int(-0.5 + v * (1 << PRECISION_BITS)) if v < 0 else int(0.5 + v * (1 << PRECISION_BITS))
        bounds[xx] = (xmin, xmax)
    return ksize, bounds, kk


def _pass_axis1(src, bounds, kk, out_size):
    """Filter along axis 1 of src[(rows, in, C)] u8 -> (rows, out, C) u8 with 8bpc rounding."""
    rows, _, ch = src.shape
    out = np.empty((rows, out_size, ch), np.uint8)
    s64 = src.astype(np.int64)
    for xx in range(out_size):
        xmin, n = int(bounds[xx, 0]), int(bounds[xx, 1])
        acc = np.tensordot(s64[:, xmin:xmin + n, :], kk[xx, :n].astype(np.int64), axes=([1], [0]))
        acc = (acc + (1 << (PRECISION_BITS - 1))) >> PRECISION_BITS
        out[:, xx, :] = np.clip(acc, 0, 255).astype(np.uint8)
    return out


def resize_u8(src, out_w, out_h):
    """``Image.resize((out_w, out_h), BICUBIC)`` on an (H, W, C) uint8 array."""
    assert src.dtype == np.uint8 and src.ndim == 3
    in_h, in_w, _ = src.shape
    if out_w == in_w and out_h == in_h:
        return src.copy()
    need_h = out_w != in_w
    need_v = out_h != in_h
    tmp = src
    if need_h:
        _, bh, kh = precompute_coeffs(in_w, 0, in_w, out_w)
        if need_v:
            _, bv, kv = precompute_coeffs(in_h, 0, in_h, out_h)
            y0 = int(bv[0, 0])
            y1 = int(bv[out_h - 1, 0] + bv[out_h - 1, 1])
            # only the rows the vertical pass will read are filtered; shift the vertical bounds accordingly
            tmp = _pass_axis1(src[y0:y1], bh, kh, out_w)
            bv = bv.copy()
            bv[:, 0] -= y0
        else:
            tmp = _pass_axis1(src, bh, kh, out_w)
    elif need_v:
        _, bv, kv = precompute_coeffs(in_h, 0, in_h, out_h)
    if need_v:
        tmp = _pass_axis1(tmp.transpose(1, 0, 2), bv, kv, out_h).transpose(1, 0, 2)
    return np.ascontiguousarray(tmp)


def crop_u8(src, box):
    """``Image.crop(box)`` on (H, W, C) uint8: outside-image area is zero; right<left raises like Pillow."""
    x0, y0, x1, y1 = (int(v) for v in box)
    if x1 < x0:
        raise ValueError("Coordinate 'right' is less than 'left'")
    if y1 < y0:
        raise ValueError("Coordinate 'lower' is less than 'upper'")
    h, w, c = src.shape
    out = np.zeros((y1 - y0, x1 - x0, c), np.uint8)
    sx0, sy0, sx1, sy1 = max(x0, 0), max(y0, 0), min(x1, w), min(y1, h)
    if sx1 > sx0 and sy1 > sy0:
        out[sy0 - y0:sy1 - y0, sx0 - x0:sx1 - x0] = src[sy0:sy1, sx0:sx1]
    return out
