"""NumPy model of ``cv2.remap`` (TEST INFRASTRUCTURE -- see oracle/__init__.py).

The reference resamples with
``cv2.remap(src, map_x, map_y, interp, borderMode=BORDER_CONSTANT, borderValue=v)``
(cli_tools/gs360_DualFisheyeDistortionCalibration.py:2001-2008, masks with
INTER_NEAREST at :2031-2038, interpolation names at :59-64).  OpenCV is a
third-party dependency (requirements.txt ``opencv-python``, 4.13.0 in this
image); its published algorithm (modules/imgproc/src/imgwarp.cpp: initInterTab1D,
initInterTab2D, remapNearest/remapBilinear/remapBicubic) is restated here:

* the float32 map is quantised to 1/32 px: ``s = cvRound(m * 32)``
  (round-half-even), integer part ``s >> 5`` saturated to int16, fraction
  ``s & 31``;
* 1-D weights for the 32 fractions: linear ``(1-t, t)``; cubic with A = -0.75,
  evaluated in float32 exactly as ``interpolateCubic``; lanczos4 (8 taps) as
  ``interpolateLanczos4`` (double sin/cos, float32 coefficients and normalisation);
* 2-D weight = float32 product ``wy * wx``.  uint8 uses 15-bit fixed-point
  weights ``saturate_cast<short>(w * 32768)``; when a table entry does not sum
  to 32768 the difference is folded into the largest (sum too small) or
  smallest (sum too large) weight among taps (k1, k2) in {ksize/2, ksize/2+1}^2;
  result ``(sum + 16384) >> 15`` saturated;
* uint16 / float32 accumulate float32 products left to right in tap order and
  uint16 rounds half-even with saturation;
* BORDER_CONSTANT applies per tap.

``border="erp"`` is this repo's panorama border (not an OpenCV mode): taps wrap
in x modulo W and clamp in y.  ``sample_cv2`` obtains the same thing from the
real ``cv2.remap`` by padding the source, which is how the model is pinned.
"""

from __future__ import annotations

import functools
import math
from typing import Tuple

import numpy as np

F32 = np.float32
INTER_BITS = 5
INTER_TAB = 1 << INTER_BITS          # 32
COEF_BITS = 15
COEF_ONE = 1 << COEF_BITS            # 32768

INTERPS = ("nearest", "linear", "cubic", "lanczos4")


def _cubic_coeffs(t: np.float32) -> np.ndarray:
    a = F32(-0.75)
    one = F32(1.0)
    x = F32(t)
    c0 = ((a * (x + one) - F32(5) * a) * (x + one) + F32(8) * a) * (x + one) - F32(4) * a
    c1 = ((a + F32(2)) * x - (a + F32(3))) * x * x + one
    xm = one - x
    c2 = ((a + F32(2)) * xm - (a + F32(3))) * xm * xm + one
    c3 = one - c0 - c1 - c2
    return np.array([c0, c1, c2, c3], dtype=F32)


def _lanczos4_coeffs(t: np.float32) -> np.ndarray:
    """interpolateLanczos4: a = 4 windowed sinc over 8 taps written with the angle-addition table
    cs[] (sin/cos in double, each coefficient rounded to float32, float32 normalisation);
    a zero fraction is the exact unit impulse."""
    x = F32(t)
    if x < np.finfo(F32).eps:
        return np.array([0, 0, 0, 1, 0, 0, 0, 0], dtype=F32)
    s45 = 0.70710678118654752440084436210485
    cs = ((1, 0), (-s45, -s45), (0, 1), (s45, -s45), (-1, 0), (s45, s45), (0, -1), (-s45, s45))
    y0 = -(float(x) + 3.0) * math.pi * 0.25
    s0, c0 = math.sin(y0), math.cos(y0)
    co = np.empty(8, dtype=F32)
    total = F32(0)
    for i in range(8):
        d = F32(x + F32(3) - F32(i))
        if abs(d) >= F32(1e-6):
            y = -float(d) * math.pi * 0.25
            co[i] = F32((cs[i][0] * s0 + cs[i][1] * c0) / (y * y))
        else:
            co[i] = F32(1e30)
        total = F32(total + co[i])
    return (co * F32(F32(1) / total)).astype(F32)


@functools.lru_cache(maxsize=None)
def tables(interp: str) -> Tuple[np.ndarray, np.ndarray, np.ndarray]:
    """(tab1d[32,k] f32, tab2d[32,32,k,k] f32, itab2d[32,32,k,k] int32)."""
    scale = F32(1.0) / F32(INTER_TAB)
    if interp == "linear":
        rows = [np.array([F32(1.0) - F32(i) * scale, F32(i) * scale], dtype=F32)
                for i in range(INTER_TAB)]
    elif interp == "cubic":
        rows = [_cubic_coeffs(F32(i) * scale) for i in range(INTER_TAB)]
    elif interp == "lanczos4":
        rows = [_lanczos4_coeffs(F32(i) * scale) for i in range(INTER_TAB)]
    else:
        raise ValueError(interp)
    t1 = np.stack(rows).astype(F32)
    k = t1.shape[1]
    tab = np.empty((INTER_TAB, INTER_TAB, k, k), dtype=F32)
    itab = np.empty((INTER_TAB, INTER_TAB, k, k), dtype=np.int32)
    half = k // 2
    for i in range(INTER_TAB):
        for j in range(INTER_TAB):
            w = (t1[i][:, None] * t1[j][None, :]).astype(F32)
            tab[i, j] = w
            iw = np.clip(np.rint(w * F32(COEF_ONE)), -32768, 32767).astype(np.int32)
            diff = int(iw.sum()) - COEF_ONE
            if diff != 0 and k == 2:
                # only entry (0, 0): 1.0 * 32768 saturates to 32767.  OpenCV's search
                # window (k1, k2 in {1, 2}) runs past a 2x2 entry; the unit lands on
                # tap (1, 1), which cannot change any 8-bit result.
                iw[1, 1] -= diff
            elif diff != 0:
                lo = hi = (half, half)
                for k1 in (half, half + 1):
                    for k2 in (half, half + 1):
                        if iw[k1, k2] < iw[lo]:
                            lo = (k1, k2)
                        elif iw[k1, k2] > iw[hi]:
                            hi = (k1, k2)
                if diff < 0:
                    iw[hi] -= diff
                else:
                    iw[lo] -= diff
            itab[i, j] = iw
    return t1, tab, itab


def quantise(m: np.ndarray) -> Tuple[np.ndarray, np.ndarray]:
    """float32 map -> (integer part saturated to int16 range, 5-bit fraction)."""
    s = np.rint(np.asarray(m, dtype=F32) * F32(INTER_TAB)).astype(np.int64)
    return np.clip(s >> INTER_BITS, -32768, 32767), s & (INTER_TAB - 1)


def _fetch(src: np.ndarray, yy: np.ndarray, xx: np.ndarray, border: str, fill):
    h, w = src.shape[:2]
    if border == "erp":
        return src[np.clip(yy, 0, h - 1), np.mod(xx, w)]
    inside = (yy >= 0) & (yy < h) & (xx >= 0) & (xx < w)
    vals = src[np.clip(yy, 0, h - 1), np.clip(xx, 0, w - 1)]
    return np.where(inside[..., None], vals, fill)


def _round_half_even_sat(acc: np.ndarray, dtype) -> np.ndarray:
    info = np.iinfo(dtype)
    return np.clip(np.rint(acc), info.min, info.max).astype(dtype)


def sample(src: np.ndarray, map_x: np.ndarray, map_y: np.ndarray, interp: str = "cubic",
           border: str = "constant", border_value: float = 0.0,
           out_dtype=None, row_block: int = 128) -> np.ndarray:
    """Model of cv2.remap for HWC (or HW) uint8 / uint16 / float32 / float16 sources.

    float16 has no OpenCV counterpart: it is sampled as float32 and rounded to
    the output type at the end.  ``out_dtype`` may be float16 for uint16 sources
    (value / 65535 in float32, then rounded to half).
    """
    if interp not in INTERPS:
        raise ValueError(interp)
    if border not in ("constant", "erp"):
        raise ValueError(border)
    squeeze = src.ndim == 2
    s3 = src[..., None] if squeeze else src
    in_dtype = s3.dtype
    out_dtype = np.dtype(out_dtype or in_dtype)
    mx = np.asarray(map_x, dtype=F32)
    my = np.asarray(map_y, dtype=F32)
    oh, ow = mx.shape
    out = np.empty((oh, ow, s3.shape[2]), dtype=out_dtype)
    for r0 in range(0, oh, row_block):
        sl = slice(r0, min(oh, r0 + row_block))
        out[sl] = _sample_block(s3, mx[sl], my[sl], interp, border, border_value, out_dtype)
    return out[..., 0] if squeeze else out


def _sample_block(s3, mx, my, interp, border, border_value, out_dtype):
    in_dtype = s3.dtype
    if interp == "nearest":
        ix = np.clip(np.rint(mx).astype(np.int64), -32768, 32767)
        iy = np.clip(np.rint(my).astype(np.int64), -32768, 32767)
        fill = np.asarray(border_value).astype(in_dtype)
        vals = _fetch(s3, iy, ix, border, fill)
        return _finish_float(vals, in_dtype, out_dtype) if vals.dtype != out_dtype else vals
    ix, fx = quantise(mx)
    iy, fy = quantise(my)
    t1, tab, itab = tables(interp)
    k = t1.shape[1]
    off = k // 2 - 1            # linear: 0, cubic: 1, lanczos4: 3
    if in_dtype == np.uint8:
        acc = np.zeros(mx.shape + (s3.shape[2],), dtype=np.int64)
        fill = np.int64(np.clip(np.rint(border_value), 0, 255))
        for k1 in range(k):
            for k2 in range(k):
                p = _fetch(s3, iy + (k1 - off), ix + (k2 - off), border, fill).astype(np.int64)
                acc += p * itab[fy, fx, k1, k2][..., None]
        res = np.clip((acc + (COEF_ONE >> 1)) >> COEF_BITS, 0, 255).astype(np.uint8)
        return res if out_dtype == np.uint8 else res.astype(out_dtype)
    # float-weight paths (uint16 / float32 / float16 sources)
    sf = s3.astype(F32)
    fill = F32(border_value)
    if k == 2:
        # remapBilinear: one left-to-right expression over the 4 taps
        acc = None
        for k1 in range(k):
            for k2 in range(k):
                p = _fetch(sf, iy + (k1 - off), ix + (k2 - off), border, fill)
                term = (p * tab[fy, fx, k1, k2][..., None]).astype(F32)
                acc = term if acc is None else (acc + term).astype(F32)
        return _finish_float(acc, in_dtype, out_dtype)
    # remapBicubic / remapLanczos4, all taps inside: each row summed left to right,
    # rows added to a running sum that starts at zero
    acc = np.zeros(mx.shape + (s3.shape[2],), dtype=F32)
    for k1 in range(k):
        row = None
        for k2 in range(k):
            p = _fetch(sf, iy + (k1 - off), ix + (k2 - off), border, fill)
            term = (p * tab[fy, fx, k1, k2][..., None]).astype(F32)
            row = term if row is None else (row + term).astype(F32)
        acc = (acc + row).astype(F32)
    if border == "constant":
        # remapBicubic near the border: sum = cval; sum += (tap - cval) * w for
        # the taps that exist, in tap order
        h, w = s3.shape[:2]
        edge = ~((ix - off >= 0) & (ix - off < max(w - (k - 1), 0)) &
                 (iy - off >= 0) & (iy - off < max(h - (k - 1), 0)))
        if edge.any():
            ey, ex = np.nonzero(edge)
            e_acc = np.full((ey.size, s3.shape[2]), fill, dtype=F32)
            for k1 in range(k):
                for k2 in range(k):
                    yy = iy[ey, ex] + (k1 - off)
                    xx = ix[ey, ex] + (k2 - off)
                    inside = (yy >= 0) & (yy < h) & (xx >= 0) & (xx < w)
                    p = sf[np.clip(yy, 0, h - 1), np.clip(xx, 0, w - 1)]
                    term = ((p - fill).astype(F32) * tab[fy[ey, ex], fx[ey, ex], k1, k2][:, None]).astype(F32)
                    e_acc = np.where(inside[:, None], (e_acc + term).astype(F32), e_acc)
            acc[ey, ex] = e_acc
    return _finish_float(acc, in_dtype, out_dtype)


def _finish_float(acc, in_dtype, out_dtype):
    if out_dtype == np.uint16:
        return _round_half_even_sat(acc, np.uint16)
    if out_dtype == np.uint8:
        return _round_half_even_sat(acc, np.uint8)
    if out_dtype == np.float16:
        if in_dtype == np.uint16:
            acc = (acc.astype(F32) * F32(1.0 / 65535.0)).astype(F32)
        return acc.astype(np.float16)
    return acc.astype(out_dtype)


def apply_invalid_fill(img: np.ndarray, valid: np.ndarray, fill_value) -> np.ndarray:
    """``rendered[~valid] = mask_value`` (DF:2009-2014)."""
    img = img.copy()
    img[~valid] = fill_value
    return img


# --------------------------------------------------------------------------
# the real thing, used to pin the model (cv2 is present in this image and on
# the GPU box; it is the library the reference itself calls)
# --------------------------------------------------------------------------

def sample_cv2(src: np.ndarray, map_x: np.ndarray, map_y: np.ndarray, interp: str = "cubic",
               border: str = "constant", border_value: float = 0.0) -> np.ndarray:
    import cv2
    flag = {"nearest": cv2.INTER_NEAREST, "linear": cv2.INTER_LINEAR,
            "cubic": cv2.INTER_CUBIC, "lanczos4": cv2.INTER_LANCZOS4}[interp]
    mx = np.ascontiguousarray(map_x, dtype=F32)
    my = np.ascontiguousarray(map_y, dtype=F32)
    nch = 1 if src.ndim == 2 else src.shape[2]
    bv = (float(border_value),) * max(nch, 1)
    if border == "constant":
        return cv2.remap(src, mx, my, flag, borderMode=cv2.BORDER_CONSTANT, borderValue=bv)
    # panorama border: wrap columns, replicate rows, then no tap can leave the image
    pad = 8
    h, w = src.shape[:2]
    padded = np.concatenate([src[:, w - pad:], src, src[:, :pad]], axis=1)
    padded = np.concatenate([padded[:1].repeat(pad, 0), padded, padded[-1:].repeat(pad, 0)], axis=0)
    if float(mx.min()) < -1.5 or float(mx.max()) > w + 0.5:
        raise ValueError("erp border expects x within one pixel of [0, W)")
    # the float32 add is exact for the model only when it does not re-round, so
    # quantise first and hand cv2 already-quantised coordinates (multiples of 1/32
    # are exact in float32 at these magnitudes)
    qx = np.rint(mx * F32(32)).astype(np.int64)
    qy = np.rint(my * F32(32)).astype(np.int64)
    if int(qy.min()) < -48 or int(qy.max()) > (h - 1) * 32 + 48:
        raise ValueError("erp border expects y within 1.5 pixels of [0, H-1]")
    px = ((qx + pad * 32).astype(np.float64) / 32.0).astype(F32)
    py = ((qy + pad * 32).astype(np.float64) / 32.0).astype(F32)
    if interp == "nearest":
        px = (np.rint(mx).astype(np.int64) + pad).astype(F32)
        py = (np.clip(np.rint(my).astype(np.int64), 0, h - 1) + pad).astype(F32)
    return cv2.remap(np.ascontiguousarray(padded), px, py, flag,
                     borderMode=cv2.BORDER_CONSTANT, borderValue=bv)
