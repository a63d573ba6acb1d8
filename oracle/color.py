"""NumPy restatement of the dual-fisheye input colour pipeline (TEST INFRASTRUCTURE -- see
oracle/__init__.py): cli_tools/gs360_DualFisheyeDistortionCalibration.py:568-725.

Pinned against outputs of the reference's own ``apply_input_color_pipeline`` recorded by
tests/golden/make_golden.py (``color_pipeline.npz``): bit-exact."""

from __future__ import annotations

import numpy as np

F32 = np.float32


def to_float01(image: np.ndarray) -> np.ndarray:
    """DF:599-609."""
    if image.dtype == np.uint8:
        return image.astype(F32) / F32(255.0)
    if image.dtype == np.uint16:
        return image.astype(F32) / F32(65535.0)
    if np.issubdtype(image.dtype, np.floating):
        return np.clip(image.astype(F32), F32(0), F32(1))
    raise TypeError("unsupported dtype %s" % image.dtype)


def from_float01(values: np.ndarray, dtype) -> np.ndarray:
    """DF:612-622 (round half to even)."""
    v = np.clip(values.astype(F32), F32(0), F32(1))
    if dtype == np.uint8:
        return np.rint(v * F32(255.0)).astype(np.uint8)
    if dtype == np.uint16:
        return np.rint(v * F32(65535.0)).astype(np.uint16)
    return v.astype(dtype)


def rec709_to_srgb(values: np.ndarray) -> np.ndarray:
    """DF:568-596: inverse Rec.709 OETF (knee 0.081, 4.5, 0.099 / 1.099, 1 / 0.45) then the sRGB
    OETF (0.0031308, 12.92, 1.055, 1 / 2.4), float32."""
    v = np.clip(values.astype(F32), F32(0), F32(1))
    lin = np.where(v < F32(0.081), v / F32(4.5),
                   np.power((v + F32(0.099)) / F32(1.099), F32(1.0 / 0.45)).astype(F32)).astype(F32)
    lin = np.clip(lin, F32(0), F32(1))
    enc = np.where(lin <= F32(0.0031308), F32(12.92) * lin,
                   (F32(1.055) * np.power(lin, F32(1.0 / 2.4)) - F32(0.055)).astype(F32)).astype(F32)
    return np.clip(enc, F32(0), F32(1))


def lut_trilinear(rgb: np.ndarray, table: np.ndarray, domain_min, domain_max) -> np.ndarray:
    """DF:625-681.  ``table`` is [b, g, r, 3] (a .cube file lists red fastest)."""
    flat = rgb.reshape(-1, 3).astype(F32)
    dmin = np.asarray(domain_min, dtype=F32).reshape(1, 3)
    span = np.asarray(domain_max, dtype=F32).reshape(1, 3) - dmin
    top = table.shape[0] - 1
    pos = np.clip((flat - dmin) / span, F32(0), F32(1)) * F32(top)
    i0 = np.floor(pos).astype(np.int32)
    i1 = np.minimum(i0 + 1, top)
    fr = pos - i0.astype(F32)
    f = [fr[:, c:c + 1] for c in range(3)]

    def node(bi, gi, ri):
        return table[(i1 if bi else i0)[:, 2], (i1 if gi else i0)[:, 1], (i1 if ri else i0)[:, 0]]

    def mix(a, b, t):
        return a + (b - a) * t

    lo = mix(mix(node(0, 0, 0), node(0, 0, 1), f[0]), mix(node(0, 1, 0), node(0, 1, 1), f[0]), f[1])
    hi = mix(mix(node(1, 0, 0), node(1, 0, 1), f[0]), mix(node(1, 1, 0), node(1, 1, 1), f[0]), f[1])
    return mix(lo, hi, f[2]).astype(F32).reshape(rgb.shape)


def apply_pipeline(image: np.ndarray, table: np.ndarray, domain_min, domain_max, output_space: str,
                   channel_order: str = "bgr") -> np.ndarray:
    """DF:684-725 for an HWC image whose first three channels are B, G, R (cv2) or R, G, B."""
    first3 = image[..., :3]
    rgb = first3[..., ::-1] if channel_order == "bgr" else first3
    val = lut_trilinear(to_float01(rgb), table, domain_min, domain_max)
    if output_space == "srgb":
        val = rec709_to_srgb(val)
    elif output_space == "passthrough":
        val = np.clip(val, F32(0), F32(1))
    else:
        raise ValueError(output_space)
    res = from_float01(val, image.dtype)
    res = res[..., ::-1] if channel_order == "bgr" else res
    if image.shape[2] == 3:
        return np.ascontiguousarray(res)
    out = image.copy()
    out[..., :3] = res
    return out


# ---- video colour step (PC:299-309): formula-level statement, float64 --------------------------------
# ffmpeg's `colorspace` filter is third-party, unversioned and absent from this image;
# this states what the filter computes on R'G'B' values (ITU-R BT.709 / SMPTE 170M / IEC 61966-2-1 curves,
# D65 primaries change) -- parity unpinned against ffmpeg itself, like the rest of that boundary.

_BT709_ALPHA, _BT709_BETA = 1.09929682680944, 0.018053968510807


def trc_decode64(name: str, v: np.ndarray) -> np.ndarray:
    a = np.abs(np.asarray(v, dtype=np.float64))
    if name in ("bt709", "smpte170m"):
        lin = np.where(a < 4.5 * _BT709_BETA, a / 4.5, ((a + (_BT709_ALPHA - 1.0)) / _BT709_ALPHA) ** (1.0 / 0.45))
    elif name in ("srgb", "iec61966-2-1"):
        lin = np.where(a <= 0.04045, a / 12.92, ((a + 0.055) / 1.055) ** 2.4)
    else:
        lin = a
    return np.copysign(lin, v)


def trc_encode64(name: str, lin: np.ndarray) -> np.ndarray:
    a = np.abs(np.asarray(lin, dtype=np.float64))
    if name in ("bt709", "smpte170m"):
        v = np.where(a < _BT709_BETA, 4.5 * a, _BT709_ALPHA * a ** 0.45 - (_BT709_ALPHA - 1.0))
    elif name in ("srgb", "iec61966-2-1"):
        v = np.where(a <= 0.0031308, 12.92 * a, 1.055 * a ** (1.0 / 2.4) - 0.055)
    else:
        v = a
    return np.copysign(v, lin)


def xyz_from_rgb64(primaries) -> np.ndarray:
    """primaries = ((xr, yr), (xg, yg), (xb, yb)); D65 white."""
    cols = np.array([[x / y, 1.0, (1.0 - x - y) / y] for x, y in primaries], dtype=np.float64).T
    white = np.array([0.3127 / 0.3290, 1.0, (1.0 - 0.3127 - 0.3290) / 0.3290])
    return cols * np.linalg.solve(cols, white)


BT709_PRIMARIES = ((0.640, 0.330), (0.300, 0.600), (0.150, 0.060))
SMPTE170M_PRIMARIES = ((0.630, 0.340), (0.310, 0.595), (0.155, 0.070))


def video_color_step64(image_rgb: np.ndarray, out_trc: str = "iec61966-2-1") -> np.ndarray:
    """R'G'B' image (RGB order, uint8 / uint16 / float) -> same dtype: BT.709 decode, 709 -> 170M primaries,
    encode with ``out_trc`` (``smpte170m`` when --keep-rec709), clip, round half to even."""
    dt = image_rgb.dtype
    scale = 255.0 if dt == np.uint8 else 65535.0 if dt == np.uint16 else 1.0
    v = image_rgb[..., :3].astype(np.float64) / scale
    if scale == 1.0:
        v = np.clip(v, 0.0, 1.0)
    m = np.linalg.solve(xyz_from_rgb64(SMPTE170M_PRIMARIES), xyz_from_rgb64(BT709_PRIMARIES))
    lin = trc_decode64("bt709", v) @ m.T
    enc = np.clip(trc_encode64(out_trc, lin), 0.0, 1.0)
    out = image_rgb.copy()
    if scale == 1.0:
        out[..., :3] = enc.astype(dt)
    else:
        out[..., :3] = np.rint(enc * scale).astype(dt)
    return out
