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
