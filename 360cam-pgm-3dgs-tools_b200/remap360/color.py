"""Input colour pipeline of the dual-fisheye stage on the GPU: ``.cube`` 3-D LUT (trilinear) and the
optional Rec.709 -> sRGB re-encoding that the reference applies to every lens image right after
reading it (cli_tools/gs360_DualFisheyeDistortionCalibration.py:494-725; GUI default
``use_input_lut=True``).  The file parser runs on the host; the per-pixel work is the
``r360_apply_lut`` kernel."""

from __future__ import annotations

import ctypes
import pathlib
from dataclasses import dataclass
from typing import Optional, Tuple

import torch

from . import _lib
from .api import _describe, _stream_handle

OUTPUT_SPACES = {"passthrough": 0, "srgb": 1}


def normalize_lut_output_color_space(value) -> str:
    """DF:481-491 (``native`` is a legacy alias of ``passthrough``)."""
    text = str(value or "passthrough").strip().lower()
    if text == "native":
        return "passthrough"
    if text in OUTPUT_SPACES:
        return text
    raise ValueError("Unsupported --lut-output-color-space: {}".format(value))


@dataclass
class CubeLUT:
    """DF:99-105, plus the device copy the kernel reads (float4 per node, red fastest)."""
    size: int
    table: "object"                       # numpy float32 [size, size, size, 3], indexed [b, g, r]
    domain_min: Tuple[float, float, float]
    domain_max: Tuple[float, float, float]
    _device: Optional[torch.Tensor] = None

    def device_table(self, device) -> torch.Tensor:
        device = torch.device(device)
        if self._device is None or self._device.device != device:
            nodes = torch.from_numpy(self.table).reshape(-1, 3)
            padded = torch.zeros((nodes.shape[0], 4), dtype=torch.float32)
            padded[:, :3] = nodes
            self._device = padded.to(device)
        return self._device


def load_cube_lut(lut_path) -> CubeLUT:
    """Parse a ``.cube`` file the way DF:494-565 does: TITLE / comments skipped, LUT_3D_SIZE,
    DOMAIN_MIN / DOMAIN_MAX, then size^3 rows of three floats; same error messages."""
    import numpy as np
    lut_path = pathlib.Path(lut_path)
    if not lut_path.is_file():
        raise FileNotFoundError("LUT file not found: {}".format(lut_path))
    size, lo, hi, rows = None, [0.0, 0.0, 0.0], [1.0, 1.0, 1.0], []
    with lut_path.open("r", encoding="utf-8", errors="ignore") as fh:
        for raw in fh:
            line = raw.strip()
            if not line or line.startswith("#"):
                continue
            key = line.upper()
            parts = line.split()
            if key.startswith("TITLE"):
                continue
            if key.startswith("LUT_3D_SIZE"):
                if len(parts) < 2:
                    raise ValueError("Invalid LUT_3D_SIZE line: {}".format(line))
                size = int(parts[1])
            elif key.startswith("DOMAIN_MIN") or key.startswith("DOMAIN_MAX"):
                if len(parts) != 4:
                    raise ValueError("Invalid {} line: {}".format(key[:10], line))
                (lo if key.startswith("DOMAIN_MIN") else hi)[:] = [float(t) for t in parts[1:]]
            elif len(parts) == 3:
                rows.append((float(parts[0]), float(parts[1]), float(parts[2])))
    if size is None:
        raise ValueError("LUT_3D_SIZE is missing in {}".format(lut_path))
    if size <= 1:
        raise ValueError("LUT_3D_SIZE must be > 1 in {}".format(lut_path))
    if len(rows) != size ** 3:
        raise ValueError("LUT row count mismatch in {}: got {}, expected {}".format(lut_path, len(rows), size ** 3))
    dmin, dmax = np.asarray(lo, dtype=np.float32), np.asarray(hi, dtype=np.float32)
    if np.any(dmax - dmin <= 0.0):
        raise ValueError("Invalid LUT domain range in {}".format(lut_path))
    table = np.asarray(rows, dtype=np.float32).reshape((size, size, size, 3))
    return CubeLUT(size=size, table=table, domain_min=tuple(float(v) for v in dmin),
                   domain_max=tuple(float(v) for v in dmax))


def apply_input_color_pipeline(images: torch.Tensor, lut: Optional[CubeLUT], lut_output_color_space: str = "srgb",
                               *, channel_order: str = "bgr", out: Optional[torch.Tensor] = None,
                               stream: Optional[torch.cuda.Stream] = None) -> torch.Tensor:
    """[N, H, W, C] (or [..., H, W, C]) uint8 / uint16 / float32 CUDA images -> same shape and dtype.
    ``lut is None`` returns the input untouched, like DF:691-692.  ``out`` may alias ``images``."""
    if lut is None:
        return images
    space = normalize_lut_output_color_space(lut_output_color_space)
    if channel_order not in ("bgr", "rgb"):
        raise ValueError("channel_order must be 'bgr' or 'rgb'")
    if images.dim() < 3 or images.shape[-1] < 3:
        raise ValueError("LUT-based input conversion requires at least 3-channel RGB image input")
    if not images.is_contiguous():
        raise ValueError("images must be contiguous")
    flat = images.reshape((-1,) + tuple(images.shape[-3:]))
    if out is None:
        out = torch.empty_like(images)
    elif out.shape != images.shape or out.dtype != images.dtype or not out.is_contiguous():
        raise ValueError("out must match images")
    src, dst = _describe(flat, "images"), _describe(out.reshape(flat.shape), "out")
    table = lut.device_table(images.device)
    desc = _lib.Lut3D(table.data_ptr(), int(lut.size), 0, (ctypes.c_float * 3)(*lut.domain_min),
                      (ctypes.c_float * 3)(*lut.domain_max))
    with torch.cuda.device(images.device):
        _lib.check(_lib.load().r360_apply_lut(ctypes.byref(src), ctypes.byref(dst), ctypes.byref(desc),
                                              OUTPUT_SPACES[space], 1 if channel_order == "rgb" else 0,
                                              _stream_handle(stream, images.device)))
    if stream is not None:
        table.record_stream(stream)
    return out


# ---- video colour step ----------------------------------------------------------------------------------
# `colorspace=iall=bt709:all=smpte170m[:trc=iec61966-2-1]` (PC:299-309, V2F:462-464).

# CIE xy chromaticities of the primaries (R, G, B) and D65 white, as in ITU-R BT.709 / SMPTE 170M
_PRIMARIES = {
    "bt709": ((0.640, 0.330), (0.300, 0.600), (0.150, 0.060)),
    "smpte170m": ((0.630, 0.340), (0.310, 0.595), (0.155, 0.070)),
}
_D65 = (0.3127, 0.3290)


def rgb_to_xyz_matrix(primaries: str):
    """Linear RGB -> CIE XYZ for a set of primaries with D65 white (float64)."""
    import numpy as np
    xy = _PRIMARIES[primaries]
    cols = np.array([[x / y, 1.0, (1.0 - x - y) / y] for x, y in xy], dtype=np.float64).T
    white = np.array([_D65[0] / _D65[1], 1.0, (1.0 - _D65[0] - _D65[1]) / _D65[1]])
    return cols * np.linalg.solve(cols, white)


def primaries_matrix(src: str, dst: str):
    """Linear RGB(src primaries) -> linear RGB(dst primaries); both D65, so no chromatic adaptation."""
    import numpy as np
    if src == dst:
        return np.eye(3)
    return np.linalg.solve(rgb_to_xyz_matrix(dst), rgb_to_xyz_matrix(src))


def parse_colorspace_filter(text: str):
    """``colorspace=iall=bt709:all=smpte170m[:trc=iec61966-2-1][:range=..][:format=..]`` -> (in_primaries, in_trc,
    out_primaries, out_trc).  ``iall`` / ``all`` set primaries, transfer and matrix together; ``trc`` overrides
    the output transfer.  ``range`` and ``format`` concern the YUV representation only and have no R'G'B' effect."""
    if not text.startswith("colorspace="):
        raise ValueError("not a colorspace filter: %r" % text)
    opts = dict(kv.split("=", 1) for kv in text[len("colorspace="):].split(":") if "=" in kv)
    src, dst = opts.get("iall", "bt709"), opts.get("all", "bt709")
    for name in (src, dst):
        if name not in _PRIMARIES:
            raise ValueError("unsupported colour space %r" % name)
    out_trc = opts.get("trc", dst)
    if out_trc not in _lib.TRC:
        raise ValueError("unsupported transfer %r" % out_trc)
    return src, src, dst, out_trc


def convert_video_color(images: torch.Tensor, *, keep_rec709: bool = False, filter_text: Optional[str] = None,
                        channel_order: str = "bgr", out: Optional[torch.Tensor] = None,
                        stream: Optional[torch.cuda.Stream] = None) -> torch.Tensor:
    """The cutter's video colour step on [.., H, W, C] uint8 / uint16 / float32 CUDA frames (``out`` may alias
    ``images``): BT.709 -> SMPTE 170M primaries, re-encoded as sRGB unless ``keep_rec709`` (PC:303-305).
    ``filter_text`` takes the job's own ``colorspace=...`` string instead."""
    if filter_text is None:
        filter_text = "colorspace=iall=bt709:all=smpte170m" + ("" if keep_rec709 else ":trc=iec61966-2-1")
    src_p, in_trc, dst_p, out_trc = parse_colorspace_filter(filter_text)
    if channel_order not in ("bgr", "rgb"):
        raise ValueError("channel_order must be 'bgr' or 'rgb'")
    if images.dim() < 3 or images.shape[-1] < 3:
        raise ValueError("colour conversion requires at least 3-channel input")
    if not images.is_contiguous():
        raise ValueError("images must be contiguous")
    flat = images.reshape((-1,) + tuple(images.shape[-3:]))
    if out is None:
        out = torch.empty_like(images)
    elif out.shape != images.shape or out.dtype != images.dtype or not out.is_contiguous():
        raise ValueError("out must match images")
    src, dst = _describe(flat, "images"), _describe(out.reshape(flat.shape), "out")
    m = primaries_matrix(src_p, dst_p)
    desc = _lib.ColorConvert(_lib.TRC[in_trc], _lib.TRC[out_trc], (ctypes.c_float * 9)(*[float(v) for v in m.reshape(-1)]), 0)
    with torch.cuda.device(images.device):
        _lib.check(_lib.load().r360_convert_color(ctypes.byref(src), ctypes.byref(dst), ctypes.byref(desc),
                                                  1 if channel_order == "rgb" else 0,
                                                  _stream_handle(stream, images.device)))
    return out


# ---- decoder matrix --------------------------------------------------------------------------------------
# Video decoders hand over R'G'B' computed from Y'CbCr with ONE matrix; OpenCV's FFmpeg reader (swscale without
# colourspace details) and JPEG decoders (JFIF) use BT.601, whereas the cutter's filter chain declares its input
# BT.709 (`iall=bt709`, PC:299-309) and ffmpeg's colorspace filter then reads the same Y'CbCr samples with the
# BT.709 matrix.  R'G'B'(709) = M709 . M601^-1 . R'G'B'(601): a 3 x 3 matrix on the ENCODED values, independent of
# the range convention because offsets and scales are common to both conversions.

_LUMA = {"bt601": (0.299, 0.114), "bt709": (0.2126, 0.0722)}       # (Kr, Kb)


def ycbcr_to_rgb_matrix(standard: str):
    """Normalised Y'CbCr (Y in [0, 1], Cb / Cr in [-0.5, 0.5]) -> R'G'B' for BT.601 / BT.709 luma coefficients."""
    import numpy as np
    kr, kb = _LUMA[standard]
    kg = 1.0 - kr - kb
    return np.array([[1.0, 0.0, 2.0 * (1.0 - kr)],
                     [1.0, -2.0 * kb * (1.0 - kb) / kg, -2.0 * kr * (1.0 - kr) / kg],
                     [1.0, 2.0 * (1.0 - kb), 0.0]], dtype=np.float64)


def decoder_matrix_correction(decoded_with: str = "bt601", declared: str = "bt709"):
    """R'G'B' as decoded -> R'G'B' the declared matrix gives for the same Y'CbCr samples (float64 3 x 3, RGB order)."""
    import numpy as np
    if decoded_with == declared:
        return np.eye(3)
    return ycbcr_to_rgb_matrix(declared) @ np.linalg.inv(ycbcr_to_rgb_matrix(decoded_with))


def correct_decoder_matrix(images: torch.Tensor, *, decoded_with: str = "bt601", declared: str = "bt709",
                           channel_order: str = "bgr", out: Optional[torch.Tensor] = None,
                           stream: Optional[torch.cuda.Stream] = None) -> torch.Tensor:
    """Apply ``decoder_matrix_correction`` to [.., H, W, C] CUDA frames (``out`` may alias ``images``): the
    r360_convert_color kernel with both transfer curves set to linear, i.e. the matrix on the encoded values, clipped
    to the code range."""
    if channel_order not in ("bgr", "rgb"):
        raise ValueError("channel_order must be 'bgr' or 'rgb'")
    if not images.is_contiguous():
        raise ValueError("images must be contiguous")
    flat = images.reshape((-1,) + tuple(images.shape[-3:]))
    if out is None:
        out = torch.empty_like(images)
    elif out.shape != images.shape or out.dtype != images.dtype or not out.is_contiguous():
        raise ValueError("out must match images")
    src, dst = _describe(flat, "images"), _describe(out.reshape(flat.shape), "out")
    m = decoder_matrix_correction(decoded_with, declared)
    desc = _lib.ColorConvert(_lib.TRC["linear"], _lib.TRC["linear"], (ctypes.c_float * 9)(*[float(v) for v in m.reshape(-1)]), 0)
    with torch.cuda.device(images.device):
        _lib.check(_lib.load().r360_convert_color(ctypes.byref(src), ctypes.byref(dst), ctypes.byref(desc),
                                                  1 if channel_order == "rgb" else 0,
                                                  _stream_handle(stream, images.device)))
    return out
