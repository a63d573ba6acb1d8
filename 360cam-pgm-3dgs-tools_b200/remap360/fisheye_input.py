"""Fisheye frames -> one perspective view: the `--fisheye-perspective` step of the reference's
gs360_Video2Frames.py (V2F:330-371 flags, :383-403 validation, :467-493 filter), which runs

    v360=<fisheye|equisolid>:rectilinear:ih_fov=F:iv_fov=F:h_fov=..:v_fov=..:interp=cubic , ... , scale=S:S

inside ffmpeg.  Here the same geometry is a lens record for the fisheye projection of the remap kernels
(an undistorted lens of the requested field of view, equidistant or equisolid) and the view is rendered
directly at S x S -- one resampling instead of v360 at its default size followed by swscale.
"""

from __future__ import annotations

import math
from typing import Optional, Tuple

import torch

from . import api
from .perspcut import fov_from_focal_mm, v_fov_from_hfov

FISHEYE_SENSOR_WIDTH_MM = 36.0        # V2F:29
FISHEYE_INPUT_FOV_DEG = 190.0         # V2F:30
_V360_INPUT_NAME = {"equidistant": "fisheye", "equisolid": "equisolid"}       # V2F:470-473


def validate_args(focal_mm: float, size_px: int, input_fov_deg: float) -> Optional[str]:
    """The message V2F prints before `sys.exit(1)` (V2F:383-403), or None."""
    if focal_mm <= 0.0:
        return "Focal length must be greater than zero when using --fisheye-perspective."
    if size_px <= 0:
        return "Output size must be greater than zero when using --fisheye-perspective."
    if input_fov_deg <= 0.0:
        return "Input fisheye FOV must be greater than zero when using --fisheye-perspective."
    return None


def view_fov_deg(focal_mm: float, size_px: int) -> Tuple[float, float]:
    """(h_fov, v_fov) of the output view: V2F:468-482 (36 mm sensor, both clamped to [1, 179])."""
    focal_mm = max(focal_mm, 1e-6)
    size_px = max(size_px, 1)
    hfov = max(1.0, min(179.0, fov_from_focal_mm(focal_mm, FISHEYE_SENSOR_WIDTH_MM)))
    vfov = max(1.0, min(179.0, v_fov_from_hfov(hfov, size_px, size_px)))
    return hfov, vfov


def v360_filter(projection: str, input_fov_deg: float, focal_mm: float, size_px: int) -> str:
    """The filter string V2F builds (V2F:483-487), for logs and for callers that still drive ffmpeg."""
    name = _V360_INPUT_NAME.get(projection, "fisheye")
    fov = max(1.0, min(360.0, input_fov_deg))
    hfov, vfov = view_fov_deg(focal_mm, size_px)
    return ("v360=%s:rectilinear:ih_fov=%.6f:iv_fov=%.6f:h_fov=%.6f:v_fov=%.6f:interp=cubic"
            % (name, fov, fov, hfov, vfov))


def ideal_calibration(width: int, height: int, projection: str, ih_fov_deg: float,
                      iv_fov_deg: Optional[float] = None, convention: str = "halfpixel") -> api.FisheyeCalibration:
    """An undistorted lens filling a width x height image with ih_fov x iv_fov degrees.  v360 normalises the
    image to [-1, 1] by r(fov / 2), r(theta) = theta (`fisheye`) or sin(theta / 2) (`equisolid`); `halfpixel`
    puts +-1 on the image edges, `v360` on the centres of the outermost pixels."""
    if projection not in _V360_INPUT_NAME:
        raise ValueError("projection must be 'equidistant' or 'equisolid'")
    iv_fov_deg = ih_fov_deg if iv_fov_deg is None else iv_fov_deg

    def radius(fov_deg: float) -> float:
        half = math.radians(max(1.0, min(360.0, float(fov_deg))) * 0.5)
        return half if projection == "equidistant" else 2.0 * math.sin(half * 0.5)

    if convention == "v360":
        half_w, half_h, shift = (width - 1) * 0.5, (height - 1) * 0.5, 0.0
    elif convention == "halfpixel":
        half_w, half_h, shift = width * 0.5, height * 0.5, -0.5
    else:
        raise ValueError("unknown convention %r" % (convention,))
    fy = half_h / radius(iv_fov_deg)
    fx = half_w / radius(ih_fov_deg)
    # no circular field mask: v360's validity test is the image rectangle, which the sensor-bounds test gives
    return api.FisheyeCalibration(width=float(width), height=float(height), f=fy, b1=fx - fy,
                                  cx=half_w + shift - width * 0.5, cy=half_h + shift - height * 0.5,
                                  lens_fov_deg=360.0, model=projection)


def fisheye_to_perspective(frames: torch.Tensor, *, projection: str = "equidistant",
                           input_fov_deg: float = FISHEYE_INPUT_FOV_DEG, focal_mm: float = 8.0,
                           size_px: int = 1600, yaw_deg: float = 0.0, pitch_deg: float = 0.0,
                           interp: str = "cubic", convention: str = "halfpixel",
                           out: Optional[torch.Tensor] = None, path: str = "auto",
                           stream: Optional[torch.cuda.Stream] = None) -> torch.Tensor:
    """Fisheye frames [B, H, W, C] -> perspective views [B, size, size, C] with V2F's parameters
    (`--fisheye-projection`, `--fisheye-input-fov`, `--fisheye-focal-mm`, `--fisheye-size`).  Pixels
    whose ray leaves the fisheye image are black, as in v360."""
    msg = validate_args(focal_mm, size_px, input_fov_deg)
    if msg:
        raise ValueError(msg)
    if frames.dim() == 3:
        frames = frames.unsqueeze(0)
    b, h, w, _ = frames.shape
    hfov, vfov = view_fov_deg(focal_mm, size_px)
    cal = ideal_calibration(w, h, projection if projection in _V360_INPUT_NAME else "equidistant",
                            input_fov_deg, input_fov_deg, convention)
    view = api.PerspectiveView(yaw_deg, pitch_deg, hfov, vfov)
    res = api.remap_fisheye(frames.contiguous().unsqueeze(1), [cal], [view], (size_px, size_px), interp=interp,
                            border_value=0.0, fill_invalid=True,
                            out=None if out is None else out.unsqueeze(1), path=path, stream=stream)
    return res[:, 0]
