"""Float64 coordinate oracle (TEST INFRASTRUCTURE -- see oracle/__init__.py).

Restates, in float64 NumPy, the geometry the reference uses to go from an
output (perspective) pixel to a source pixel position:

* panorama (ERP) source -- gs360_GUI.py:342-395 (normalise, pitch about X, yaw
  about Y, lon = atan2(x, z), lat = asin(y)) and gs360_GUI.py:419-424
  (lon/lat -> ERP pixel position, edge origin).  Output pixel centres follow
  cli_tools/gs360_DualFisheyeDistortionCalibration.py:1771-1778
  (``((i + 0.5) / w) * 2 - 1``), which is also what FFmpeg's v360 filter uses
  for its rectilinear output.
* dual-fisheye source -- gs360_DualFisheyeDistortionCalibration.py:1759-1823
  (equisolid projection, Brown distortion :975-1005, affinity b1/b2, validity
  by lens half-FOV and image bounds) and the lens choice of :1857-1907.

Two ERP pixel conventions are exposed because the replaced filter and the
in-repo code disagree by up to half a pixel:

``halfpixel``  x = (lon/2pi + 0.5) * W - 0.5      y = (0.5 - lat/pi) * H - 0.5
``v360``       x = (lon/2pi + 0.5) * (W - 1)      y = (0.5 - lat/pi) * (H - 1)
"""

from __future__ import annotations

import math
from typing import Dict, Mapping, Sequence, Tuple

import numpy as np

CONVENTIONS = ("halfpixel", "v360")


def view_rotation(yaw_deg: float, pitch_deg: float, roll_deg: float = 0.0) -> np.ndarray:
    """World-from-camera rotation: roll about the view axis, then pitch about X,
    then yaw about Y, with the signs of gs360_GUI.py:351-374.

    The reference never rolls (``roll=0`` at gs360_360PerspCut.py:312); roll is
    defined here as a right-handed turn about +z in the y-up camera frame.
    """
    cy, sy = math.cos(math.radians(yaw_deg)), math.sin(math.radians(yaw_deg))
    cp, sp = math.cos(math.radians(pitch_deg)), math.sin(math.radians(pitch_deg))
    cr, sr = math.cos(math.radians(roll_deg)), math.sin(math.radians(roll_deg))
    r_yaw = np.array([[cy, 0.0, sy], [0.0, 1.0, 0.0], [-sy, 0.0, cy]])
    r_pitch = np.array([[1.0, 0.0, 0.0], [0.0, cp, sp], [0.0, -sp, cp]])
    r_roll = np.array([[cr, -sr, 0.0], [sr, cr, 0.0], [0.0, 0.0, 1.0]])
    return r_yaw @ r_pitch @ r_roll


def _clamped_fov_rad(fov_deg: float) -> float:
    # gs360_GUI.py:437-438 and DF:1781-1782 clamp to [1e-3, 179.9] degrees.
    return math.radians(min(max(float(fov_deg), 1e-3), 179.9))


def fisheye_fov_from_dfov(d_fov_deg: float, width: int, height: int) -> Tuple[float, float]:
    """v360's (h_fov, v_fov) for ``output=fisheye:d_fov=...`` [upstream FFmpeg vf_v360.c
    ``fov_from_dfov``; not verifiable here -- no ffmpeg binary in the image]."""
    d = 0.5 * math.hypot(width, height)
    return d / width * d_fov_deg, d / height * d_fov_deg


def camera_rays(out_w: int, out_h: int, hfov_deg: float, vfov_deg: float,
                yaw_deg: float, pitch_deg: float, roll_deg: float = 0.0,
                projection: str = "rectilinear") -> np.ndarray:
    """Unit rays (h, w, 3), y up, after the view rotation (float64).

    ``projection="fisheye"`` is v360's ``output=fisheye`` (gs360_360PerspCut.py:375-379, preset
    fisheyeXY) [upstream ``fisheye_to_xyz``, unverified here]: flat coordinates
    (h_fov/180 * u, v_fov/180 * v), angle from the view axis = 90 degrees x their length
    (equidistant), azimuth = their direction."""
    u = ((np.arange(out_w, dtype=np.float64) + 0.5) / float(out_w)) * 2.0 - 1.0
    v = ((np.arange(out_h, dtype=np.float64) + 0.5) / float(out_h)) * 2.0 - 1.0
    uu, vv = np.meshgrid(u, v)
    rays = np.empty((out_h, out_w, 3), dtype=np.float64)
    if projection == "fisheye":
        fu = min(max(float(hfov_deg), 1e-3), 360.0) / 180.0 * uu
        fv = min(max(float(vfov_deg), 1e-3), 360.0) / 180.0 * vv
        alpha = 0.5 * math.pi * np.hypot(fu, fv)
        phi = np.arctan2(fv, fu)
        rays[..., 0] = np.sin(alpha) * np.cos(phi)
        rays[..., 1] = -np.sin(alpha) * np.sin(phi)
        rays[..., 2] = np.cos(alpha)
    elif projection == "rectilinear":
        rays[..., 0] = math.tan(_clamped_fov_rad(hfov_deg) * 0.5) * uu
        rays[..., 1] = math.tan(_clamped_fov_rad(vfov_deg) * 0.5) * (-vv)
        rays[..., 2] = 1.0
        rays /= np.linalg.norm(rays, axis=2, keepdims=True)
    else:
        raise ValueError("unknown projection: %r" % (projection,))
    if roll_deg:
        rays = rays @ view_rotation(0.0, 0.0, roll_deg).T
    # pitch then yaw, written out as in gs360_GUI.py:351-374 / DF:1310-1339
    cp, sp = math.cos(math.radians(pitch_deg)), math.sin(math.radians(pitch_deg))
    cy, sy = math.cos(math.radians(yaw_deg)), math.sin(math.radians(yaw_deg))
    x, y, z = rays[..., 0], rays[..., 1], rays[..., 2]
    y1 = cp * y + sp * z
    z1 = -sp * y + cp * z
    x2 = cy * x + sy * z1
    z2 = -sy * x + cy * z1
    return np.stack([x2, y1, z2], axis=-1)


def erp_map64(src_w: int, src_h: int, out_w: int, out_h: int,
              yaw_deg: float, pitch_deg: float, hfov_deg: float, vfov_deg: float,
              convention: str = "halfpixel", roll_deg: float = 0.0, projection: str = "rectilinear"
              ) -> Tuple[np.ndarray, np.ndarray]:
    """Source pixel position (x, y) in an ERP of size src_w x src_h for every
    output pixel.  x is NOT wrapped into [0, W): it lies in [-0.5, W - 0.5) for
    ``halfpixel`` and [0, W - 1] for ``v360``; the sampler wraps taps."""
    if convention not in CONVENTIONS:
        raise ValueError("unknown convention: %r" % (convention,))
    rays = camera_rays(out_w, out_h, hfov_deg, vfov_deg, yaw_deg, pitch_deg, roll_deg, projection)
    lon = np.arctan2(rays[..., 0], rays[..., 2])
    lat = np.arcsin(np.clip(rays[..., 1], -1.0, 1.0))
    if convention == "halfpixel":
        mx = (lon / (2.0 * math.pi) + 0.5) * src_w - 0.5
        my = (0.5 - lat / math.pi) * src_h - 0.5
    else:
        mx = (lon / (2.0 * math.pi) + 0.5) * (src_w - 1)
        my = (0.5 - lat / math.pi) * (src_h - 1)
    return mx, my


# --------------------------------------------------------------------------
# dual fisheye (equisolid + Brown), DF:1759-1823
# --------------------------------------------------------------------------

CALIB_KEYS = ("width", "height", "f", "cx", "cy", "k1", "k2", "k3", "k4",
              "p1", "p2", "b1", "b2")


def brown_distort(x: np.ndarray, y: np.ndarray, c: Mapping[str, float]):
    """DF:975-1005."""
    r2 = x * x + y * y
    r4 = r2 * r2
    radial = 1.0 + c["k1"] * r2 + c["k2"] * r4 + c["k3"] * (r4 * r2) + c["k4"] * (r4 * r4)
    xd, yd = x * radial, y * radial
    if c["p1"] != 0.0 or c["p2"] != 0.0:
        xy = x * y
        xd = xd + c["p1"] * (r2 + 2.0 * x * x) + 2.0 * c["p2"] * xy
        yd = yd + c["p2"] * (r2 + 2.0 * y * y) + 2.0 * c["p1"] * xy
    return xd, yd


def fisheye_map64(calib: Mapping[str, float], yaw_rel_deg: float, pitch_deg: float,
                  hfov_deg: float, vfov_deg: float, out_w: int, out_h: int,
                  lens_fov_deg: float) -> Tuple[np.ndarray, np.ndarray, np.ndarray]:
    """(map_x, map_y, valid) for one lens, float64 restatement of DF:1759-1823."""
    rays = camera_rays(out_w, out_h, hfov_deg, vfov_deg, yaw_rel_deg, pitch_deg)
    rx, ry, rz = rays[..., 0], rays[..., 1], rays[..., 2]
    theta = np.arccos(np.clip(rz, -1.0, 1.0))
    theta_max = math.radians(max(1.0, min(360.0, float(lens_fov_deg))) * 0.5)
    rho = np.sqrt(rx * rx + ry * ry)
    scale = np.zeros_like(rho)
    nz = rho > 1e-12
    if calib.get("model", "equisolid") == "equidistant":      # v360 input=fisheye: r = f * theta
        scale[nz] = theta[nz] / rho[nz]
    else:
        scale[nz] = 2.0 * np.sin(theta[nz] * 0.5) / rho[nz]
    xn = rx * scale
    yn = -ry * scale
    xd, yd = brown_distort(xn, yn, calib)
    cx0 = calib["width"] * 0.5 + calib["cx"]
    cy0 = calib["height"] * 0.5 + calib["cy"]
    mx = cx0 + xd * calib["f"] + xd * calib["b1"] + yd * calib["b2"]
    my = cy0 + yd * calib["f"]
    valid = theta <= theta_max
    valid &= (mx >= 0.0) & (mx <= calib["width"] - 1)
    valid &= (my >= 0.0) & (my <= calib["height"] - 1)
    return mx, my, valid


def ideal_fisheye_calib(width: int, height: int, projection: str, ih_fov_deg: float,
                        iv_fov_deg: float = None, convention: str = "halfpixel") -> Dict[str, float]:
    """An undistorted lens that fills a width x height image with ``ih_fov`` x ``iv_fov`` degrees: the source
    model behind ``v360=<fisheye|equisolid>:rectilinear:ih_fov=..:iv_fov=..`` (V2F:466-487) as a record of
    :func:`fisheye_map64`.  v360 normalises the image to [-1, 1] by ``r(fov / 2)`` with r(theta) = theta
    (``fisheye``, equidistant) or sin(theta / 2) (``equisolid``) *[upstream vf_v360.c, unverified here]*;
    ``halfpixel`` puts +-1 on the image edges (pixel index = (u + 1) W / 2 - 0.5), ``v360`` on the centres
    of the outermost pixels (index = (u + 1) (W - 1) / 2)."""
    iv_fov_deg = ih_fov_deg if iv_fov_deg is None else iv_fov_deg
    if projection not in ("equidistant", "equisolid"):
        raise ValueError("projection must be 'equidistant' or 'equisolid'")

    def radius(fov_deg):
        half = math.radians(max(1.0, min(360.0, float(fov_deg))) * 0.5)
        return half if projection == "equidistant" else 2.0 * math.sin(half * 0.5)

    if convention == "v360":
        half_w, half_h, shift = (width - 1) * 0.5, (height - 1) * 0.5, 0.0
    else:
        half_w, half_h, shift = width * 0.5, height * 0.5, -0.5
    fy = half_h / radius(iv_fov_deg)
    fx = half_w / radius(ih_fov_deg)
    return {"width": float(width), "height": float(height), "f": fy, "b1": fx - fy, "b2": 0.0,
            "cx": half_w + shift - width * 0.5, "cy": half_h + shift - height * 0.5,
            "k1": 0.0, "k2": 0.0, "k3": 0.0, "k4": 0.0, "p1": 0.0, "p2": 0.0, "model": projection}


def v360_fisheye_input_map64(width: int, height: int, projection: str, ih_fov_deg: float, iv_fov_deg: float,
                             hfov_deg: float, vfov_deg: float, out_w: int, out_h: int,
                             yaw_deg: float = 0.0, pitch_deg: float = 0.0, convention: str = "halfpixel"):
    """Independent statement of the same mapping, written the way v360 does it (normalised (uf, vf) in
    [-1, 1], no focal length): used by the tests to cross-check :func:`ideal_fisheye_calib`."""
    rays = camera_rays(out_w, out_h, hfov_deg, vfov_deg, yaw_deg, pitch_deg)
    rx, ry, rz = rays[..., 0], rays[..., 1], rays[..., 2]
    h = np.sqrt(rx * rx + ry * ry)
    theta = np.arctan2(h, rz)
    lh = np.where(h > 0.0, h, 1.0)
    if projection == "equidistant":
        rad = theta
        ru = math.radians(ih_fov_deg * 0.5)
        rv = math.radians(iv_fov_deg * 0.5)
    else:
        rad = np.sin(theta * 0.5)
        ru = math.sin(math.radians(ih_fov_deg * 0.25))
        rv = math.sin(math.radians(iv_fov_deg * 0.25))
    uf = rx / lh * rad / ru
    vf = -ry / lh * rad / rv
    if convention == "v360":
        return (uf + 1.0) * (width - 1) * 0.5, (vf + 1.0) * (height - 1) * 0.5
    return (uf + 1.0) * width * 0.5 - 0.5, (vf + 1.0) * height * 0.5 - 0.5


def undistort_map64(calib: Mapping[str, float], zoom: float, lens_fov_deg: float,
                    out_w: int = 0, out_h: int = 0, grid_x=None, grid_y=None):
    """Fisheye -> undistorted fisheye: (map_x, map_y, valid, valid_model), float64 restatement of
    ``_remap_for_zoom`` (DF:1008-1051) on the integer pixel grid of ``build_remap_cache``
    (DF:1136-1140) or on an explicit grid (``estimate_auto_undistort_zoom``, DF:1069-1071)."""
    w = int(out_w) or int(calib["width"])
    h = int(out_h) or int(calib["height"])
    gx = np.arange(w, dtype=np.float64) if grid_x is None else np.asarray(grid_x, dtype=np.float64)
    gy = np.arange(h, dtype=np.float64) if grid_y is None else np.asarray(grid_y, dtype=np.float64)
    dst_x, dst_y = np.meshgrid(gx, gy)
    cx0 = calib["width"] * 0.5 + calib["cx"]
    cy0 = calib["height"] * 0.5 + calib["cy"]
    den_y, den_x = calib["f"], calib["f"] + calib["b1"]
    if abs(den_y) < 1e-12 or abs(den_x) < 1e-12:
        raise ValueError("Invalid focal/b1 configuration caused division by zero.")
    y0 = (dst_y - cy0) / den_y
    x0 = (dst_x - cx0 - y0 * calib["b2"]) / den_x
    x, y = x0 / zoom, y0 / zoom
    xd, yd = brown_distort(x, y, calib)
    mx = cx0 + xd * calib["f"] + xd * calib["b1"] + yd * calib["b2"]
    my = cy0 + yd * calib["f"]
    r = np.sqrt(np.maximum(x * x + y * y, 0.0))
    theta = 2.0 * np.arcsin(np.clip(r * 0.5, 0.0, 1.0))
    theta_max = math.radians(max(1.0, min(360.0, float(lens_fov_deg))) * 0.5)
    valid_model = theta <= theta_max
    valid = valid_model & (mx >= 0.0) & (mx <= calib["width"] - 1) & (my >= 0.0) & (my <= calib["height"] - 1)
    return mx, my, valid, valid_model


def auto_undistort_zoom(calib: Mapping[str, float], lens_fov_deg: float = 190.0, sample_count: int = 192) -> float:
    """Smallest zoom (to bisection accuracy) at which no model-valid sample of a coarse grid reads
    outside the sensor: DF:1054-1117 (x1.2 bracketing up to 20 times, then 20 bisections)."""
    w, h = int(calib["width"]), int(calib["height"])
    steps = max(32, int(sample_count))
    gx, gy = np.linspace(0.0, w - 1.0, steps), np.linspace(0.0, h - 1.0, steps)

    def overflow(zoom):
        mx, my, _, vm = undistort_map64(calib, zoom, lens_fov_deg, grid_x=gx, grid_y=gy)
        if not vm.any():
            return 0.0
        sx, sy = mx[vm], my[vm]
        return float(max(np.max(np.maximum(0.0, -sx)), np.max(np.maximum(0.0, sx - (w - 1))),
                         np.max(np.maximum(0.0, -sy)), np.max(np.maximum(0.0, sy - (h - 1)))))

    if overflow(1.0) <= 0.0:
        return 1.0
    low = high = 1.0
    for _ in range(20):
        high *= 1.2
        if overflow(high) <= 0.0:
            break
    if overflow(high) > 0.0:
        return high
    for _ in range(20):
        mid = (low + high) * 0.5
        if overflow(mid) <= 0.0:
            high = mid
        else:
            low = mid
    return high


def wrap_angle_deg(a: float) -> float:
    """DF:1342-1345: wrap to [-180, 180)."""
    return ((float(a) + 180.0) % 360.0) - 180.0


def dualfisheye_view_maps(calib_x: Mapping[str, float], calib_y: Mapping[str, float],
                          specs: Sequence[Mapping[str, object]],
                          lens_x_yaw_deg: float = 0.0, lens_y_yaw_deg: float = 180.0,
                          lens_fov_deg: float = 190.0) -> Dict[str, Dict[str, object]]:
    """Per view: maps for both lenses, keep the one with the larger valid ratio,
    ties broken by smaller |relative yaw| (DF:1857-1907)."""
    out: Dict[str, Dict[str, object]] = {}
    for spec in specs:
        best = None
        for lens_key, lens_yaw, calib in (("X", lens_x_yaw_deg, calib_x),
                                          ("Y", lens_y_yaw_deg, calib_y)):
            yaw_rel = wrap_angle_deg(float(spec["yaw_deg"]) - lens_yaw)
            mx, my, valid = fisheye_map64(
                calib, yaw_rel, float(spec["pitch_deg"]), float(spec["hfov_deg"]),
                float(spec["vfov_deg"]), int(spec["width"]), int(spec["height"]),
                lens_fov_deg)
            key = (float(np.mean(valid)), -abs(yaw_rel))
            if best is None or key > best[0]:
                best = (key, lens_key, yaw_rel, mx, my, valid)
        out[str(spec["view_id"])] = {
            "lens_key": best[1], "yaw_rel_deg": best[2],
            "map_x": best[3], "map_y": best[4], "valid": best[5],
        }
    return out
