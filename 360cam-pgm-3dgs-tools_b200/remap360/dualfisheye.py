"""Host side of the dual-fisheye -> perspective stage (DualFisheyePipeline stage 3).

Mirrors the pieces of cli_tools/gs360_DualFisheyeDistortionCalibration.py that feed the remap:
Metashape calibration XML loading (DF:453-492, :754-828), the SFM10 ten-view layout
(DF:1243-1307), relative yaw and lens choice per view (DF:1342-1345, :1857-1907), and the
per-pair rendering loop (DF:1996-2055) -- which here is one batched CUDA call instead of ten
``cv2.remap`` calls on CPU maps -- and the zoom selection of the optional fisheye -> undistorted
fisheye output (DF:1054-1170).

The lens choice needs each candidate's valid ratio; it is obtained from the device projection
(``sample_coordinates``), i.e. with the same float64 math that the kernels sample with.
"""

from __future__ import annotations

import math
import pathlib
import xml.etree.ElementTree as ET
from dataclasses import dataclass
from typing import Dict, List, Optional, Sequence, Tuple

SUPPORTED_MODELS = {"equisolid_fisheye"}

SFM10_LAYOUT = (          # (view id, yaw as a function of the yaw delta, pitch sign)  DF:1281-1292
    ("A", lambda d: 0.0, 0), ("A_U", lambda d: 0.0, +1), ("A_D", lambda d: 0.0, -1),
    ("B", lambda d: +d, 0), ("E", lambda d: 180.0 - d, 0), ("F", lambda d: 180.0, 0),
    ("F_U", lambda d: 180.0, +1), ("F_D", lambda d: 180.0, -1), ("G", lambda d: 180.0 + d, 0),
    ("J", lambda d: 360.0 - d, 0),
)


@dataclass
class SensorCalibration:
    """DF:67-85."""
    sensor_id: str
    model_type: str
    width: int
    height: int
    f: float
    cx: float = 0.0
    cy: float = 0.0
    k1: float = 0.0
    k2: float = 0.0
    k3: float = 0.0
    k4: float = 0.0
    p1: float = 0.0
    p2: float = 0.0
    b1: float = 0.0
    b2: float = 0.0


def _float_child(node: Optional[ET.Element], name: str, default: float) -> float:
    if node is None:
        return default
    child = node.find(name)
    if child is None or child.text is None:
        return default
    try:
        return float(child.text.strip())
    except ValueError:
        return default


def load_metashape_calibration(xml_path) -> Tuple[Dict[str, SensorCalibration], Dict[str, str]]:
    """Sensor calibrations (adjusted before initial, DF:754-764) and camera label -> sensor id."""
    root = ET.parse(str(xml_path)).getroot()
    sensors: Dict[str, SensorCalibration] = {}
    for sensor in root.findall(".//sensors/sensor"):
        sid = sensor.attrib.get("id", "").strip()
        calibs = sensor.findall("calibration")
        if not sid or not calibs:
            continue
        node = next((c for cls in ("adjusted", "initial") for c in calibs
                     if c.attrib.get("class", "").strip().lower() == cls), calibs[0])
        model = (node.attrib.get("type") or sensor.attrib.get("type") or "").strip().lower()
        res = node.find("resolution")
        if res is None:
            res = sensor.find("resolution")
        if res is None:
            continue
        width, height = int(res.attrib.get("width", "0")), int(res.attrib.get("height", "0"))
        if width <= 0 or height <= 0:
            continue
        cal = SensorCalibration(sensor_id=sid, model_type=model, width=width, height=height,
                                **{k: _float_child(node, k, 0.0) for k in
                                   ("f", "cx", "cy", "k1", "k2", "k3", "k4", "p1", "p2", "b1", "b2")})
        if cal.f > 0.0:
            sensors[sid] = cal
    cameras = {}
    for cam in root.findall(".//cameras/camera"):
        label, sid = cam.attrib.get("label", "").strip(), cam.attrib.get("sensor_id", "").strip()
        if label and sid:
            cameras[label] = sid
    return sensors, cameras


def parse_sensor_dimensions(sensor_mm: str) -> Tuple[float, float]:
    """DF:1220-1240."""
    values = []
    for tok in str(sensor_mm or "").strip().replace("x", " ").replace("X", " ").replace(",", " ").split():
        try:
            values.append(float(tok))
        except ValueError:
            pass
    if not values:
        raise ValueError("Invalid --perspective-sensor-mm: '{}'".format(sensor_mm))
    w, h = float(values[0]), float(values[1] if len(values) > 1 else values[0])
    if w <= 0.0 or h <= 0.0:
        raise ValueError("Sensor dimensions must be positive: '{}'".format(sensor_mm))
    return w, h


def compute_view_fov_deg(focal_mm: float, sensor_mm: str) -> Tuple[float, float]:
    """DF:1243-1255 (clamped to [1, 179.9] degrees)."""
    f = float(focal_mm)
    if f <= 0.0:
        raise ValueError("--perspective-focal-mm must be > 0")
    sw, sh = parse_sensor_dimensions(sensor_mm)
    fov = [math.degrees(2.0 * math.atan(s / (2.0 * f))) for s in (sw, sh)]
    return tuple(max(1.0, min(179.9, a)) for a in fov)


def build_sfm10_specs(output_size: int, focal_mm: float, sensor_mm: str, yaw_delta_deg: float,
                      pitch_delta_deg: float) -> List[Dict[str, float]]:
    """DF:1258-1307."""
    size = int(output_size)
    if size <= 0:
        raise ValueError("--perspective-size must be > 0")
    yaw_delta, pitch_delta = float(yaw_delta_deg), float(pitch_delta_deg)
    if yaw_delta <= 0.0 or yaw_delta >= 180.0:
        raise ValueError("--perspective-yaw-delta-deg must be in (0, 180)")
    if pitch_delta <= 0.0 or pitch_delta >= 89.9:
        raise ValueError("--perspective-pitch-delta-deg must be in (0, 89.9)")
    hfov, vfov = compute_view_fov_deg(focal_mm, sensor_mm)
    return [{"view_id": vid, "yaw_deg": float(yaw(yaw_delta)), "pitch_deg": float(sign * pitch_delta) if sign else 0.0,
             "hfov_deg": float(hfov), "vfov_deg": float(vfov), "width": size, "height": size}
            for vid, yaw, sign in SFM10_LAYOUT]


def wrap_angle_deg(angle_deg: float) -> float:
    """DF:1342-1345: wrap to [-180, 180)."""
    return ((float(angle_deg) + 180.0) % 360.0) - 180.0


def to_device_calibration(cal: SensorCalibration, lens_fov_deg: float = 190.0):
    from .api import FisheyeCalibration
    if cal.model_type not in SUPPORTED_MODELS:
        raise ValueError("Unsupported sensor model '{}' (supported: {}).".format(
            cal.model_type, ", ".join(sorted(SUPPORTED_MODELS))))
    return FisheyeCalibration(width=cal.width, height=cal.height, f=cal.f, cx=cal.cx, cy=cal.cy, k1=cal.k1,
                              k2=cal.k2, k3=cal.k3, k4=cal.k4, p1=cal.p1, p2=cal.p2, b1=cal.b1, b2=cal.b2,
                              lens_fov_deg=float(lens_fov_deg))


def _undistort_overflow(cal: SensorCalibration, zoom: float, lens_fov_deg: float, steps: int) -> float:
    """How far (px) the model-valid samples of a steps x steps grid over the output read outside the
    sensor at this zoom (the ``overflow`` closure of DF:1073-1099).  Host-side planning, float64."""
    import numpy as np
    w, h = int(cal.width), int(cal.height)
    cx0, cy0 = w * 0.5 + cal.cx, h * 0.5 + cal.cy
    den_y, den_x = cal.f, cal.f + cal.b1
    if abs(den_y) < 1e-12 or abs(den_x) < 1e-12:
        raise ValueError("Invalid focal/b1 configuration caused division by zero.")
    jj, ii = np.meshgrid(np.linspace(0.0, h - 1.0, steps), np.linspace(0.0, w - 1.0, steps), indexing="ij")
    yn = (jj - cy0) / den_y
    xn = (ii - cx0 - yn * cal.b2) / den_x
    xn, yn = xn / zoom, yn / zoom
    r2 = xn * xn + yn * yn
    in_model = np.minimum(0.5 * np.sqrt(r2), 1.0) <= math.sin(math.radians(max(1.0, min(360.0, lens_fov_deg)) * 0.25))
    if not in_model.any():
        return 0.0
    radial = 1.0 + r2 * (cal.k1 + r2 * (cal.k2 + r2 * (cal.k3 + r2 * cal.k4)))
    xd = xn * radial + cal.p1 * (r2 + 2.0 * xn * xn) + 2.0 * cal.p2 * xn * yn
    yd = yn * radial + cal.p2 * (r2 + 2.0 * yn * yn) + 2.0 * cal.p1 * xn * yn
    sx = (cx0 + xd * (cal.f + cal.b1) + yd * cal.b2)[in_model]
    sy = (cy0 + yd * cal.f)[in_model]
    return float(max(0.0, (-sx).max(), (sx - (w - 1)).max(), (-sy).max(), (sy - (h - 1)).max()))


def estimate_auto_undistort_zoom(cal: SensorCalibration, sample_count: int = 192,
                                 lens_fov_deg: float = 190.0) -> float:
    """DF:1054-1117: 1.0 if nothing overflows; otherwise grow the zoom by 1.2 (at most 20 times)
    until nothing does and bisect 20 times between the last two values."""
    steps = max(32, int(sample_count))
    if _undistort_overflow(cal, 1.0, lens_fov_deg, steps) <= 0.0:
        return 1.0
    lo = hi = 1.0
    for _ in range(20):
        hi *= 1.2
        if _undistort_overflow(cal, hi, lens_fov_deg, steps) <= 0.0:
            break
    else:
        if _undistort_overflow(cal, hi, lens_fov_deg, steps) > 0.0:
            return hi
    for _ in range(20):
        mid = 0.5 * (lo + hi)
        if _undistort_overflow(cal, mid, lens_fov_deg, steps) <= 0.0:
            hi = mid
        else:
            lo = mid
    return hi


def build_undistort_items(calibs: Sequence[SensorCalibration], undistort_zoom: Optional[float] = None,
                          lens_fov_deg: float = 190.0):
    """One UndistortItem per lens image, zoom as ``build_remap_cache`` picks it (DF:1142-1152):
    the given value, else the auto estimate; never below 1e-6."""
    from .api import UndistortItem
    items = []
    for slot, cal in enumerate(calibs):
        if cal.model_type not in SUPPORTED_MODELS:
            raise ValueError("Unsupported sensor model '{}' (supported: {}).".format(
                cal.model_type, ", ".join(sorted(SUPPORTED_MODELS))))
        zoom = float(undistort_zoom) if undistort_zoom is not None else estimate_auto_undistort_zoom(
            cal, lens_fov_deg=float(lens_fov_deg))
        items.append(UndistortItem(zoom=max(1e-6, zoom), src_slot=slot, view_id=str(cal.sensor_id)))
    return items


def choose_lenses(calib_x: SensorCalibration, calib_y: SensorCalibration, specs: Sequence[Dict[str, object]],
                  lens_x_yaw_deg: float = 0.0, lens_y_yaw_deg: float = 180.0, lens_fov_deg: float = 190.0,
                  device="cuda"):
    """Per view: the lens with the larger valid ratio, ties to the smaller |relative yaw|
    (DF:1857-1907).  Returns (views, info): ``views`` are PerspectiveView objects with yaw relative
    to the chosen lens and ``src_slot`` 0 (X) / 1 (Y); ``info`` maps view id -> lens key and ratio."""
    from .api import PerspectiveView, sample_coordinates
    cals = [to_device_calibration(calib_x, lens_fov_deg), to_device_calibration(calib_y, lens_fov_deg)]
    views, info = [], {}
    for spec in specs:
        w, h = int(spec["width"]), int(spec["height"])
        cand = []
        for slot, (key, lens_yaw) in enumerate((("X", lens_x_yaw_deg), ("Y", lens_y_yaw_deg))):
            yaw_rel = wrap_angle_deg(float(spec["yaw_deg"]) - lens_yaw)
            cand.append(PerspectiveView(yaw_rel, float(spec["pitch_deg"]), float(spec["hfov_deg"]),
                                        float(spec["vfov_deg"]), src_slot=slot, view_id=str(spec["view_id"])))
        got = sample_coordinates(cand, (w, h), calibs=cals, path="direct", device=device)
        ratios = [float(got["valid"][n].float().mean()) for n in range(2)]
        best = max(range(2), key=lambda n: (ratios[n], -abs(cand[n].yaw_deg), -n))
        views.append(cand[best])
        info[str(spec["view_id"])] = {"lens_key": "XY"[best], "valid_ratio": ratios[best]}
    return views, cals, info
