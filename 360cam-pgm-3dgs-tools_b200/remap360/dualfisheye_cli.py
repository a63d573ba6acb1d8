"""Drop-in command line of the dual-fisheye tool with the remap stages on the GPU.

Mirrors the process contract of cli_tools/gs360_DualFisheyeDistortionCalibration.py: the flags of
``parse_arguments`` (DF:124-449), the validation order and ``[ERR]`` texts of ``main`` (DF:2067-2372,
exit code 1), the ``[INFO]`` banner (DF:2386-2470), the ``[DRY]`` / ``[OK ]`` / ``[DONE]`` lines
(DF:2603-2846) and the output layout (``<root>/Images``, ``<root>/Masks``, DF:1536-1551; default roots
``<input>_perspective_colmap`` / ``_undistorted`` / ``_colorcorrected``, DF:2207-2239); exit code 2 when
any pair failed.

What runs where: file decode and PNG / TIFF encode stay on the host (``--workers`` threads); per pair ONE
upload of both lens images, then on the device the input colour pipeline (``r360_apply_lut``), the optional
fisheye -> undistorted fisheye remap, the ten perspective views and their masks (nearest, border 0), the
JPEG encode of the views (nvJPEG, remap360/codec.py), and one download per output group.  The reference does the same work with NumPy + ``cv2.remap`` on the CPU
(``process_pair_task``, DF:1910-2064).

The pose / COLMAP / Metashape-XML export behind ``--camera-extrinsics-xml``, ``--pointcloud-ply`` and
``--metadata-only`` (DF:1348-1686, :2812-2833) is host bookkeeping in ``remap360/pose_export.py``; it uses the
lens choice of the remap (made on the device) and writes the same files as the reference."""

from __future__ import annotations

import argparse
import os
import pathlib
import sys
from concurrent.futures import ThreadPoolExecutor
from dataclasses import dataclass, field
from typing import Dict, List, Optional, Sequence, Tuple

from . import dualfisheye as dfh

TEMPLATES = pathlib.Path(__file__).resolve().parent / "templates"
DEFAULT_CAMERA_XML = TEMPLATES / "Osmo360-Fisheye-Distortion.xml"
DEFAULT_DLOGM_LUT = TEMPLATES / "DJI Osmo 360 D-Log M to Rec.709 V1.cube"     # not shipped: pass --dlogm-lut
SUPPORTED_EXTS = ("jpg", "jpeg", "png", "tif", "tiff")
INTERPOLATIONS = ("nearest", "linear", "cubic", "lanczos4")


class UsageError(Exception):
    """A condition the reference answers with ``[ERR] ...`` on stderr and exit code 1."""


def create_arg_parser() -> argparse.ArgumentParser:
    ap = argparse.ArgumentParser(
        description="Dual-fisheye frames -> colour-corrected / undistorted fisheye / perspective views (CUDA remap).")
    add = ap.add_argument
    add("-i", "--input-dir", default=None, help="directory of *_X / *_Y fisheye frames")
    add("--metadata-only", action="store_true", help="export camera metadata only (not available in this backend)")
    add("-x", "--camera-xml", default=str(DEFAULT_CAMERA_XML), help="Metashape calibration XML")
    add("-o", "--output-dir", default=None, help="undistorted fisheye output directory")
    add("--suffixes", default="_X,_Y", help="lens suffixes, X first")
    add("--ext", default="jpg,jpeg,png,tif,tiff", help="input extensions")
    add("--input-lut", default=None, help=".cube 3-D LUT applied to every input frame")
    add("--lut-output-color-space", default="srgb", help="passthrough | srgb")
    add("--input-color-profile", choices=("native", "osmo360-dlogm"), default="native")
    add("--dlogm-lut", default=str(DEFAULT_DLOGM_LUT))
    add("--sensor-id-x", default=None)
    add("--sensor-id-y", default=None)
    add("--interpolation", choices=INTERPOLATIONS, default="cubic")
    add("--undistort-zoom", default="auto")
    add("--mask-outside-model", dest="mask_outside_model", action="store_true")
    add("--no-mask-outside-model", dest="mask_outside_model", action="store_false")
    ap.set_defaults(mask_outside_model=True)
    add("--mask-value", type=int, default=0)
    add("--limit", type=int, default=0, help="deprecated, ignored")
    add("--workers", type=int, default=max(1, os.cpu_count() or 1), help="host decode / encode threads")
    add("--memory-throttle-percent", type=float, default=80.0)
    add("--dry-run", action="store_true")
    add("--report-json", default=None, help="deprecated, ignored")
    add("--no-perspective", action="store_true")
    add("--save-fisheye-output", action="store_true")
    add("--save-color-corrected-output", action="store_true")
    add("--color-corrected-output-dir", default=None)
    add("--fisheye-output-dir", default=None)
    add("--no-fisheye-output", action="store_true")
    add("--perspective-output-dir", default=None)
    add("--perspective-ext", default="jpg")
    add("--perspective-mask-ext", default="png")
    add("--perspective-size", type=int, default=1750)
    add("--perspective-focal-mm", type=float, default=14.0)
    add("--perspective-sensor-mm", default="36 36")
    add("--perspective-yaw-delta-deg", type=float, default=40.0)
    add("--perspective-pitch-delta-deg", type=float, default=40.0)
    add("--perspective-jpeg-quality", type=int, default=95)
    add("--lens-fov-deg", type=float, default=190.0)
    add("--lens-x-yaw-deg", type=float, default=0.0)
    add("--lens-y-yaw-deg", type=float, default=180.0)
    add("--camera-extrinsics-xml", default=None)
    add("--pointcloud-ply", default=None)
    add("--mask-input-dir", default=None)
    add("--perspective-metashape-xml-name", default="perspective_cams.xml")
    return ap


def parse_undistort_zoom_arg(value) -> Optional[float]:
    """DF:465-478: ``auto`` / empty -> None, else a positive float."""
    text = (value or "").strip().lower()
    if text in ("", "auto"):
        return None
    zoom = float(text)
    if zoom <= 0.0:
        raise ValueError("undistort zoom must be > 0")
    return zoom


def _abs(path) -> pathlib.Path:
    return pathlib.Path(path).expanduser().resolve()


@dataclass
class PairJob:
    index: int
    base: str
    x_path: pathlib.Path
    y_path: pathlib.Path
    sensor_x: str
    sensor_y: str
    x_mask: Optional[pathlib.Path] = None
    y_mask: Optional[pathlib.Path] = None


@dataclass
class RunPlan:
    """Everything ``main`` decides before touching an image."""
    args: argparse.Namespace
    metadata_only: bool
    frames_dir: Optional[pathlib.Path]
    fisheye_dir: pathlib.Path
    persp_root: pathlib.Path
    color_dir: pathlib.Path
    want_fisheye: bool
    want_color: bool
    want_persp: bool
    calibration_xml: pathlib.Path
    extrinsics_xml: Optional[pathlib.Path]
    pointcloud: Optional[pathlib.Path]
    mask_dir: Optional[pathlib.Path]
    lut_path: Optional[pathlib.Path]
    lut: object
    lut_space: str
    zoom: Optional[float]
    suffixes: Tuple[str, str]
    sensors: Dict[str, dfh.SensorCalibration] = field(default_factory=dict)
    camera_to_sensor: Dict[str, str] = field(default_factory=dict)
    pairs: List[Tuple[str, pathlib.Path, pathlib.Path]] = field(default_factory=list)
    mask_value: int = 0
    workers: int = 1
    throttle: float = 0.8

    @property
    def images_dir(self) -> pathlib.Path:
        return self.persp_root / "Images"

    @property
    def masks_dir(self) -> pathlib.Path:
        return self.persp_root / "Masks"


def find_pairs(frames_dir: pathlib.Path, exts: Sequence[str], suffixes: Sequence[str]):
    """DF:831-914: files with a wanted extension whose stem ends in one of the suffixes, grouped by the
    stem without the suffix; a group counts when it has both an X and a Y file.  Returns (files, pairs)."""
    x_suffix, y_suffix = suffixes[0], suffixes[1]
    files = [p for p in sorted(frames_dir.iterdir())
             if p.is_file() and p.suffix.lower().lstrip(".") in exts and any(p.stem.endswith(s) for s in suffixes)]
    groups: Dict[str, Dict[str, pathlib.Path]] = {}
    for p in files:
        if p.stem.endswith(x_suffix):
            groups.setdefault(p.stem[:-len(x_suffix)], {})["X"] = p
        elif p.stem.endswith(y_suffix):
            groups.setdefault(p.stem[:-len(y_suffix)], {})["Y"] = p
    pairs = [(b, g["X"], g["Y"]) for b, g in sorted(groups.items()) if "X" in g and "Y" in g]
    return files, pairs


def sensor_for_file(path: pathlib.Path, plan: RunPlan) -> Optional[str]:
    """DF:851-876: camera label, then the --sensor-id-x/-y suffix fallback, then a lone sensor."""
    sid = plan.camera_to_sensor.get(path.stem)
    if sid in plan.sensors:
        return sid
    a = plan.args
    if a.sensor_id_x and path.stem.endswith(plan.suffixes[0]) and a.sensor_id_x in plan.sensors:
        return a.sensor_id_x
    if a.sensor_id_y and path.stem.endswith(plan.suffixes[1]) and a.sensor_id_y in plan.sensors:
        return a.sensor_id_y
    return next(iter(plan.sensors)) if len(plan.sensors) == 1 else None


def resolve_plan(args: argparse.Namespace) -> RunPlan:
    """The checks of DF:2071-2372 in the reference's order; raises UsageError with its message."""
    from . import color
    try:
        zoom = parse_undistort_zoom_arg(args.undistort_zoom)
    except Exception as exc:
        raise UsageError("--undistort-zoom: {}".format(exc))
    metadata_only = bool(args.metadata_only)
    frames_dir = _abs(args.input_dir) if str(args.input_dir or "").strip() else None
    if frames_dir is None and not metadata_only:
        raise UsageError("--input-dir is required unless --metadata-only is used.")
    profile = str(args.input_color_profile).strip().lower()
    lut_path = None
    if args.input_lut:
        lut_path = _abs(args.input_lut)
    elif profile == "osmo360-dlogm":
        lut_path = _abs(args.dlogm_lut)
    elif profile != "native":
        raise UsageError("Unsupported --input-color-profile: {}".format(profile))
    lut = None
    if lut_path is not None:
        try:
            lut = color.load_cube_lut(lut_path)
        except Exception as exc:
            raise UsageError("Failed to load input LUT: {}".format(exc))
    try:
        lut_space = color.normalize_lut_output_color_space(str(args.lut_output_color_space).strip().lower())
    except Exception as exc:
        raise UsageError(str(exc))
    suffixes = [t.strip() for t in args.suffixes.split(",") if t.strip()]
    if len(suffixes) < 2:
        raise UsageError("--suffixes must include at least two values like '_X,_Y'.")
    if frames_dir is not None:
        if frames_dir.is_file():
            raise UsageError("Input must be a directory of fisheye frames, not a video file.\n"
                             "Use gs360_Video2Frames.py to extract *_X/*_Y images first.")
        if not frames_dir.is_dir():
            raise UsageError("Input path not found: {}".format(frames_dir))
    want_fisheye = bool(args.save_fisheye_output) and not metadata_only
    want_color = bool(args.save_color_corrected_output) and not metadata_only
    want_persp = not bool(args.no_perspective) and not metadata_only
    if not metadata_only and not (want_fisheye or want_persp or want_color):
        raise UsageError("All outputs are disabled. Enable perspective, "
                         "--save-fisheye-output, or --save-color-corrected-output.")
    extrinsics = None
    if str(args.camera_extrinsics_xml or "").strip():
        extrinsics = _abs(args.camera_extrinsics_xml)
        if not extrinsics.is_file():
            raise UsageError("Camera extrinsics XML not found: {}".format(extrinsics))
        if not want_persp and not metadata_only:
            raise UsageError("--camera-extrinsics-xml requires perspective output.")

    def root(explicit, suffix, unused_name, from_extrinsics=False):
        if explicit:
            return _abs(explicit)
        if frames_dir is not None:
            return frames_dir.with_name(frames_dir.name + suffix)
        if from_extrinsics and extrinsics is not None:
            return extrinsics.with_name(extrinsics.stem + suffix)
        return pathlib.Path.cwd() / unused_name

    fisheye_dir = root(args.output_dir or args.fisheye_output_dir, "_undistorted", "_unused_dualfisheye_undistorted")
    persp_root = root(args.perspective_output_dir, "_perspective_colmap", "perspective_colmap", True)
    color_dir = root(args.color_corrected_output_dir, "_colorcorrected", "_unused_colorcorrected")
    pointcloud = None
    if str(args.pointcloud_ply or "").strip():
        pointcloud = _abs(args.pointcloud_ply)
        if not pointcloud.is_file():
            raise UsageError("Point cloud PLY not found: {}".format(pointcloud))
    if metadata_only and extrinsics is None:
        raise UsageError("--metadata-only requires --camera-extrinsics-xml.")
    if metadata_only and pointcloud is None:
        raise UsageError("--metadata-only requires --pointcloud-ply.")
    camera_xml = _abs(args.camera_xml) if str(args.camera_xml or "").strip() else None
    calibration_xml = extrinsics or camera_xml
    if calibration_xml is None:
        raise UsageError("Specify --camera-extrinsics-xml or --camera-xml.")
    if not calibration_xml.is_file():
        raise UsageError("Calibration XML not found: {}".format(calibration_xml))
    mask_dir = None
    if str(args.mask_input_dir or "").strip():
        mask_dir = _abs(args.mask_input_dir)
        if not mask_dir.is_dir():
            raise UsageError("Mask input directory not found: {}".format(mask_dir))
        if not want_persp and not metadata_only:
            raise UsageError("--mask-input-dir requires perspective output.")
    exts = [t.strip().lower().lstrip(".") for t in args.ext.split(",") if t.strip()] or list(SUPPORTED_EXTS)
    sensors, camera_to_sensor = dfh.load_metashape_calibration(calibration_xml)
    if not sensors:
        raise UsageError("No usable calibration found in XML.")
    bad = sorted(s.sensor_id for s in sensors.values() if s.model_type not in dfh.SUPPORTED_MODELS)
    if bad:
        raise UsageError("Unsupported model types in sensors: {}".format(", ".join(bad)))
    plan = RunPlan(args=args, metadata_only=metadata_only, frames_dir=frames_dir, fisheye_dir=fisheye_dir,
                   persp_root=persp_root, color_dir=color_dir, want_fisheye=want_fisheye, want_color=want_color,
                   want_persp=want_persp, calibration_xml=calibration_xml, extrinsics_xml=extrinsics,
                   pointcloud=pointcloud, mask_dir=mask_dir, lut_path=lut_path, lut=lut, lut_space=lut_space,
                   zoom=zoom, suffixes=(suffixes[0], suffixes[1]), sensors=sensors,
                   camera_to_sensor=camera_to_sensor)
    if frames_dir is not None:
        files, plan.pairs = find_pairs(frames_dir, exts, suffixes)
        if not files:
            raise UsageError("No target images found in {}".format(frames_dir))
        if not plan.pairs:
            raise UsageError("No valid X/Y fisheye pairs found in {}".format(frames_dir))
    if args.limit:
        print("[WARN] --limit is deprecated and ignored. Processing all pairs.")
    if args.report_json:
        print("[WARN] --report-json is deprecated and ignored.")
    plan.mask_value = int(max(0, min(255, args.mask_value)))
    plan.workers = int(args.workers)
    if plan.workers < 1:
        raise UsageError("--workers must be >= 1.")
    plan.throttle = float(args.memory_throttle_percent) / 100.0
    if plan.throttle <= 0.0 or plan.throttle > 1.0:
        raise UsageError("--memory-throttle-percent must be > 0 and <= 100.")
    return plan


def announce(plan: RunPlan) -> None:
    """The [INFO] banner of DF:2386-2470."""
    a = plan.args
    say = lambda text: print("[INFO] " + text)   # noqa: E731
    say("input:  {}".format(plan.frames_dir) if plan.frames_dir is not None else "input:  disabled (--metadata-only)")
    say("fisheye output: {}".format(plan.fisheye_dir if plan.want_fisheye else "disabled"))
    if plan.want_persp or plan.metadata_only:
        say("perspective output: {}".format(plan.persp_root))
        say("perspective xml: {}".format(plan.persp_root / a.perspective_metashape_xml_name))
        say("perspective images dir: {}".format(plan.images_dir))
        say("perspective sparse dir: {}".format(plan.persp_root / "Sparse" / "0"))
        say("perspective masks dir: {}".format(plan.masks_dir))
    else:
        say("perspective output: disabled")
    say("color-corrected output: {}".format(plan.color_dir if plan.want_color else "disabled"))
    say("calibration xml: {}".format(plan.calibration_xml))
    say("pairs:  {}".format(len(plan.pairs)))
    say("files:  {}".format(2 * len(plan.pairs)))
    say("camera extrinsics xml: {}".format(plan.extrinsics_xml if plan.extrinsics_xml is not None else "disabled"))
    say("pointcloud ply: {}".format(plan.pointcloud if plan.pointcloud is not None else "disabled"))
    if plan.mask_dir is not None:
        say("mask input dir: {}".format("ignored (--metadata-only)" if plan.metadata_only else plan.mask_dir))
    else:
        say("mask input dir: disabled")
    say("workers: {} (memory auto-throttle > {:.1f}%)".format(plan.workers, plan.throttle * 100.0))
    say("pair worker mode: {}".format("disabled (--metadata-only)" if plan.metadata_only else "enabled"))
    if plan.lut_path is not None:
        say("input LUT: {}".format(plan.lut_path))
        say("LUT output color space: {}".format(plan.lut_space))
    else:
        say("input LUT: disabled")
    if not plan.want_fisheye:
        say("undistort zoom: unused (direct perspective path)")
    elif plan.zoom is None:
        say("undistort zoom: auto")
    else:
        say("undistort zoom: {:.6f}".format(plan.zoom))


def match_masks(mask_dir: pathlib.Path, jobs: Sequence[PairJob]) -> None:
    """DF:1563-1596: masks are matched to the pair's files by exact file name."""
    names = {p.name: p for p in sorted(mask_dir.iterdir()) if p.is_file()}
    missing = set()
    for job in jobs:
        job.x_mask, job.y_mask = names.get(job.x_path.name), names.get(job.y_path.name)
        missing.update(p.name for p, m in ((job.x_path, job.x_mask), (job.y_path, job.y_mask)) if m is None)
    if missing:
        listed = sorted(missing)
        raise ValueError("Missing mask images in {}: {}".format(
            mask_dir, ", ".join(listed[:8]) + (", ..." if len(listed) > 8 else "")))


class PairRenderer:
    """Device side of one run: cached view sets per sensor pair, one call per X/Y pair."""

    def __init__(self, plan: RunPlan, specs, zooms: Dict[str, float], jpeg_views: bool = False):
        self.plan, self.specs, self.zooms, self.jpeg_views = plan, specs, zooms, jpeg_views
        self._views: Dict[Tuple[str, str], tuple] = {}
        self._codec = None

    def _jpeg_codec(self):
        if self._codec is None:
            from .executor import _gpu_codec
            self._codec = _gpu_codec() or False                  # False: tried, not available
        return self._codec or None

    def views_for(self, sensor_x: str, sensor_y: str):
        key = (sensor_x, sensor_y)
        if key not in self._views:
            a = self.plan.args
            self._views[key] = dfh.choose_lenses(self.plan.sensors[sensor_x], self.plan.sensors[sensor_y], self.specs,
                                                 float(a.lens_x_yaw_deg), float(a.lens_y_yaw_deg), float(a.lens_fov_deg))
        return self._views[key]

    def render(self, job: PairJob, image_x, image_y, mask_x, mask_y) -> Dict[str, list]:
        """NumPy images in (as cv2 reads them), NumPy outputs back: {"color": [x, y], "fisheye": [x, y],
        "views": [...], "masks": [...]} for the stages that are enabled."""
        import numpy as np
        import torch
        from . import api, color
        plan, a = self.plan, self.plan.args
        pair = None
        if isinstance(image_x, bytes) and isinstance(image_y, bytes):
            # JPEG frames arrive undecoded: nvJPEG writes them straight into the pair tensor
            jc = self._jpeg_codec()
            try:
                if jc is None:
                    raise RuntimeError("no GPU codec")
                wx, hx, cx = jc.info(image_x)
                if jc.info(image_y) != (wx, hx, cx) or cx != 3:
                    raise RuntimeError("X/Y differ")
                pair = torch.empty((1, 2, hx, wx, 3), dtype=torch.uint8, device="cuda")
                jc.decode(image_x, out=pair[0, 0])
                jc.decode(image_y, out=pair[0, 1])
                image_x = image_y = np.empty((hx, wx, 3), dtype=np.uint8)       # shape / dtype carriers
            except Exception:
                import cv2
                pair = None
                image_x, image_y = (cv2.imdecode(np.frombuffer(b, np.uint8), cv2.IMREAD_UNCHANGED) for b in (image_x, image_y))
                if image_x is None or image_y is None:
                    raise RuntimeError("Failed to read image: {}".format(job.x_path if image_x is None else job.y_path))
        if image_x.shape != image_y.shape or image_x.dtype != image_y.dtype:
            raise RuntimeError("X/Y images differ in size or type: {} vs {}".format(image_x.shape, image_y.shape))
        if image_x.dtype not in (np.uint8, np.uint16):
            raise RuntimeError("Unsupported image dtype: {}".format(image_x.dtype))

        def up(arrs):
            host = np.stack([x if x.ndim == 3 else x[..., None] for x in arrs])
            if host.dtype == np.uint16:
                return torch.from_numpy(host.view(np.int16)).cuda().view(torch.uint16)
            return torch.from_numpy(host).cuda()

        def down(t):
            t = t.contiguous()
            host = t.view(torch.int16).cpu().numpy().view(np.uint16) if t.dtype == torch.uint16 else t.cpu().numpy()
            return [img[..., 0] if img.shape[-1] == 1 else img for img in host]

        if pair is None:
            pair = up([image_x, image_y])[None]                          # [1, 2, H, W, C]
        if plan.lut is not None:
            color.apply_input_color_pipeline(pair, plan.lut, plan.lut_space, out=pair)
        out: Dict[str, list] = {}
        if plan.want_color:
            out["color"] = down(pair[0])
        interp, fill, bv = a.interpolation, bool(a.mask_outside_model), plan.mask_value
        if plan.want_fisheye:
            cals = [dfh.to_device_calibration(plan.sensors[s], float(a.lens_fov_deg)) for s in (job.sensor_x, job.sensor_y)]
            for cal, img, name in zip(cals, (image_x, image_y), (job.x_path.name, job.y_path.name)):
                if (img.shape[1], img.shape[0]) != (int(cal.width), int(cal.height)):         # DF:1187-1196
                    raise RuntimeError("Resolution mismatch for {}: got {}x{}, expected {}x{}".format(
                        name, img.shape[1], img.shape[0], int(cal.width), int(cal.height)))
            items = [api.UndistortItem(self.zooms[job.sensor_x], 0), api.UndistortItem(self.zooms[job.sensor_y], 1)]
            out["fisheye"] = down(api.undistort_fisheye(pair, cals, items, interp=interp, border_value=bv,
                                                         fill_invalid=fill)[0])
        if plan.want_persp:
            views, cals, _info = self.views_for(job.sensor_x, job.sensor_y)
            size = (int(a.perspective_size), int(a.perspective_size))
            rendered = api.remap_fisheye(pair, cals, views, size, interp=interp, border_value=bv, fill_invalid=fill)[0]
            jc = None
            if self.jpeg_views and rendered.dtype == torch.uint8 and rendered.shape[-1] in (1, 3):
                jc = self._jpeg_codec()
            if jc is not None:          # JPEG views leave the device already encoded (4:4:4, --perspective-jpeg-quality)
                out["views"] = [jc.encode(img, int(a.perspective_jpeg_quality)) for img in rendered]
            else:
                out["views"] = down(rendered)
            if plan.mask_dir is not None:
                if mask_x is None or mask_y is None:
                    raise RuntimeError("Mask source missing for pair '{}'.".format(job.base))
                if mask_x.shape != mask_y.shape or mask_x.dtype != mask_y.dtype:
                    raise RuntimeError("X/Y masks differ in size or type")
                masks = up([mask_x, mask_y])[None]
                out["masks"] = down(api.remap_fisheye(masks, cals, views, size, interp="nearest", border_value=0,
                                                      fill_invalid=fill)[0])
        return out


def _write(path: pathlib.Path, image, jpeg_quality: Optional[int] = None) -> None:
    import cv2
    path.parent.mkdir(parents=True, exist_ok=True)
    if isinstance(image, (bytes, bytearray)):          # encoded on the GPU
        path.write_bytes(image)
        return
    params: List[int] = []
    if jpeg_quality is not None and path.suffix.lower() in (".jpg", ".jpeg"):
        params = [int(cv2.IMWRITE_JPEG_QUALITY), int(max(1, min(100, jpeg_quality)))]
    if not cv2.imwrite(str(path), image, params):
        raise RuntimeError("Failed to write image: {}".format(path))


def _read(path: pathlib.Path, what: str = "image"):
    import cv2
    image = cv2.imread(str(path), cv2.IMREAD_UNCHANGED)
    if image is None:
        raise RuntimeError("Failed to read {}: {}".format(what, path))
    return image


def run(plan: RunPlan) -> int:
    a = plan.args
    errors: List[str] = []
    counts = {"processed": 0, "skipped": 0, "color": 0, "persp": 0, "mask": 0}
    jobs: List[PairJob] = []
    label_pairs = []
    if plan.metadata_only:
        # no images are read: the X/Y pairs are the camera labels of the extrinsics XML (DF:2477-2499)
        from . import pose_export
        labels = set(pose_export.camera_transform_map(plan.extrinsics_xml)) if plan.extrinsics_xml is not None else None
        label_pairs = pose_export.metadata_only_pairs(plan.camera_to_sensor, plan.sensors, plan.suffixes[0],
                                                      plan.suffixes[1], labels)
        if not label_pairs:
            raise UsageError("No valid X/Y camera label pairs found in extrinsics XML.")
    for index, (base, x_path, y_path) in enumerate(plan.pairs, start=1):
        sx, sy = sensor_for_file(x_path, plan), sensor_for_file(y_path, plan)
        if sx is None or sy is None:
            counts["skipped"] += 2
            print("[SKIP] {}: sensor_id unresolved".format(base))
            continue
        jobs.append(PairJob(index, base, x_path, y_path, sx, sy))
    if plan.mask_dir is not None and not plan.metadata_only:
        try:
            match_masks(plan.mask_dir, jobs)
        except Exception as exc:
            raise UsageError(str(exc))
    zooms: Dict[str, float] = {}
    if plan.want_fisheye:
        for sid in sorted({s for j in jobs for s in (j.sensor_x, j.sensor_y)}):
            try:
                zooms[sid] = dfh.build_undistort_items([plan.sensors[sid]], plan.zoom, float(a.lens_fov_deg))[0].zoom
                print("[INFO] sensor {} undistort_zoom={:.6f}".format(sid, zooms[sid]))
            except Exception as exc:
                errors.append("[ERR] sensor {}: remap build failed ({})".format(sid, exc))
                print(errors[-1])
        if errors:
            return 2
    specs = []
    if plan.want_persp or plan.metadata_only:
        specs = dfh.build_sfm10_specs(int(a.perspective_size), float(a.perspective_focal_mm), str(a.perspective_sensor_mm),
                                      float(a.perspective_yaw_delta_deg), float(a.perspective_pitch_delta_deg))
    persp_ext = "." + a.perspective_ext.strip().lstrip(".").lower()
    mask_ext = "." + a.perspective_mask_ext.strip().lstrip(".").lower()
    n_pairs = len(plan.pairs)
    ok_bases = set()

    if a.dry_run:
        total = max(1, len(jobs))
        for job in jobs:
            tag = "{:4d}/{:4d}".format(job.index, total)
            if plan.want_color:
                for p in (job.x_path, job.y_path):
                    print("[DRY][COLOR] {} {} -> {}".format(tag, p.name, p.name))
                counts["color"] += 2
            if plan.want_fisheye:
                for p, sid in ((job.x_path, job.sensor_x), (job.y_path, job.sensor_y)):
                    print("[DRY] {} {} -> {} (sensor_id={})".format(tag, p.name, p.name, sid))
            if plan.want_persp:
                for spec in specs:
                    print("[DRY][PERSP] {} {}_{}{}".format(tag, job.base, spec["view_id"], persp_ext))
                    if plan.mask_dir is not None:
                        print("[DRY][MASK ] {} {}_{}{}".format(tag, job.base, spec["view_id"], mask_ext))
                counts["persp"] += len(specs)
                if plan.mask_dir is not None:
                    counts["mask"] += len(specs)
            if not plan.metadata_only:
                counts["processed"] += 2
            ok_bases.add(job.base)
    elif not plan.metadata_only and jobs:
        renderer = PairRenderer(plan, specs, zooms, jpeg_views=persp_ext in (".jpg", ".jpeg"))
        quality = int(a.perspective_jpeg_quality)
        gpu_decode = renderer._jpeg_codec() is not None

        def load(job: PairJob):
            masks = (None, None)
            if plan.mask_dir is not None:
                masks = (_read(job.x_mask, "mask image"), _read(job.y_mask, "mask image"))
            if gpu_decode and all(p.suffix.lower() in (".jpg", ".jpeg") for p in (job.x_path, job.y_path)):
                return job.x_path.read_bytes(), job.y_path.read_bytes(), masks[0], masks[1]
            return _read(job.x_path), _read(job.y_path), masks[0], masks[1]

        def finish(job: PairJob, out: Dict[str, list]):
            """Encode and write one pair's outputs; returns the names per group."""
            done = {"color": [], "fisheye": [], "views": [], "masks": []}
            for key, folder in (("color", plan.color_dir), ("fisheye", plan.fisheye_dir)):
                for img, p in zip(out.get(key, []), (job.x_path, job.y_path)):
                    _write(folder / p.name, img)
                    done[key].append(p.name)
            for key, folder, ext in (("views", plan.images_dir, persp_ext), ("masks", plan.masks_dir, mask_ext)):
                for img, spec in zip(out.get(key, []), specs):
                    name = "{}_{}{}".format(job.base, spec["view_id"], ext)
                    _write(folder / name, img, quality)
                    done[key].append(name)
            return done

        def report(job: PairJob, done) -> None:
            tag = "{:4d}/{:4d}".format(job.index, n_pairs)
            for name in done["color"]:
                print("[OK ][COLOR] {} {} -> {}".format(tag, name, name))
            for name in done["fisheye"]:
                print("[OK ][FISH] {} {} -> {}".format(tag, name, name))
            if done["views"]:
                print("[OK ][PERSP] {} {} -> {} views".format(tag, job.base, len(done["views"])))
            if done["masks"]:
                print("[OK ][MASK ] {} {} -> {} masks".format(tag, job.base, len(done["masks"])))
            counts["processed"] += 2
            counts["color"] += len(done["color"])
            counts["persp"] += len(done["views"])
            counts["mask"] += len(done["masks"])
            ok_bases.add(job.base)

        def fail(job: PairJob, exc) -> None:
            counts["skipped"] += 2
            errors.append("[ERR] {}: {}".format(job.base, exc))
            print(errors[-1])

        # decode ahead / encode behind on host threads, one pair at a time on the device
        depth = max(1, min(plan.workers, 4))
        with ThreadPoolExecutor(max_workers=plan.workers) as pool:
            loads = {k: pool.submit(load, jobs[k]) for k in range(min(depth, len(jobs)))}
            writes = []
            for k, job in enumerate(jobs):
                nxt = k + depth
                if nxt < len(jobs):
                    loads[nxt] = pool.submit(load, jobs[nxt])
                try:
                    out = renderer.render(job, *loads.pop(k).result())
                    writes.append((job, pool.submit(finish, job, out)))
                except Exception as exc:
                    fail(job, exc)
                while writes and (len(writes) > depth or writes[0][1].done()):
                    wjob, fut = writes.pop(0)
                    try:
                        report(wjob, fut.result())
                    except Exception as exc:
                        fail(wjob, exc)
            for wjob, fut in writes:
                try:
                    report(wjob, fut.result())
                except Exception as exc:
                    fail(wjob, exc)

    if plan.extrinsics_xml is not None:
        try:
            if plan.metadata_only:
                resolved, done_bases = label_pairs, {rec[1] for rec in label_pairs}
            else:
                resolved = [(j.index, j.base, j.x_path, j.y_path, j.sensor_x, j.sensor_y) for j in jobs]
                done_bases = ok_bases
            export_camera_metadata(plan, resolved, done_bases, specs, persp_ext)
        except Exception as exc:
            errors.append("[ERR] perspective camera metadata export failed ({})".format(exc))
            print(errors[-1], file=sys.stderr)
    print("[DONE] processed={} skipped={} total={} persp_outputs={} mask_outputs={} color_outputs={} errors={}".format(
        counts["processed"], counts["skipped"], 2 * len(plan.pairs), counts["persp"], counts["mask"], counts["color"],
        len(errors)))
    return 2 if errors else 0


def export_camera_metadata(plan: RunPlan, resolved_pairs, ok_bases, specs, persp_ext: str) -> None:
    """DF:1599-1686: poses of every written view from the aligned fisheye cameras, then the Metashape XML and the
    COLMAP text model next to the images (``[DRY][META]`` line only on a dry run)."""
    from . import pose_export
    a = plan.args
    if bool(a.no_perspective) and not plan.metadata_only:
        raise ValueError("--camera-extrinsics-xml requires perspective output to be enabled.")
    if not plan.extrinsics_xml.is_file():
        raise ValueError("Camera extrinsics XML not found: {}".format(plan.extrinsics_xml))
    class _LensChoice(dict):
        """(sensor X, sensor Y) -> {view id: "X" | "Y"}: the remap's own lens choice, made on the device the first
        time a sensor pair is asked for (the reference reads it from its map cache, DF:1380-1397)."""
        def get(self, key, default=None):
            if key not in self:
                _views, _cals, info = dfh.choose_lenses(plan.sensors[key[0]], plan.sensors[key[1]], specs,
                                                        float(a.lens_x_yaw_deg), float(a.lens_y_yaw_deg),
                                                        float(a.lens_fov_deg))
                self[key] = {vid: rec["lens_key"] for vid, rec in info.items()}
            return self[key]

    lens_keys = _LensChoice()
    frames = pose_export.perspective_pose_frames(pose_export.camera_transform_map(plan.extrinsics_xml), resolved_pairs,
                                                 ok_bases, specs, lens_keys, persp_ext, float(a.lens_x_yaw_deg),
                                                 float(a.lens_y_yaw_deg))
    cameras, images = pose_export.colmap_model(frames, int(a.perspective_size), float(a.perspective_focal_mm),
                                               str(a.perspective_sensor_mm))
    points = []
    if plan.pointcloud is not None:
        if not plan.pointcloud.is_file():
            raise ValueError("Point cloud PLY not found: {}".format(plan.pointcloud))
        points = pose_export.colmap_points_from_ply(plan.pointcloud)
    out_xml = plan.persp_root / str(a.perspective_metashape_xml_name)
    out_colmap = plan.persp_root / "Sparse" / "0"
    if a.dry_run:
        print("[DRY][META] frames={} images={} xml={} colmap={} masks={} points={}".format(
            len(frames), plan.images_dir, out_xml, out_colmap, plan.masks_dir, len(points)))
        return
    pose_export.write_metashape_perspective_xml(out_xml, cameras, images)
    pose_export.write_colmap_text_model(out_colmap, cameras, images, points)
    print("[OK] Perspective images root: {}".format(plan.images_dir))
    print("[OK] Perspective Metashape XML: {}".format(out_xml))
    print("[OK] Perspective COLMAP text: {} (images={}, points={})".format(out_colmap, len(images), len(points)))
    print("[OK] Perspective masks root: {}".format(plan.masks_dir))


def main(argv: Optional[Sequence[str]] = None) -> int:
    args = create_arg_parser().parse_args(argv)
    try:
        plan = resolve_plan(args)
        announce(plan)
        return run(plan)
    except UsageError as exc:
        print("[ERR] {}".format(exc), file=sys.stderr)
        return 1


if __name__ == "__main__":
    sys.exit(main())
