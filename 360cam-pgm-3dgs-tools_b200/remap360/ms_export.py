"""Virtual perspective cameras for the ERP cutter's views: the pose exporters of the reference's
``gs360_MS360xmlToPersCams.py`` (drop-in: same flags, same files, same log lines).

A Metashape project aligned on spherical (360) images is turned into one pinhole camera per (image, view of the
cutter's preset): every view's rotation is the cutter's own yaw / pitch (MS:348-353, the convention the remap kernels
implement -- tests/test_oracle_geometry.py), so the poses written here describe exactly the images
``gs360_360PerspCut`` cuts on the GPU.  Outputs (MS:987-1249): ``transforms.json`` (OpenGL camera axes),
COLMAP text model, RealityScan XMP side-cars, Metashape cameras XML, and the rotated point cloud.

Pure host code: nothing here touches the device.  Matrix helpers, the Metashape camera loader and the PLY reader are
shared with the dual-fisheye tool's exporter (``pose_export.py``, which follows the same reference functions).
Not provided: the ``metashape-multi-camera-system`` format (MS:1250-1797, a template-driven rig export); asking for
it is an error here.

Parity: ``tests/test_ms_export.py`` replays whole ``main()`` runs recorded from the reference
(tests/golden/ms_export.json) -- transcripts and files byte for byte."""

from __future__ import annotations

import argparse
import json
import math
import pathlib
import struct
import subprocess
import sys
import xml.etree.ElementTree as ET
from typing import Callable, Dict, List, Optional, Sequence, Tuple

from . import pose_export as pe

Mat = List[List[float]]

PRESET_CHOICES = ["default", "fisheyelike", "full360coverage", "2views", "evenMinus30", "evenPlus30", "cube105"]
FORMAT_MULTI = "metashape-multi-camera-system"
SENSOR_MM = 36.0                       # the cutter's square 36 mm sensor (MS:49-50)
EXTRA_PITCH_DEG = 30.0                 # the _U / _D companions of a view (MS:55)
TRANSFORMS_X_FIX_DEG = 270.0           # transforms.json is written in a frame turned about X (MS:57)
COLMAP_X_BASE_DEG = 0.0
POINTCLOUD_PLY_X_DEG = 180.0           # only reported in the log (MS:59)
REALITYSCAN_AXES: Mat = [[1.0, 0.0, 0.0], [0.0, 0.0, -1.0], [0.0, 1.0, 0.0]]
REALITYSCAN_DIR = "cameras_RealityScan"

# preset -> (views around the horizon, focal length in mm or None, output size, letters dropped, letters that get
# _U / _D companions, pitch of every second view or None, horizontal FOV that defines the focal length or None)
_RING_PRESETS: Dict[str, Tuple[int, Optional[float], int, str, str, Optional[float], Optional[float]]] = {
    "default": (8, 12.0, 1600, "", "", None, None),
    "fisheyelike": (10, 17.0, 1600, "CDHI", "AF", None, None),
    "full360coverage": (8, 14.0, 1600, "BDFH", "BDFH", None, None),
    "2views": (8, 6.0, 3600, "BCDFGH", "", None, None),
    "evenMinus30": (8, 12.0, 1600, "", "", -30.0, None),
    "evenPlus30": (8, 12.0, 1600, "", "", 30.0, None),
}
_CUBE_VIEWS = [("A", 0.0, 0.0), ("B", 90.0, 0.0), ("C", 180.0, 0.0), ("D", -90.0, 0.0), ("E", 0.0, 90.0), ("F", 0.0, -90.0)]
_CUBE_HFOV_DEG = 105.0


# ---- view sets and intrinsics (MS:240-257, :588-720) -----------------------------------------------------------

def _tag(idx: int) -> str:
    return chr(ord("A") + idx) if idx < 26 else "%02d" % (idx + 1)


def _wrap_deg(angle: float) -> float:
    angle = ((angle + 180.0) % 360.0) - 180.0
    return 180.0 if abs(angle + 180.0) < 1e-6 else angle


def _companion_suffix(delta: float) -> str:
    sign, mag = ("_U" if delta > 0 else "_D"), abs(delta)
    if abs(mag - EXTRA_PITCH_DEG) < 1e-6:
        return sign
    return "%s%d" % (sign, int(round(mag))) if float(mag).is_integer() else "%s%g" % (sign, mag)


def preset_size_and_focal(preset: str) -> Tuple[int, float]:
    """Output size (square) and focal length in mm of a preset."""
    if preset == "cube105":
        return 1600, SENSOR_MM / (2.0 * math.tan(math.radians(_CUBE_HFOV_DEG) / 2.0))
    if preset not in _RING_PRESETS:
        raise ValueError("unknown preset: " + preset)
    _count, focal, size, _drop, _extra, _even, _hfov = _RING_PRESETS[preset]
    return size, float(focal)


def build_views(preset: str) -> List[Tuple[str, float, float]]:
    """(view id, yaw, pitch) in the cutter's naming and order."""
    if preset == "cube105":
        return list(_CUBE_VIEWS)
    if preset not in _RING_PRESETS:
        raise ValueError("unknown preset: " + preset)
    count, _focal, _size, drop, extra, even_pitch, _hfov = _RING_PRESETS[preset]
    views = []
    for idx in range(count):
        tag, yaw = _tag(idx), _wrap_deg(idx * (360.0 / float(count)))
        pitch = float(even_pitch) if even_pitch is not None and (idx + 1) % 2 == 0 else 0.0
        if tag not in drop:
            views.append((tag, yaw, pitch))
        if tag in extra:
            for delta in (EXTRA_PITCH_DEG, -EXTRA_PITCH_DEG):
                views.append((tag + _companion_suffix(delta), yaw, max(-90.0, min(90.0, pitch + delta))))
    return views


def compute_intrinsics(focal_mm: float, width: int, height: int):
    """(fl_x, fl_y, cx, cy, hfov, vfov) of the pinhole camera of a view."""
    fl_x, fl_y = focal_mm / (SENSOR_MM / float(width)), focal_mm / (SENSOR_MM / float(height))
    fov = math.degrees(2.0 * math.atan(SENSOR_MM / (2.0 * focal_mm)))
    return fl_x, fl_y, width * 0.5, height * 0.5, fov, fov


# ---- frames ------------------------------------------------------------------------------------------------------

def axis_angle_matrix(axis: Sequence[float], deg: float) -> Mat:
    """Rodrigues rotation about `axis` (identity for a zero axis or angle), MS:314-334."""
    x, y, z = axis
    norm = math.sqrt(x * x + y * y + z * z)
    if norm <= 0.0 or abs(deg) < 1e-12:
        return [row[:] for row in pe.IDENTITY3]
    x, y, z = x / norm, y / norm, z / norm
    c, s = math.cos(math.radians(deg)), math.sin(math.radians(deg))
    t = 1.0 - c
    return [[t * x * x + c, t * x * y - s * z, t * x * z + s * y],
            [t * x * y + s * z, t * y * y + c, t * y * z - s * x],
            [t * x * z - s * y, t * y * z + s * x, t * z * z + c]]


def _base_name(label: str, view_ids: Sequence[str]) -> str:
    """Image stem of a camera label: a trailing _<view id> is dropped (labels of already-cut images), path separators
    become underscores (MS:755-771)."""
    upper = str(label).upper()
    for vid in sorted({str(v).upper() for v in view_ids}, key=len, reverse=True):
        if upper.endswith("_" + vid):
            label = label[:-(len(vid) + 1)]
            break
    return label.replace("\\", "_").replace("/", "_").strip()


def build_frames(cameras, preset: str, ext: str, scale: float, world: Mat, axis: Sequence[float], deg: float,
                 log: Callable[[str], None] = print):
    """One frame per (camera, view): camera-to-world in GL and CV axes (MS:1800-1875), plus the shared intrinsics
    (fl_x, fl_y, cx, cy, width, height)."""
    views = build_views(preset)
    size, focal_mm = preset_size_and_focal(preset)
    fl_x, fl_y, cx, cy, hfov, vfov = compute_intrinsics(focal_mm, size, size)
    log("[INFO] preset={} views={} focal_mm={}".format(preset, len(views), focal_mm))
    log("[INFO] intrinsics: size={}x{} hfov={:.2f} vfov={:.2f}".format(size, size, hfov, vfov))
    log("[INFO] scale factor: {:.6g}".format(scale))
    log("[INFO] WORLD_FROM_METASHAPE axis=({:.6f} {:.6f} {:.6f}) deg={:.3f}".format(axis[0], axis[1], axis[2], deg))
    log("[INFO] WORLD_FROM_METASHAPE matrix:")
    for row in world:
        log("       " + " ".join("{: .6f}".format(v) for v in row))
    log("[INFO] transforms X fix: +{:.1f} deg".format(TRANSFORMS_X_FIX_DEG))
    log("[INFO] colmap X base: +{:.1f} deg".format(COLMAP_X_BASE_DEG))
    log("[INFO] pointcloud ply X: +{:.1f} deg".format(POINTCLOUD_PLY_X_DEG))
    ids = [vid for vid, _y, _p in views]
    frames = []
    for _cid, label, mat in cameras:
        base = _base_name(label, ids)
        scaled = [row[:] for row in mat]
        for r in range(3):
            scaled[r][3] *= scale
        base_gl = pe.matmul(pe.matmul(world, scaled), pe.CV_TO_GL)
        for vid, yaw, pitch in views:
            c2w_gl = pe.matmul(base_gl, pe.rigid(pe.yaw_pitch_to_rot_gl(yaw, pitch)))
            frames.append({"file_path": "{}_{}.{}".format(base, vid, ext), "c2w_gl": c2w_gl, "c2w_cv": pe.matmul(c2w_gl, pe.CV_TO_GL),
                           "source_name": base, "view_id": vid})
    return frames, (fl_x, fl_y, cx, cy, size, size)


def _x_fixed(c2w_gl: Mat, deg: Optional[float]) -> Mat:
    if deg is None or abs(deg) < 1e-6:
        return c2w_gl
    return pe.matmul(pe.rigid(pe.rot_x_deg(deg)), c2w_gl)


def colmap_pose(frame, x_fix_deg: float):
    """World-to-camera rotation and translation of a frame (MS:393-400)."""
    return pe.colmap_pose(_x_fixed(frame["c2w_gl"], x_fix_deg))


def colmap_images(frames, x_fix_deg: float):
    images = []
    for image_id, frame in enumerate(frames, start=1):
        r_wc, t = colmap_pose(frame, x_fix_deg)
        qw, qx, qy, qz = pe.rotmat_to_quat_wxyz(r_wc)
        images.append({"image_id": image_id, "qw": qw, "qx": qx, "qy": qy, "qz": qz, "tx": t[0], "ty": t[1], "tz": t[2],
                       "name": frame["file_path"]})
    return images


# ---- writers -------------------------------------------------------------------------------------------------------

def write_transforms_json(path: pathlib.Path, frames, intrinsics, x_fix_deg: float = 0.0) -> None:
    fl_x, fl_y, cx, cy, width, height = intrinsics
    payload = {"camera_model": "OPENCV", "fl_x": fl_x, "fl_y": fl_y, "cx": cx, "cy": cy, "w": width, "h": height,
               "k1": 0.0, "k2": 0.0, "p1": 0.0, "p2": 0.0,
               "frames": [{"file_path": f["file_path"], "transform_matrix": _x_fixed(f["c2w_gl"], x_fix_deg)} for f in frames]}
    with path.open("w", encoding="utf-8") as fp:
        json.dump(payload, fp, indent=2)


def write_colmap(out_dir: pathlib.Path, images, intrinsics, points) -> None:
    """cameras.txt (one PINHOLE camera), images.txt (an empty observation line per image), points3D.txt."""
    out_dir.mkdir(parents=True, exist_ok=True)
    fl_x, fl_y, cx, cy, width, height = intrinsics
    with (out_dir / "cameras.txt").open("w", encoding="utf-8") as fp:
        fp.write("# Camera list with one line of data per camera:\n#   CAMERA_ID, MODEL, WIDTH, HEIGHT, PARAMS[]\n"
                 "# Number of cameras: 1\n")
        fp.write("1 PINHOLE {} {} {:.12g} {:.12g} {:.12g} {:.12g}\n".format(width, height, fl_x, fl_y, cx, cy))
    with (out_dir / "images.txt").open("w", encoding="utf-8") as fp:
        fp.write("# Image list with two lines of data per image:\n#   IMAGE_ID, QW, QX, QY, QZ, TX, TY, TZ, CAMERA_ID, NAME\n"
                 "#   POINTS2D[] as (X, Y, POINT3D_ID)\n")
        fp.write("# Number of images: {}, mean observations per image: 0\n".format(len(images)))
        for im in images:
            fp.write("{image_id} {qw:.12g} {qx:.12g} {qy:.12g} {qz:.12g} {tx:.12g} {ty:.12g} {tz:.12g} 1 {name}\n\n".format(**im))
    with (out_dir / "points3D.txt").open("w", encoding="utf-8") as fp:
        fp.write("# 3D point list with one line of data per point:\n"
                 "#   POINT3D_ID, X, Y, Z, R, G, B, ERROR, TRACK[] as (IMAGE_ID, POINT2D_IDX)\n")
        fp.write("# Number of points: {}, mean track length: 0\n".format(len(points)))
        for pt in points:
            fp.write("{id} {x:.12g} {y:.12g} {z:.12g} {r} {g} {b} {err:.6g}\n".format(**pt))


_XMP = """<x:xmpmeta xmlns:x="adobe:ns:meta/">
  <rdf:RDF xmlns:rdf="http://www.w3.org/1999/02/22-rdf-syntax-ns#">
    <rdf:Description xcr:Version="3" xcr:PosePrior="initial" xcr:Coordinates="absolute"
       xcr:DistortionModel="perspective" xcr:DistortionCoeficients="0 0 0 0 0 0"
       xcr:FocalLength35mm="{focal:g}" xcr:Skew="0" xcr:AspectRatio="1" xcr:PrincipalPointU="0"
       xcr:PrincipalPointV="0" xcr:CalibrationPrior="initial" xcr:CalibrationGroup="0"
       xcr:DistortionGroup="0" xcr:InTexturing="1" xcr:InMeshing="1" xmlns:xcr="http://www.capturingreality.com/ns/xcr/1.1#">
      <xcr:Rotation>{rotation}</xcr:Rotation>
      <xcr:Position>{position}</xcr:Position>
    </rdf:Description>
  </rdf:RDF>
</x:xmpmeta>
"""


def write_realityscan_xmp(out_dir: pathlib.Path, frames, intrinsics, x_fix_deg: float = 0.0,
                          log: Callable[[str], None] = print) -> None:
    """One <stem>.xmp per frame: rotation and camera centre in RealityScan's axes (Z up), MS:1087-1132."""
    out_dir.mkdir(parents=True, exist_ok=True)
    focal_mm = intrinsics[0] * (SENSOR_MM / float(intrinsics[4]))
    axes_t = pe.transpose3(REALITYSCAN_AXES)
    for frame in frames:
        r_wc, t = colmap_pose(frame, x_fix_deg)
        centre = pe.matvec3(pe.transpose3(r_wc), [-t[0], -t[1], -t[2]])
        rotation = " ".join("{:.15g}".format(v) for row in pe.matmul(r_wc, REALITYSCAN_AXES) for v in row)
        position = " ".join("{:.15g}".format(v) for v in pe.matvec3(axes_t, centre))
        stem = pathlib.Path(frame["file_path"]).stem
        with (out_dir / (stem + ".xmp")).open("w", encoding="utf-8") as fp:
            fp.write(_XMP.format(focal=focal_mm, rotation=rotation, position=position))
    log("[OK] RealityScan XMP: {}".format(out_dir))


def write_metashape_xml(xml_in: pathlib.Path, out_path: pathlib.Path, frames, intrinsics, preset: str) -> None:
    """A Metashape document with one frame sensor and one camera per frame (<transform> = camera-to-world in CV axes);
    data type / black level / sensitivity are carried over from the project's first sensor (MS:1135-1247)."""
    fl_x, _fl_y, _cx, _cy, width, height = intrinsics
    carried = {"data_type": "uint8", "black_level": "0 0 0", "sensitivity": "1 1 1"}
    chunk_in = ET.parse(str(xml_in)).getroot().find("chunk")
    sensors_in = chunk_in.find("sensors") if chunk_in is not None else None
    first = sensors_in.find("sensor") if sensors_in is not None else None
    if first is not None:
        for key in carried:
            carried[key] = (first.findtext(key) or carried[key]).strip()
    res = {"width": str(int(width)), "height": str(int(height))}
    doc = ET.Element("document", {"version": "1.2.0"})
    chunk = ET.SubElement(doc, "chunk", {"label": "unknown", "enabled": "true"})
    sensor = ET.SubElement(ET.SubElement(chunk, "sensors", {"next_id": "1"}), "sensor",
                           {"id": "0", "label": "virtual_" + preset, "type": "frame"})
    ET.SubElement(sensor, "resolution", res)
    ET.SubElement(sensor, "property", {"name": "layer_index", "value": "0"})
    calib = ET.SubElement(sensor, "calibration", {"type": "frame", "class": "initial"})
    ET.SubElement(calib, "resolution", res)
    ET.SubElement(calib, "f").text = "{:.6f}".format(fl_x)
    for name in ("cx", "cy", "k1", "k2", "p1", "p2"):
        ET.SubElement(calib, name).text = "0"
    for key in ("data_type", "black_level", "sensitivity"):
        ET.SubElement(sensor, key).text = carried[key]
    comps = ET.SubElement(chunk, "components", {"next_id": "1", "active_id": "0"})
    ET.SubElement(ET.SubElement(comps, "component", {"id": "0", "label": "Component 1"}), "partition")
    cams = ET.SubElement(chunk, "cameras", {"next_id": str(len(frames)), "next_group_id": "0"})
    for idx, frame in enumerate(frames):
        cam = ET.SubElement(cams, "camera", id=str(idx), sensor_id="0", component_id="0", label=pathlib.Path(frame["file_path"]).stem)
        ET.SubElement(cam, "transform").text = " ".join("{:.15g}".format(v) for row in frame["c2w_cv"] for v in row)
    ET.ElementTree(doc).write(str(out_path), encoding="UTF-8", xml_declaration=True)


def write_ply(path: pathlib.Path, rows, names: Sequence[str]) -> None:
    """Binary little-endian PLY: x / y / z as float32, everything else as uchar (MS:891-919)."""
    rec = struct.Struct("<" + "".join("f" if n in ("x", "y", "z") else "B" for n in names))
    with path.open("wb") as fp:
        fp.write(b"ply\nformat binary_little_endian 1.0\n")
        fp.write("element vertex {}\n".format(len(rows)).encode("ascii"))
        for n in names:
            fp.write(("property %s %s\n" % ("float" if n in ("x", "y", "z") else "uchar", n)).encode("ascii"))
        fp.write(b"end_header\n")
        for row in rows:
            fp.write(rec.pack(*[row[n] for n in names]))


def points_from_ply(ply_path: pathlib.Path, out_dir: pathlib.Path, world: Mat, ply_x_deg: float, scale: float,
                    write_transforms_ply: bool = True, log: Callable[[str], None] = print):
    """COLMAP points of the project's point cloud (world rotation, scale) and, for transforms.json users, the cloud
    once more turned about X as ``pointcloud_for_transforms.ply`` (MS:922-984)."""
    rows, names = pe.read_ply_vertices(ply_path)
    coloured = all(c in names for c in ("red", "green", "blue"))
    rot_world = pe.rotation_part(world)
    rot_ply = pe.rot_x_deg(ply_x_deg) if abs(ply_x_deg) > 1e-6 else None
    points, cloud = [], []
    for idx, v in enumerate(rows, start=1):
        w = pe.matvec3(rot_world, [float(v.get("x", 0.0)), float(v.get("y", 0.0)), float(v.get("z", 0.0))])
        p = pe.matvec3(rot_ply, w) if rot_ply is not None else list(w)
        w = [c * scale for c in w]
        p = [c * scale for c in p]
        r, g, b = (int(v.get("red", 128)), int(v.get("green", 128)), int(v.get("blue", 128))) if coloured else (128, 128, 128)
        points.append({"id": idx, "x": w[0], "y": w[1], "z": w[2], "r": r, "g": g, "b": b, "err": 0.0})
        row = {"x": p[0], "y": p[1], "z": p[2]}
        if coloured:
            row.update({"red": r, "green": g, "blue": b})
        cloud.append(row)
    if write_transforms_ply:
        out_ply = out_dir / "pointcloud_for_transforms.ply"
        write_ply(out_ply, cloud, ["x", "y", "z", "red", "green", "blue"] if coloured else ["x", "y", "z"])
        log("[OK] Rotated pointcloud: {}".format(out_ply))
    return points


# ---- command line (MS:1878-2177) -------------------------------------------------------------------------------------

def build_arg_parser() -> argparse.ArgumentParser:
    ap = argparse.ArgumentParser(description="Convert Metashape 360 XML to virtual camera transforms.",
                                 formatter_class=argparse.ArgumentDefaultsHelpFormatter)
    ap.add_argument("xml", help="Metashape cameras_XML.xml path")
    ap.add_argument("--preset", choices=PRESET_CHOICES, default="full360coverage",
                    help="Virtual camera preset (matches gs360_360PerspCut)")
    ap.add_argument("-o", "--out", default=None, help="Output directory (default: <xml_dir>/perspective_cams)")
    ap.add_argument("--format", choices=["transforms", "colmap", "metashape", FORMAT_MULTI, "realityscan", "all"], default="metashape",
                    help="Output format (all=transforms+colmap+metashape+realityscan)")
    ap.add_argument("--ext", default="jpg", help="Image extension for file paths (without dot)")
    ap.add_argument("--scale", type=float, default=1.0, help="Global scale factor applied to translations")
    ap.add_argument("--world-rot-axis", default="0 1 0", help="World rotation axis (x y z) for Metashape->PostShot")
    ap.add_argument("--world-rot-deg", type=float, default=0.0, help="World rotation angle in degrees for Metashape->PostShot")
    ap.add_argument("--persp-cut", dest="cut", action="store_true", help="Run gs360_360PerspCut.py to cut perspective images")
    ap.add_argument("--cut", dest="cut", action="store_true", help=argparse.SUPPRESS)
    ap.add_argument("--cut-input", default=None, help="Input folder for equirectangular images (default: <xml_dir>/360imgs)")
    ap.add_argument("--cut-out", default=None, help="Output folder for cut images (default: tool's own default)")
    ap.add_argument("--points-ply", default=None,
                    help="Input pointcloud PLY (required for --format colmap; optional for transforms to write rotated PLY)")
    ap.add_argument("--pc-rotate-x-plus180", dest="pc_rotate_x_deg", action="store_const", const=180.0, default=0.0,
                    help="Rotate output pointcloud PLY around X by +180 degrees")
    ap.add_argument("--pc-rotate-x-plus90", dest="pc_rotate_x_deg", action="store_const", const=90.0, help=argparse.SUPPRESS)
    ap.add_argument("--pc-rotate-x-minus90", dest="pc_rotate_x_deg", action="store_const", const=-90.0, help=argparse.SUPPRESS)
    return ap


def parse_axis(text: str) -> List[float]:
    parts = text.replace(",", " ").split()
    if len(parts) != 3:
        raise ValueError("axis must have 3 values (x y z)")
    return [float(p) for p in parts]


def cut_command(preset: str, cut_in: pathlib.Path, cut_out: Optional[pathlib.Path], tool: pathlib.Path) -> List[str]:
    """argv of the cutter run behind --persp-cut (MS:2019-2049): the preset by name, cube105 spelled out."""
    cmd = [sys.executable, str(tool), "-i", str(cut_in)]
    cmd += ["--count", "4", "--hfov", str(_CUBE_HFOV_DEG), "--add-top", "--add-bottom"] if preset == "cube105" else ["--preset", preset]
    return cmd + (["-o", str(cut_out)] if cut_out is not None else [])


def main(argv: Optional[Sequence[str]] = None) -> None:
    args = build_arg_parser().parse_args(argv)
    err = lambda *a: print(*a, file=sys.stderr)
    if args.format == FORMAT_MULTI and args.preset != "fisheyelike":
        err("[ERR] --format metashape-multi-camera-system requires --preset fisheyelike")
        sys.exit(1)
    if args.format == FORMAT_MULTI:
        err("[ERR] --format metashape-multi-camera-system is not available in the CUDA tool set (template-driven rig export)")
        sys.exit(1)
    xml_path = pathlib.Path(args.xml).expanduser().resolve()
    if not xml_path.exists():
        err("[ERR] XML not found:", xml_path)
        sys.exit(1)
    out_dir = pathlib.Path(args.out).expanduser().resolve() if args.out else xml_path.parent / "perspective_cams"
    out_dir.mkdir(parents=True, exist_ok=True)
    axis = parse_axis(args.world_rot_axis)
    world = pe.rigid(axis_angle_matrix(axis, args.world_rot_deg))
    cameras = pe.load_metashape_cameras(xml_path)
    if not cameras:
        err("[WARN] No camera transforms found")
        sys.exit(1)
    frames, intrinsics = build_frames(cameras, args.preset, args.ext.lstrip("."), args.scale, world, axis, args.world_rot_deg)
    if args.format in ("transforms", "all"):
        write_transforms_json(out_dir / "transforms.json", frames, intrinsics, x_fix_deg=TRANSFORMS_X_FIX_DEG)
        print("[OK] transforms.json:", out_dir / "transforms.json")
    points = []
    needs_colmap = args.format in ("colmap", "all")
    if needs_colmap and not args.points_ply:
        err("[ERR] --points-ply is required when --format includes colmap")
        sys.exit(1)
    if args.points_ply and args.format in ("transforms", "colmap", "all"):
        ply = pathlib.Path(args.points_ply).expanduser().resolve()
        if not ply.exists():
            err("[ERR] points PLY not found: {}".format(ply))
            sys.exit(1)
        points = points_from_ply(ply, out_dir, world, args.pc_rotate_x_deg, args.scale,
                                 write_transforms_ply=args.format in ("transforms", "all"))
    if needs_colmap:
        write_colmap(out_dir / "sparse" / "0", colmap_images(frames, COLMAP_X_BASE_DEG), intrinsics, points)
        print("[OK] COLMAP text:", out_dir / "sparse" / "0")
    if args.format in ("realityscan", "all"):
        write_realityscan_xmp(out_dir / REALITYSCAN_DIR, frames, intrinsics, COLMAP_X_BASE_DEG)
    if args.format in ("metashape", "all"):
        write_metashape_xml(xml_path, out_dir / "perspective_cams.xml", frames, intrinsics, args.preset)
        print("[OK] Metashape cameras XML:", out_dir / "perspective_cams.xml")
    if args.cut:
        cut_in = pathlib.Path(args.cut_input).expanduser().resolve() if args.cut_input else xml_path.parent / "360imgs"
        if not cut_in.exists():
            raise ValueError("cut input not found: {}".format(cut_in))
        cut_out = pathlib.Path(args.cut_out).expanduser().resolve() if args.cut_out else None
        tool = pathlib.Path(__file__).resolve().parent.parent / "gs360_360PerspCut.py"
        if not tool.exists():
            raise ValueError("gs360_360PerspCut.py not found: {}".format(tool))
        cmd = cut_command(args.preset, cut_in, cut_out, tool)
        print("[INFO] Running cut: " + " ".join(cmd))
        sys.stdout.flush()
        subprocess.run(cmd, check=True)
    print("[INFO] If you still need to cut images, run gs360_360PerspCut.py separately.")
