"""Camera poses of the cut views and their COLMAP / Metashape-XML export -- the bookkeeping behind
``--camera-extrinsics-xml`` / ``--pointcloud-ply`` / ``--metadata-only`` of the dual-fisheye tool.

Pure host code (no pixels are touched); it consumes the same view set and lens choice as the remap:

* aligned fisheye cameras from a Metashape ``cameras.xml`` -- chunk / component similarity applied to
  every camera-to-world matrix (MS:409-585);
* per cut view, ``c2w_gl = (c2w_cv . CV_TO_GL) . R_gl(yaw relative to the chosen lens, pitch)`` and back to
  the OpenCV convention (DF:1348-1461; ``R_gl = Ry(-yaw) . Rx(pitch)``, MS:292-353 -- the same camera the
  remap kernels render, tests/test_oracle_geometry.py holds the witness);
* one shared PINHOLE camera, world-to-camera quaternion + translation per image (DF:1464-1512, MS:393-400,
  quaternion extraction as the converter's, CameraFormatConverter:204-233);
* optional sparse points straight from a Metashape PLY (DF:1515-1533, MS:782-985 with identity world
  transform and unit scale);
* COLMAP text model and Metashape perspective XML writers (CameraFormatConverter:471-545, :938-1035).

Numbers are written with ``{:.12g}`` / ``{:.15g}``, so every matrix product below keeps the reference's
summation order (left to right over k) -- the files come out byte-identical, which is what the tests check
against files produced by running the reference (tests/golden/df_metadata.json)."""

from __future__ import annotations

import math
import pathlib
import struct
import xml.etree.ElementTree as ET
from typing import Dict, Iterable, List, Mapping, Optional, Sequence, Set, Tuple

from . import dualfisheye as dfh

Mat = List[List[float]]

CV_TO_GL: Mat = [[1.0, 0.0, 0.0, 0.0], [0.0, -1.0, 0.0, 0.0], [0.0, 0.0, -1.0, 0.0], [0.0, 0.0, 0.0, 1.0]]
IDENTITY3: Mat = [[1.0, 0.0, 0.0], [0.0, 1.0, 0.0], [0.0, 0.0, 1.0]]


# ---- small dense algebra on nested lists ------------------------------------------------------------------

def matmul(a: Mat, b: Mat) -> Mat:
    n = len(a)
    return [[sum(a[i][k] * b[k][j] for k in range(n)) for j in range(n)] for i in range(n)]


def transpose3(a: Mat) -> Mat:
    return [[a[c][r] for c in range(3)] for r in range(3)]


def matvec3(a: Mat, v: Sequence[float]) -> List[float]:
    return [a[r][0] * v[0] + a[r][1] * v[1] + a[r][2] * v[2] for r in range(3)]


def rigid(r: Mat, t: Sequence[float] = (0.0, 0.0, 0.0)) -> Mat:
    """3x3 rotation + translation -> 4x4."""
    return [[r[0][0], r[0][1], r[0][2], t[0]], [r[1][0], r[1][1], r[1][2], t[1]], [r[2][0], r[2][1], r[2][2], t[2]],
            [0.0, 0.0, 0.0, 1.0]]


def rotation_part(m: Mat) -> Mat:
    return [list(m[r][:3]) for r in range(3)]


def rot_x_deg(deg: float) -> Mat:
    c, s = math.cos(math.radians(deg)), math.sin(math.radians(deg))
    return [[1.0, 0.0, 0.0], [0.0, c, -s], [0.0, s, c]]


def rot_y_deg(deg: float) -> Mat:
    c, s = math.cos(math.radians(deg)), math.sin(math.radians(deg))
    return [[c, 0.0, s], [0.0, 1.0, 0.0], [-s, 0.0, c]]


def yaw_pitch_to_rot_gl(yaw_deg: float, pitch_deg: float) -> Mat:
    """MS:348-353: positive yaw turns right, pitch is applied first."""
    return matmul(rot_y_deg(-float(yaw_deg)), rot_x_deg(float(pitch_deg)))


def rotmat_to_quat_wxyz(r: Mat) -> Tuple[float, float, float, float]:
    """Largest-pivot extraction, normalised (CameraFormatConverter:204-233)."""
    trace = r[0][0] + r[1][1] + r[2][2]
    if trace > 0.0:
        s = math.sqrt(trace + 1.0) * 2.0
        q = (0.25 * s, (r[2][1] - r[1][2]) / s, (r[0][2] - r[2][0]) / s, (r[1][0] - r[0][1]) / s)
    elif r[0][0] > r[1][1] and r[0][0] > r[2][2]:
        s = math.sqrt(1.0 + r[0][0] - r[1][1] - r[2][2]) * 2.0
        q = ((r[2][1] - r[1][2]) / s, 0.25 * s, (r[0][1] + r[1][0]) / s, (r[0][2] + r[2][0]) / s)
    elif r[1][1] > r[2][2]:
        s = math.sqrt(1.0 + r[1][1] - r[0][0] - r[2][2]) * 2.0
        q = ((r[0][2] - r[2][0]) / s, (r[0][1] + r[1][0]) / s, 0.25 * s, (r[1][2] + r[2][1]) / s)
    else:
        s = math.sqrt(1.0 + r[2][2] - r[0][0] - r[1][1]) * 2.0
        q = ((r[1][0] - r[0][1]) / s, (r[0][2] + r[2][0]) / s, (r[1][2] + r[2][1]) / s, 0.25 * s)
    n = math.sqrt(q[0] * q[0] + q[1] * q[1] + q[2] * q[2] + q[3] * q[3])
    if n <= 0.0:
        return (1.0, 0.0, 0.0, 0.0)
    return (q[0] / n, q[1] / n, q[2] / n, q[3] / n)


def quat_wxyz_to_rotmat(qw: float, qx: float, qy: float, qz: float) -> Mat:
    n = math.sqrt(qw * qw + qx * qx + qy * qy + qz * qz)
    if n <= 0.0:
        return [row[:] for row in IDENTITY3]
    qw, qx, qy, qz = qw / n, qx / n, qy / n, qz / n
    return [[1.0 - 2.0 * (qy * qy + qz * qz), 2.0 * (qx * qy - qz * qw), 2.0 * (qx * qz + qy * qw)],
            [2.0 * (qx * qy + qz * qw), 1.0 - 2.0 * (qx * qx + qz * qz), 2.0 * (qy * qz - qx * qw)],
            [2.0 * (qx * qz - qy * qw), 2.0 * (qy * qz + qx * qw), 1.0 - 2.0 * (qx * qx + qy * qy)]]


# ---- Metashape cameras.xml -------------------------------------------------------------------------------------

def _floats(text, count: Optional[int], what: str) -> List[float]:
    values = [float(tok) for tok in str(text or "").strip().split()]
    if count is not None and len(values) != count:
        raise ValueError("{} must have {} floats".format(what, count))
    return values


def parse_transform16(text) -> Mat:
    v = _floats(text, 16, "transform")
    return [v[0:4], v[4:8], v[8:12], v[12:16]]


def parse_similarity(node) -> Optional[Dict[str, object]]:
    """A Metashape <transform> as rotation / translation / uniform scale: either 16 numbers in the node's text
    (scale = mean row norm of the 3x3 block) or <rotation>/<translation>/<scale> children (MS:453-517)."""
    if node is None:
        return None
    text = (node.text or "").strip()
    if text:
        m = parse_transform16(text)
        block = [[float(m[r][c]) for c in range(3)] for r in range(3)]
        norms = [math.sqrt(sum(v * v for v in row)) for row in block]
        norms = [v for v in norms if v > 1e-12]
        scale = sum(norms) / float(len(norms)) if norms else 1.0
        if scale <= 1e-12:
            scale = 1.0
        return {"rotation": [[v / scale for v in row] for row in block],
                "translation": [m[0][3], m[1][3], m[2][3]], "scale": scale}
    rot_node, tr_node, sc_node = node.find("rotation"), node.find("translation"), node.find("scale")
    if rot_node is None and tr_node is None and sc_node is None:
        return None
    rot = [row[:] for row in IDENTITY3]
    if rot_node is not None and (rot_node.text or "").strip():
        v = _floats(rot_node.text, 9, "rotation")
        rot = [v[0:3], v[3:6], v[6:9]]
    tvec = [0.0, 0.0, 0.0]
    if tr_node is not None and (tr_node.text or "").strip():
        tvec = _floats(tr_node.text, 3, "translation")
    scale = 1.0
    if sc_node is not None and (sc_node.text or "").strip():
        v = _floats(sc_node.text, None, "scale")
        if not v:
            raise ValueError("scale is empty")
        if len(v) == 3:
            if max(abs(x - v[0]) for x in v[1:]) > 1e-9:
                raise ValueError("non-uniform scale is not supported")
        elif len(v) != 1:
            raise ValueError("scale must have 1 or 3 floats")
        scale = v[0]
    return {"rotation": rot, "translation": tvec, "scale": float(scale)}


def apply_similarity(sim: Mapping[str, object], cam: Mat) -> Mat:
    """world = s * R * camera-centre + t, rotation = R * camera rotation (MS:520-540)."""
    rot, t, s = sim["rotation"], sim["translation"], float(sim["scale"])
    centre = matvec3(rot, [cam[0][3], cam[1][3], cam[2][3]])
    return rigid(matmul(rot, rotation_part(cam)), [(s * centre[k]) + t[k] for k in range(3)])


def load_metashape_cameras(xml_path) -> List[Tuple[int, str, Mat]]:
    """(camera id, label, camera-to-world 4x4 in the OpenCV convention), sorted by id; disabled cameras and
    cameras without a transform are skipped; the chunk transform wins over the component's (MS:543-585)."""
    chunk = ET.parse(str(xml_path)).getroot().find("chunk")
    if chunk is None:
        raise ValueError("missing <chunk> in XML")
    cams_node = chunk.find("cameras")
    if cams_node is None:
        raise ValueError("missing <cameras> in XML")
    chunk_sim = parse_similarity(chunk.find("transform"))
    comp_sims = {}
    comps = chunk.find("components")
    if comps is not None:
        for comp in comps.findall("component"):
            cid = (comp.get("id") or "").strip()
            sim = parse_similarity(comp.find("transform")) if cid else None
            if sim is not None:
                comp_sims[cid] = sim
    out = []
    for cam in cams_node.findall("camera"):
        if (cam.get("enabled") or "").lower() == "false":
            continue
        tnode = cam.find("transform")
        if tnode is None or not (tnode.text or "").strip():
            continue
        mat = parse_transform16(tnode.text)
        sim = chunk_sim if chunk_sim is not None else comp_sims.get((cam.get("component_id") or "").strip())
        if sim is not None:
            mat = apply_similarity(sim, mat)
        out.append((int(cam.get("id", "0")), cam.get("label") or "camera_{}".format(cam.get("id", "0")), mat))
    out.sort(key=lambda rec: rec[0])
    return out


def camera_transform_map(xml_path) -> Dict[str, Mat]:
    """DF:966-972."""
    return {str(label): mat for _id, label, mat in load_metashape_cameras(xml_path)}


# ---- pairs, frames, COLMAP model -------------------------------------------------------------------------

ResolvedPair = Tuple[int, str, pathlib.Path, pathlib.Path, str, str]      # index, base, X path, Y path, sensor ids


def metadata_only_pairs(camera_to_sensor: Mapping[str, str], sensors: Mapping[str, object], x_suffix: str,
                        y_suffix: str, available_labels: Optional[Set[str]] = None) -> List[ResolvedPair]:
    """X/Y pairs named by the XML's camera labels alone (DF:917-963)."""
    table: Dict[str, Dict[str, Tuple[str, str]]] = {}
    for label, sid in sorted(camera_to_sensor.items()):
        if sid not in sensors or (available_labels is not None and label not in available_labels):
            continue
        if label.endswith(x_suffix):
            base, key = label[:-len(x_suffix)], "X"
        elif label.endswith(y_suffix):
            base, key = label[:-len(y_suffix)], "Y"
        else:
            continue
        table.setdefault(base, {})[key] = (label, sid)
    pairs: List[ResolvedPair] = []
    for idx, base in enumerate(sorted(table), start=1):
        x, y = table[base].get("X"), table[base].get("Y")
        if x is not None and y is not None:
            pairs.append((idx, base, pathlib.Path(x[0] + ".jpg"), pathlib.Path(y[0] + ".jpg"), x[1], y[1]))
    return pairs


def perspective_pose_frames(transforms: Mapping[str, Mat], pairs: Sequence[ResolvedPair],
                            ok_bases: Optional[Set[str]], specs: Sequence[Mapping[str, object]],
                            lens_keys: Mapping[Tuple[str, str], Mapping[str, str]], out_ext: str,
                            lens_x_yaw_deg: float, lens_y_yaw_deg: float) -> List[Dict[str, object]]:
    """One pose per (pair, view): DF:1348-1461.  ``lens_keys[(sensor_x, sensor_y)][view_id]`` is "X" or "Y"."""
    frames: List[Dict[str, object]] = []
    missing: List[str] = []
    for _idx, base, x_path, y_path, sid_x, sid_y in pairs:
        if ok_bases is not None and base not in ok_bases:
            continue
        x_c2w, y_c2w = transforms.get(x_path.stem), transforms.get(y_path.stem)
        if x_c2w is None:
            missing.append(x_path.stem)
            continue
        if y_c2w is None:
            missing.append(y_path.stem)
            continue
        chosen = lens_keys.get((sid_x, sid_y))
        if not chosen:
            raise ValueError("Perspective remap cache missing for sensor pair {} / {}".format(sid_x, sid_y))
        for spec in specs:
            view_id = str(spec["view_id"])
            if view_id not in chosen:
                raise ValueError("Perspective view '{}' missing from remap cache.".format(view_id))
            key = str(chosen[view_id]).upper()
            if key == "X":
                base_cv, label, lens_yaw = x_c2w, x_path.stem, float(lens_x_yaw_deg)
            elif key == "Y":
                base_cv, label, lens_yaw = y_c2w, y_path.stem, float(lens_y_yaw_deg)
            else:
                raise ValueError("Unsupported lens key '{}' for view '{}'.".format(key, view_id))
            yaw_rel = dfh.wrap_angle_deg(float(spec["yaw_deg"]) - lens_yaw)
            pitch = float(spec["pitch_deg"])
            c2w_gl = matmul(matmul(base_cv, CV_TO_GL), rigid(yaw_pitch_to_rot_gl(yaw_rel, pitch)))
            frames.append({"file_path": "{}_{}{}".format(base, view_id, out_ext), "c2w_gl": c2w_gl,
                           "c2w_cv": matmul(c2w_gl, CV_TO_GL), "source_name": base, "source_label": label,
                           "view_id": view_id, "lens_key": key, "yaw_rel_deg": yaw_rel, "pitch_deg": pitch})
    if missing:
        names = sorted(set(missing))
        raise ValueError("Missing camera transforms in extrinsics XML: {}".format(
            ", ".join(names[:8]) + (", ..." if len(names) > 8 else "")))
    if not frames:
        raise ValueError("No perspective pose frames could be generated.")
    return frames


def colmap_pose(c2w_gl: Mat) -> Tuple[Mat, List[float]]:
    """World-to-camera rotation and translation of a GL camera-to-world matrix (MS:393-400, no X fix)."""
    c2w_cv = matmul(c2w_gl, CV_TO_GL)
    r_wc = transpose3(rotation_part(c2w_cv))
    return r_wc, matvec3(r_wc, [-c2w_cv[0][3], -c2w_cv[1][3], -c2w_cv[2][3]])


def colmap_model(frames: Sequence[Mapping[str, object]], size: int, focal_mm: float, sensor_mm: str):
    """One PINHOLE camera shared by every view + one image record per frame (DF:1464-1512)."""
    w = h = int(size)
    sw, sh = dfh.parse_sensor_dimensions(sensor_mm)
    fx, fy = float(focal_mm) / (float(sw) / float(w)), float(focal_mm) / (float(sh) / float(h))
    cameras = [{"camera_id": 1, "model": "PINHOLE", "width": w, "height": h, "params": [fx, fy, w * 0.5, h * 0.5]}]
    images = []
    for image_id, frame in enumerate(frames, start=1):
        r_wc, t = colmap_pose(frame["c2w_gl"])
        qw, qx, qy, qz = rotmat_to_quat_wxyz(r_wc)
        images.append({"image_id": image_id, "qw": qw, "qx": qx, "qy": qy, "qz": qz, "tx": t[0], "ty": t[1], "tz": t[2],
                       "camera_id": 1, "name": str(frame["file_path"]), "points2d_line": ""})
    return cameras, images


# ---- Metashape PLY -> sparse points --------------------------------------------------------------------------

_PLY_TYPES = {"float": "f", "float32": "f", "double": "d", "float64": "d", "uchar": "B", "uint8": "B", "char": "b",
              "int8": "b", "short": "h", "int16": "h", "ushort": "H", "uint16": "H", "int": "i", "int32": "i",
              "uint": "I", "uint32": "I"}
_PLY_FLOATS = ("float", "float32", "double", "float64")


def read_ply_vertices(ply_path) -> Tuple[List[Dict[str, float]], List[str]]:
    """Vertex records of an ascii or little-endian binary PLY (MS:782-888); list properties are refused.
    Deliberate difference: the reference's ASCII branch keys each record by the property TYPE instead of its name
    (MS:866-871), which turns every ASCII point into (0, 0, 0) grey; here ASCII and binary files agree."""
    with pathlib.Path(ply_path).open("rb") as fp:
        fmt, count, props, in_vertex = None, 0, [], False
        while True:
            raw = fp.readline()
            if not raw:
                raise ValueError("unexpected EOF while reading PLY header")
            line = raw.decode("ascii", "ignore").strip()
            if line == "end_header":
                break
            parts = line.split()
            if line.startswith("format "):
                fmt = parts[1]
            elif line.startswith("element "):
                in_vertex = len(parts) >= 3 and parts[1] == "vertex"
                if in_vertex:
                    count = int(parts[2])
            elif line.startswith("property ") and in_vertex:
                if parts[1] == "list":
                    raise ValueError("PLY list properties are not supported")
                if len(parts) >= 3:
                    props.append((parts[1], parts[2]))
        if fmt is None:
            raise ValueError("PLY format not found")
        if fmt not in ("binary_little_endian", "ascii"):
            raise ValueError("unsupported PLY format: {}".format(fmt))
        for typ, _name in props:
            if typ not in _PLY_TYPES:
                raise ValueError("unsupported PLY type: {}".format(typ))
        names = [name for _typ, name in props]
        rows: List[Dict[str, float]] = []
        if fmt == "ascii":
            for _ in range(count):
                raw = fp.readline()
                if not raw:
                    raise ValueError("unexpected EOF in PLY vertices")
                toks = raw.decode("ascii", "ignore").strip().split()
                if len(toks) < len(names):
                    raise ValueError("invalid PLY vertex row")
                rows.append({name: (float(tok) if typ in _PLY_FLOATS else int(float(tok)))
                             for (typ, name), tok in zip(props, toks)})
        else:
            rec = struct.Struct("<" + "".join(_PLY_TYPES[typ] for typ, _ in props))
            blob = fp.read(rec.size * count)
            if len(blob) != rec.size * count:
                raise ValueError("unexpected EOF in PLY vertices")
            rows = [dict(zip(names, vals)) for vals in rec.iter_unpack(blob)]
    return rows, names


def colmap_points_from_ply(ply_path) -> List[Dict[str, object]]:
    """DF:1515-1533: the PLY's vertices as COLMAP points, identity world transform, unit scale, grey when the file
    has no colours."""
    rows, names = read_ply_vertices(pathlib.Path(ply_path))
    has_color = all(c in names for c in ("red", "green", "blue"))
    points = []
    for idx, v in enumerate(rows, start=1):
        x, y, z = matvec3(IDENTITY3, [float(v.get("x", 0.0)), float(v.get("y", 0.0)), float(v.get("z", 0.0))])
        r, g, b = (int(v.get("red", 128)), int(v.get("green", 128)), int(v.get("blue", 128))) if has_color else (128,) * 3
        points.append({"id": idx, "x": x * 1.0, "y": y * 1.0, "z": z * 1.0, "r": r, "g": g, "b": b, "err": 0.0})
    return points


# ---- writers -----------------------------------------------------------------------------------------------

def write_colmap_text_model(out_dir, cameras: Iterable[Mapping[str, object]], images: Sequence[Mapping[str, object]],
                            points: Sequence[Mapping[str, object]]) -> None:
    """cameras.txt / images.txt / points3D.txt with COLMAP's comment headers (CameraFormatConverter:471-545)."""
    out_dir = pathlib.Path(out_dir)
    out_dir.mkdir(parents=True, exist_ok=True)
    cameras = list(cameras)
    mean_obs = 0.0
    if images:
        mean_obs = sum(len((im.get("points2d_line", "") or "").split()) // 3 for im in images) / float(len(images))
    mean_track = 0.0
    if points:
        mean_track = sum(len(pt.get("track_tokens", []) or []) // 2 for pt in points) / float(len(points))
    with (out_dir / "cameras.txt").open("w", encoding="utf-8") as f:
        f.write("# Camera list with one line of data per camera:\n#   CAMERA_ID, MODEL, WIDTH, HEIGHT, PARAMS[]\n")
        f.write("# Number of cameras: {}\n".format(len(cameras)))
        for cam in sorted(cameras, key=lambda c: c["camera_id"]):
            f.write("{} {} {} {} {}\n".format(cam["camera_id"], cam["model"], cam["width"], cam["height"],
                                              " ".join("{:.12g}".format(v) for v in cam["params"])))
    with (out_dir / "images.txt").open("w", encoding="utf-8") as f:
        f.write("# Image list with two lines of data per image:\n"
                "#   IMAGE_ID, QW, QX, QY, QZ, TX, TY, TZ, CAMERA_ID, NAME\n#   POINTS2D[] as (X, Y, POINT3D_ID)\n")
        f.write("# Number of images: {}, mean observations per image: {:.3f}\n".format(len(images), mean_obs))
        for im in sorted(images, key=lambda m: m["image_id"]):
            f.write("{image_id} {qw:.12g} {qx:.12g} {qy:.12g} {qz:.12g} {tx:.12g} {ty:.12g} {tz:.12g} "
                    "{camera_id} {name}\n".format(**im))
            f.write((im.get("points2d_line", "") or "") + "\n")
    with (out_dir / "points3D.txt").open("w", encoding="utf-8") as f:
        f.write("# 3D point list with one line of data per point:\n"
                "#   POINT3D_ID, X, Y, Z, R, G, B, ERROR, TRACK[] as (IMAGE_ID, POINT2D_IDX)\n")
        f.write("# Number of points: {}, mean track length: {:.6f}\n".format(len(points), mean_track))
        for pt in points:
            line = "{id} {x:.12g} {y:.12g} {z:.12g} {r} {g} {b} {err:.6g}".format(**pt)
            tokens = pt.get("track_tokens", []) or []
            if tokens:
                line += " " + " ".join(str(t) for t in tokens)
            f.write(line + "\n")


def _indent(elem, level: int = 0) -> None:
    pad = "\n" + "  " * level
    if len(elem):
        if not (elem.text or "").strip():
            elem.text = pad + "  "
        for child in elem:
            _indent(child, level + 1)
        if not (elem[-1].tail or "").strip():
            elem[-1].tail = pad
    if level and not (elem.tail or "").strip():
        elem.tail = pad


def write_metashape_perspective_xml(path, cameras: Iterable[Mapping[str, object]],
                                    images: Sequence[Mapping[str, object]]) -> None:
    """Metashape document: one frame sensor per distinct (size, fx, fy), one camera per image whose <transform> is
    the camera-to-world matrix in the OpenCV convention (CameraFormatConverter:938-1035)."""
    path = pathlib.Path(path)
    path.parent.mkdir(parents=True, exist_ok=True)
    by_id = {int(c["camera_id"]): c for c in cameras}
    sensors: Dict[tuple, Dict[str, object]] = {}
    sensor_of_cam: Dict[int, int] = {}
    for im in images:
        cam = by_id[int(im["camera_id"])]
        if str(cam["model"]).upper() != "PINHOLE":
            raise ValueError("unsupported COLMAP camera model: {}".format(cam["model"]))
        fx, fy = float(cam["params"][0]), float(cam["params"][1])
        key = (int(cam["width"]), int(cam["height"]), round(fx, 9), round(fy, 9))
        if key not in sensors:
            sensors[key] = {"id": len(sensors), "w": key[0], "h": key[1], "f": 0.5 * (fx + fy)}
        sensor_of_cam[int(cam["camera_id"])] = sensors[key]["id"]
    doc = ET.Element("document", {"version": "1.2.0"})
    chunk = ET.SubElement(doc, "chunk", {"label": "unknown", "enabled": "true"})
    sensors_node = ET.SubElement(chunk, "sensors", {"next_id": str(len(sensors))})
    for s in sorted(sensors.values(), key=lambda d: d["id"]):
        node = ET.SubElement(sensors_node, "sensor", {"id": str(s["id"]), "label": "virtual_fisheyelike", "type": "frame"})
        ET.SubElement(node, "resolution", {"width": str(s["w"]), "height": str(s["h"])})
        ET.SubElement(node, "property", {"name": "layer_index", "value": "0"})
        ET.SubElement(node, "data_type").text = "uint8"
        calib = ET.SubElement(node, "calibration", {"type": "frame", "class": "initial"})
        ET.SubElement(calib, "resolution", {"width": str(s["w"]), "height": str(s["h"])})
        ET.SubElement(calib, "f").text = "{:.15g}".format(s["f"])
        ET.SubElement(node, "black_level").text = "0 0 0"
        ET.SubElement(node, "sensitivity").text = "1 1 1"
    comps = ET.SubElement(chunk, "components", {"next_id": "1", "active_id": "0"})
    ET.SubElement(ET.SubElement(comps, "component", {"id": "0", "label": "Component 1"}), "partition")
    cams_node = ET.SubElement(chunk, "cameras", {"next_id": str(len(images)), "next_group_id": "0"})
    for idx, im in enumerate(images):
        r_wc = quat_wxyz_to_rotmat(im["qw"], im["qx"], im["qy"], im["qz"])
        r_cw = transpose3(r_wc)
        centre = matvec3(r_cw, [-im["tx"], -im["ty"], -im["tz"]])
        node = ET.SubElement(cams_node, "camera", {"id": str(idx), "sensor_id": str(sensor_of_cam[int(im["camera_id"])]),
                                                   "component_id": "0", "label": pathlib.Path(im["name"]).stem})
        ET.SubElement(node, "transform").text = " ".join("{:.15g}".format(float(v)) for row in rigid(r_cw, centre) for v in row)
    _indent(doc)
    with path.open("wb") as f:
        f.write(b"<?xml version='1.0' encoding='UTF-8'?>\n")
        f.write(ET.tostring(doc, encoding="utf-8"))
        f.write(b"\n")
