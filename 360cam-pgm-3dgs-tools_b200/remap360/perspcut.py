"""Drop-in for the module surface of cli_tools/gs360_360PerspCut.py, backed by the CUDA remap.

What stays identical to the reference (checked against fixtures recorded from it,
tests/golden/perspcut_views.json): the argument parser (gs360_360PerspCut.py:417-532), the view
sets / names / FOV numbers / log lines of ``build_view_jobs`` (:593-980) -- including the
ffmpeg-style argv kept in ``BuildResult.jobs`` because the GUI edits those lists in place
(gs360_GUI.py:19092-19147) -- and the ``run_one`` / ``stop_event`` / ``parse_jobs`` contract
(:535-590).  What changes: ``run_one`` does not spawn ffmpeg; it reads the job back out of the
argv and runs it on the GPU (remap360.executor).

Presets: default 12 mm, fisheyelike 17 mm, full360coverage 14 mm, 2views 6 mm / 3600 px,
evenMinus30 / evenPlus30, fisheyeXY (gs360_360PerspCut.py:616-680).
"""

from __future__ import annotations

import argparse
import math
import os
import pathlib
import re
import shlex
import signal
import sys
import threading
from concurrent.futures import ThreadPoolExecutor, as_completed
from dataclasses import dataclass
from typing import Dict, Iterable, List, Optional, Sequence, Set, Tuple

EXTS = {".tif", ".tiff", ".jpg", ".jpeg", ".png"}
PROGRESS_INTERVAL = 5
PRESET_NAMES = ("default", "fisheyelike", "full360coverage", "2views", "evenMinus30", "evenPlus30", "fisheyeXY")


class StoreWithFlag(argparse.Action):
    """Stores the value and marks ``<dest>_explicit`` (gs360_360PerspCut.py:24-29): presets only
    override --size / --hfov / --focal-mm when the user did not pass them."""

    def __call__(self, parser, namespace, values, option_string=None):
        setattr(namespace, self.dest, values)
        setattr(namespace, self.dest + "_explicit", True)


@dataclass
class ViewSpec:
    """gs360_360PerspCut.py:32-45."""
    source_path: pathlib.Path
    output_name: str
    view_id: str
    yaw_deg: float
    pitch_deg: float
    hfov_deg: float
    vfov_deg: float
    width: int
    height: int
    projection: str = "perspective"


@dataclass
class BuildResult:
    """gs360_360PerspCut.py:48-65."""
    jobs: List[Tuple[List[str], str, str]]
    view_specs: List[ViewSpec]
    focal_used_mm: float
    focal_35mm_equiv: Optional[float]
    hfov_deg: float
    vfov_deg: float
    preview_views_line: str
    sensor_line: str
    realityscan_line: str
    metashape_line: str

    @property
    def total(self) -> int:
        return len(self.jobs)


# ---- small pure helpers (gs360_360PerspCut.py:67-180) -----------------------------------------

def update_progress(label: str, completed: int, total: int, last_pct: int) -> int:
    if total <= 0:
        return last_pct
    pct = int(completed * 100 / total)
    if last_pct < 0 or pct >= 100 or pct - last_pct >= PROGRESS_INTERVAL:
        sys.stdout.write("%s... %3d%% (%d/%d)\r" % (label, pct, completed, total))
        sys.stdout.flush()
        return pct
    return last_pct


def fov_from_focal_mm(f_mm: float, sensor_w_mm: float) -> float:
    return math.degrees(2.0 * math.atan(sensor_w_mm / (2.0 * f_mm)))


def focal_from_hfov_deg(hfov_deg: float, sensor_w_mm: float) -> float:
    return sensor_w_mm / (2.0 * math.tan(math.radians(hfov_deg) / 2.0))


def v_fov_from_hfov(hfov_deg: float, w: int, h: int) -> float:
    half = math.radians(hfov_deg) / 2.0
    return math.degrees(2.0 * math.atan(math.tan(half) * (h / float(w))))


def letter_tag(idx: int) -> str:
    return chr(ord("A") + idx) if idx < 26 else "%02d" % (idx + 1)


def letter_to_index1(s: str) -> int:
    s = s.strip()
    if not s:
        raise ValueError("empty key")
    if s.isdigit():
        return int(s)
    first = s.upper()[0]
    if "A" <= first <= "Z":
        return ord(first) - ord("A") + 1
    raise ValueError("invalid key: " + s)


def normalize_angle_deg(a: float) -> float:
    a = ((a + 180.0) % 360.0) - 180.0
    return 180.0 if abs(a + 180.0) < 1e-6 else a


def clamp(v: float, lo: float, hi: float) -> float:
    return max(lo, min(hi, v))


def map_interp_for_v360(name: str) -> str:
    return {"bicubic": "cubic", "bilinear": "linear", "lanczos": "lanczos"}.get((name or "").lower(), "cubic")


def _sensor_tokens(text: str) -> List[str]:
    norm = text.lower().replace("×", "x").replace(",", " ").strip()
    if "x" in norm:
        return [t.strip() for t in norm.split("x") if t.strip()]
    return [t for t in norm.split() if t]


def parse_sensor(s: str) -> float:
    norm = s.lower().replace("×", "x").replace(",", " ").strip()
    return float(norm.split("x")[0].strip() if "x" in norm else norm.split()[0])


def parse_sensor_dimensions(s: str) -> Tuple[float, ...]:
    dims = []
    for tok in _sensor_tokens(s):
        try:
            dims.append(float(tok))
        except ValueError:
            pass
    return tuple(dims)


def extra_suffix(delta_pitch: float, default_deg: float = 30.0) -> str:
    head = "_U" if delta_pitch > 0 else "_D"
    mag = abs(delta_pitch)
    if abs(mag - default_deg) < 1e-6:
        return head
    if float(mag).is_integer():
        return "%s%d" % (head, int(round(mag)))
    return "%s%g" % (head, mag)


def bit_depth_from_pixel_format_tag(tag: int) -> int:
    """Nominal bit depth from the codec pixel-format tag OpenCV reports (CAP_PROP_CODEC_PIXEL_FORMAT =
    avcodec_pix_fmt_to_codec_tag): libavcodec's raw tags spell high-bit-depth planar YUV as 'Y' '3' <chroma> <bits>
    (yuv420p10le = Y3 0x0b 0x0a), semi-planar ones as P010 / P012 / P016 / P210 / P216 / P410 / P416, packed ones as
    v210 / Y210 / Y410 / Y216 / Y416, 16-bit RGB as 'b48r' / '0RGB'-style 48 / 64 tags.  Anything above 8 bits counts
    as 10, like the reference (gs360_360PerspCut.py:133-146); unknown tags are 8."""
    raw = int(tag) & 0xFFFFFFFF
    b = raw.to_bytes(4, "little")
    if b[:2] == b"Y3" and b[3] in (9, 10, 12, 14, 16):
        return 10
    if b[::-1][:2] == b"Y3" and b[0] in (9, 10, 12, 14, 16):       # big-endian variants are stored reversed
        return 10
    text = b.decode("latin-1")
    if text in ("P010", "P012", "P016", "P210", "P212", "P216", "P410", "P412", "P416", "v210", "Y210", "Y212", "Y216",
                "Y410", "Y412", "Y416", "b48r", "b64a", "RBA@", "BRA@"):
        return 10
    if b[:2] == b"G3" and b[3] in (9, 10, 12, 14, 16):             # planar GBR
        return 10
    return 8


def detect_input_bit_depth(in_path: pathlib.Path) -> int:
    """Nominal bit depth of a video (reference: ffprobe's bits_per_raw_sample / pix_fmt, gs360_360PerspCut.py:111-149).
    There is no ffprobe here; the container is opened with OpenCV and the STREAM's pixel format is read from its
    metadata (the decoded frames are always 8-bit BGR and say nothing)."""
    try:
        import cv2
        cap = cv2.VideoCapture(str(in_path))
        if cap.isOpened():
            tag = int(cap.get(cv2.CAP_PROP_CODEC_PIXEL_FORMAT))
            cap.release()
            return bit_depth_from_pixel_format_tag(tag)
    except Exception:
        pass
    return 8


# ---- --addcam / --delcam / --setcam (gs360_360PerspCut.py:183-283) ---------------------------------

_UD_TOKEN = re.compile(r"^([UD])\s*([+-]?\d+(?:\.\d+)?)?$")


def parse_addcam_spec(spec: str, default_deg: float) -> Dict[int, List[float]]:
    result: Dict[int, List[float]] = {}
    for raw in (spec or "").split(","):
        tok = raw.strip()
        if not tok:
            continue
        if ":" not in tok and "=" not in tok:
            result.setdefault(letter_to_index1(tok), []).extend([+default_deg, -default_deg])
            continue
        key, val = re.split(r"[:=]", tok, maxsplit=1)
        m = _UD_TOKEN.match(val.strip().upper())
        if not m:
            raise ValueError("invalid --addcam token: " + tok)
        deg = float(m.group(2)) if m.group(2) else default_deg
        result.setdefault(letter_to_index1(key), []).append(deg if m.group(1) == "U" else -deg)
    return result


def parse_delcam_spec(spec: str) -> Set[int]:
    return {letter_to_index1(t.strip()) for t in (spec or "").split(",") if t.strip()}


def parse_setcam_spec(spec: str, default_deg: float):
    """Returns (abs_map, delta_map, extra_abs_map, extra_delta_map); the extra maps are keyed by
    (index, suffix) and address add-cam views such as ``A_U``."""
    abs_map: Dict[int, float] = {}
    delta_map: Dict[int, float] = {}
    extra_abs: Dict[Tuple[int, str], float] = {}
    extra_delta: Dict[Tuple[int, str], float] = {}
    for raw in (spec or "").split(","):
        tok = raw.strip()
        if not tok:
            continue
        if ":" not in tok and "=" not in tok:
            raise ValueError("invalid --setcam token: " + tok)
        key_txt, val_txt = re.split(r"[:=]", tok, maxsplit=1)
        key_txt = key_txt.strip()
        suffix = None
        if "_" in key_txt:
            base, tail = key_txt.split("_", 1)
            suffix = "_" + tail.strip()
            key_txt = base
        idx = letter_to_index1(key_txt)
        key = (idx, suffix) if suffix else idx
        t_abs, t_delta = (extra_abs, extra_delta) if suffix else (abs_map, delta_map)
        val = val_txt.strip()
        if re.match(r"^[+|-]\s*\d+(?:\.\d+)?$", val):
            t_delta[key] = float(val.replace(" ", ""))
            continue
        up = re.match(r"^[Uu]\s*(\d+(?:\.\d+)?)?$", val)
        down = re.match(r"^[Dd]\s*(\d+(?:\.\d+)?)?$", val)
        if up:
            t_abs[key] = +(float(up.group(1)) if up.group(1) else default_deg)
        elif down:
            t_abs[key] = -(float(down.group(1)) if down.group(1) else default_deg)
        else:
            try:
                t_abs[key] = float(val.replace(" ", ""))
            except Exception as exc:
                raise ValueError("invalid --setcam token: " + tok) from exc
    return abs_map, delta_map, extra_abs, extra_delta


# ---- ffmpeg-style argv (kept for GUI compatibility; gs360_360PerspCut.py:286-414) -----------------

def _job_argv(ffmpeg: str, inp: pathlib.Path, out: pathlib.Path, v360_filter: str, ext: str, *,
              video_mode: bool, fps: Optional[float], keep_rec709: bool, bit_depth: int,
              jpeg_quality_95: bool, start_time: Optional[float], end_time: Optional[float]) -> List[str]:
    ext = ext.lower()
    is_jpeg = ext in (".jpg", ".jpeg")
    chain: List[str] = []
    if video_mode:
        if fps is None or fps <= 0:
            raise ValueError("fps must be specified and > 0 when processing a video input")
        chain.append("fps=%s" % fps)
        cs = "colorspace=iall=bt709:all=smpte170m" + ("" if keep_rec709 else ":trc=iec61966-2-1")
        chain.append(cs + (":range=jpeg:format=yuv444p" if is_jpeg else ":format=yuv444p"))
    chain.append(v360_filter)
    argv = [ffmpeg, "-hide_banner", "-loglevel", "error", "-y"]
    if video_mode and start_time is not None:
        argv += ["-ss", "%s" % max(0.0, float(start_time))]
    argv += ["-i", str(inp)]
    if video_mode and end_time is not None:
        argv += ["-to", "%s" % max(0.0, float(end_time))]
    argv += ["-vf", ",".join(chain), "-threads", "1"]
    argv += ["-vsync", "vfr", "-start_number", "0"] if video_mode else ["-frames:v", "1"]
    if is_jpeg:
        q = "2" if jpeg_quality_95 else "1"
        argv += ["-c:v", "mjpeg", "-q:v", q, "-qmin", q, "-qmax", q, "-pix_fmt", "yuvj444p", "-huffman", "optimal"]
        if video_mode:
            argv += ["-colorspace", "smpte170m", "-color_primaries", "smpte170m", "-color_trc", "smpte170m"]
    elif video_mode and ext in (".png", ".tif", ".tiff"):
        argv += ["-pix_fmt", "rgb48le" if bit_depth > 8 else "rgb24"]
    argv.append(str(out))
    return argv


def build_ffmpeg_cmd(ffmpeg: str, inp: pathlib.Path, out: pathlib.Path, w: int, h: int, yaw: float,
                     pitch: float, hfov: float, vfov: float, interp_v360: str, ext: str, *,
                     video_mode: bool = False, fps: Optional[float] = None, keep_rec709: bool = False,
                     bit_depth: int = 8, jpeg_quality_95: bool = False,
                     start_time: Optional[float] = None, end_time: Optional[float] = None) -> List[str]:
    flt = ("v360=input=equirect:output=rectilinear:w=%s:h=%s:yaw=%s:pitch=%s:roll=0:h_fov=%s:v_fov=%s:interp=%s"
           % (w, h, yaw, pitch, hfov, vfov, interp_v360))
    return _job_argv(ffmpeg, inp, out, flt, ext, video_mode=video_mode, fps=fps, keep_rec709=keep_rec709,
                     bit_depth=bit_depth, jpeg_quality_95=jpeg_quality_95, start_time=start_time, end_time=end_time)


def build_ffmpeg_equisolid_cmd(ffmpeg: str, inp: pathlib.Path, out: pathlib.Path, w: int, h: int, yaw: float,
                               pitch: float, fov_deg: float, interp_v360: str, ext: str, *,
                               video_mode: bool = False, fps: Optional[float] = None,
                               keep_rec709: bool = False, bit_depth: int = 8, jpeg_quality_95: bool = False,
                               start_time: Optional[float] = None, end_time: Optional[float] = None) -> List[str]:
    flt = ("v360=input=equirect:output=fisheye:w=%s:h=%s:yaw=%s:pitch=%s:roll=0:d_fov=%s:interp=%s"
           % (w, h, yaw, pitch, fov_deg, interp_v360))
    return _job_argv(ffmpeg, inp, out, flt, ext, video_mode=video_mode, fps=fps, keep_rec709=keep_rec709,
                     bit_depth=bit_depth, jpeg_quality_95=jpeg_quality_95, start_time=start_time, end_time=end_time)


# ---- argument parser (gs360_360PerspCut.py:417-532) ---------------------------------------------------

def create_arg_parser() -> argparse.ArgumentParser:
    ap = argparse.ArgumentParser(
        description=("Batch convert equirectangular images into perspective views on the GPU "
                     "(drop-in for the ffmpeg/v360 cutter), including optional virtual camera "
                     "add/delete/set operations."),
        formatter_class=argparse.ArgumentDefaultsHelpFormatter,
        epilog=("Notes: presets can be overridden with --focal-mm / --size / --sensor-mm. "
                "Priority: --hfov overrides --focal-mm. "
                "Use --setcam to specify absolute or relative pitch values per camera."))
    ap.add_argument("-i", "--in", dest="input_dir", required=True,
                    help="Input folder (equirectangular images) or a video file of equirectangular frames")
    ap.add_argument("-o", "--out", dest="out_dir", default=None,
                    help="Output folder. Defaults to <input>/_geometry if omitted")
    ap.add_argument("--preset", choices=list(PRESET_NAMES), default="default",
                    help=("default=8-view baseline / fisheyelike=10-view mix (17mm) / "
                          "full360coverage=12-view wide cover (14mm) / 2views=front/back (6mm, 3600px) / "
                          "evenMinus30, evenPlus30=even slots pitched / fisheyeXY=fisheye X/Y pair (3600px FOV180)"))
    ap.add_argument("--count", type=int, default=8, help="Horizontal division count (4=90deg, 8=45deg)")
    ap.add_argument("--addcam", default="",
                    help="Add virtual cameras, e.g. 'B' (+/-default pitch), 'B:U', 'D:D20', 'F:U15' (comma separated)")
    ap.add_argument("--addcam-deg", type=float, default=30.0,
                    help="Default magnitude in degrees when 'U/D' in --addcam/--setcam omit a value")
    ap.add_argument("--add-top", action="store_true", help="Include cube-map style top view (pitch +90 deg)")
    ap.add_argument("--add-bottom", action="store_true", help="Include cube-map style bottom view (pitch -90 deg)")
    ap.add_argument("--add-topdown", action="store_true", dest="add_topdown", help=argparse.SUPPRESS)
    ap.add_argument("--delcam", default="", help="Remove baseline cameras by letter, e.g. 'B,D'")
    ap.add_argument("--setcam", default="",
                    help="Override/adjust baseline pitch. Absolute: 'A=30','A=U','A=D20'. Relative: 'A:+10','B:-5'.")
    ap.add_argument("--size", type=int, default=1600, action=StoreWithFlag, help="Square output size per view")
    ap.add_argument("--ext", default="jpg", help="Output extension (jpg=high quality)")
    ap.add_argument("--jpeg-quality-95", action="store_true",
                    help="With --ext jpg, encode at approximately 95%% JPEG quality instead of maximum.")
    ap.add_argument("-f", "--fps", type=float, default=None, help="Frame extraction rate (fps) for video input")
    ap.add_argument("--start", type=float, default=None, help="Optional start time in seconds (video input)")
    ap.add_argument("--end", type=float, default=None, help="Optional end time in seconds (video input)")
    ap.add_argument("--keep-rec709", action="store_true",
                    help="Keep Rec.709 transfer characteristics for video inputs (default: convert to sRGB)")
    ap.add_argument("--hfov", type=float, default=None, action=StoreWithFlag,
                    help="Horizontal FOV in degrees (overrides focal length)")
    ap.add_argument("--focal-mm", type=float, default=12.0, action=StoreWithFlag,
                    help="Focal length in millimetres when --hfov is not set")
    ap.add_argument("--sensor-mm", default="36 36", help="Sensor width/height in millimetres, e.g. '36 36' or '36x24'")
    ap.add_argument("-j", "--jobs", default="auto", help="Concurrent jobs (number or 'auto'=cores/2)")
    ap.add_argument("--print-cmd", choices=["once", "none", "all"], default="once",
                    help="How many job command lines to print: once/none/all")
    ap.add_argument("--ffmpeg", default="ffmpeg",
                    help="Kept for compatibility: first word of the job argv (no ffmpeg process is started)")
    ap.add_argument("--dry-run", action="store_true", help="Print all commands without executing them")
    return ap


# ---- execution state and cancellation (gs360_360PerspCut.py:535-590) ------------------------------------

stop_event = threading.Event()
_signal_hits = 0


def on_signal(sig, frame):
    global _signal_hits
    _signal_hits += 1
    if not stop_event.is_set():
        print("\n[INFO] Cancel requested. Stopping new jobs...", file=sys.stderr)
        stop_event.set()
    if _signal_hits >= 2:
        print("[INFO] Force exiting", file=sys.stderr)


def install_signal_handlers() -> None:
    try:
        signal.signal(signal.SIGINT, on_signal)
        signal.signal(signal.SIGTERM, on_signal)
    except Exception:
        pass


def parse_jobs(s: str) -> int:
    if str(s).lower() == "auto":
        return max(1, (os.cpu_count() or 1) // 2)
    return max(1, int(s))


def run_one(cmd: List[str]) -> Tuple[int, str]:
    """Execute one (source, view) job given as ffmpeg-style argv.  Returns (rc, stderr_text);
    (130, "") when cancelled.  Thread-safe."""
    if stop_event.is_set():
        return 130, ""
    from . import executor
    return executor.run_job_argv(cmd, stop_event)


# ---- the planner (gs360_360PerspCut.py:593-980) --------------------------------------------------------

# preset -> (forced count, focal mm, size, slots deleted, slots that get +/- add-cams)
_PRESET_TABLE = {
    "fisheyelike": (10, 17.0, None, "CDHI", "AF"),
    "full360coverage": (8, 14.0, None, "BDFH", "BDFH"),
    "2views": (None, 6.0, 3600, "BCDFGH", ""),
}


class _PitchRules:
    """--setcam lookups (gs360_360PerspCut.py:802-819)."""

    def __init__(self, spec: str, default_deg: float):
        self.abs, self.delta, self.extra_abs, self.extra_delta = parse_setcam_spec(spec, default_deg)

    def apply(self, idx: int, pitch: float, suffix: Optional[str] = None) -> float:
        if suffix:
            key = (idx, suffix)
            if key in self.extra_abs:
                pitch = float(self.extra_abs[key])
            elif idx in self.abs:
                pitch = float(self.abs[idx])
            if key in self.extra_delta:
                pitch += float(self.extra_delta[key])
            elif idx in self.delta:
                pitch += float(self.delta[idx])
            return pitch
        if idx in self.abs:
            pitch = float(self.abs[idx])
        if idx in self.delta:
            pitch += float(self.delta[idx])
        return pitch


def _ensure_pair(slot: List[float], deg: float) -> None:
    if not any(abs(v - deg) < 1e-6 for v in slot):
        slot.append(deg)
    if not any(abs(v + deg) < 1e-6 for v in slot):
        slot.append(-deg)


def _view_id_from_name(out_name: str, stem: str, video_mode: bool) -> str:
    out_stem = pathlib.Path(out_name).stem
    if video_mode and out_stem.startswith(stem + "_%07d_"):
        return out_stem[len(stem) + 6:]
    if out_stem.startswith(stem + "_"):
        return out_stem[len(stem) + 1:]
    return out_stem


def build_view_jobs(args, files: List[pathlib.Path], out_dir: pathlib.Path) -> BuildResult:
    """Pure planning: view set, names, FOVs, job argv and the four log lines.  Mutates ``args``
    the way the reference does (count, size, focal_mm, add_top, add_bottom)."""
    size_explicit = getattr(args, "size_explicit", False)
    hfov_explicit = getattr(args, "hfov_explicit", False)
    focal_explicit = getattr(args, "focal_mm_explicit", False)
    video_mode = bool(getattr(args, "input_is_video", False))
    job_kwargs = dict(video_mode=video_mode, fps=getattr(args, "fps", None),
                      keep_rec709=bool(getattr(args, "keep_rec709", False)),
                      bit_depth=int(getattr(args, "video_bit_depth", 8)),
                      jpeg_quality_95=args.jpeg_quality_95,
                      start_time=getattr(args, "start", None), end_time=getattr(args, "end", None))

    add_top = bool(getattr(args, "add_top", False))
    add_bottom = bool(getattr(args, "add_bottom", False))
    if getattr(args, "add_topdown", False):
        add_top = add_bottom = True
    args.add_top, args.add_bottom = add_top, add_bottom

    preset = args.preset
    fisheye_xy = preset == "fisheyeXY"
    even_pitch = {"evenMinus30": -30.0, "evenPlus30": +30.0}.get(preset)
    forced_count, preset_focal, preset_size, preset_del, preset_add = _PRESET_TABLE.get(
        preset, (None, None, None, "", ""))
    if forced_count is not None:
        args.count = forced_count
    elif fisheye_xy:
        if args.count != 8:
            print("[INFO] preset 'fisheyeXY' forces count=8")
        args.count = 8
    if preset_size is not None and not size_explicit:
        args.size = preset_size
    if preset_focal is not None and not hfov_explicit and not focal_explicit:
        args.focal_mm = preset_focal

    add_map = parse_addcam_spec(args.addcam, args.addcam_deg)
    del_set = parse_delcam_spec(args.delcam)
    user_add = bool(str(getattr(args, "addcam", "")).strip()) or bool(getattr(args, "addcam_explicit", False))
    user_del = bool(str(getattr(args, "delcam", "")).strip()) or bool(getattr(args, "delcam_explicit", False))
    if preset == "2views":
        del_set.update(letter_to_index1(ch) for ch in preset_del)
    elif preset in _PRESET_TABLE:
        if not user_del:
            del_set.update(letter_to_index1(ch) for ch in preset_del)
        if not user_add:
            for ch in preset_add:
                _ensure_pair(add_map.setdefault(letter_to_index1(ch), []), float(args.addcam_deg))
    rules = _PitchRules(args.setcam, args.addcam_deg)

    # ---- optics -----------------------------------------------------------------------------
    sensor_w = parse_sensor(args.sensor_mm)
    dims = parse_sensor_dimensions(args.sensor_mm)
    sensor_long = max(dims) if dims else sensor_w
    sensor_h = float(dims[1]) if len(dims) >= 2 else sensor_w
    if sensor_h <= 0:
        sensor_h = None
    if args.hfov is not None:
        hfov = float(args.hfov)
        focal = focal_from_hfov_deg(hfov, sensor_w)
    else:
        focal = float(args.focal_mm)
        hfov = fov_from_focal_mm(focal, sensor_w)
    focal_35 = None
    if sensor_long and sensor_long > 0 and abs(sensor_long - 36.0) > 1e-6:
        focal_35 = focal * (36.0 / sensor_long)
    w = h = int(args.size)
    if sensor_h and focal > 1e-6:
        vfov = max(1.0, min(179.9, math.degrees(2.0 * math.atan(sensor_h / (2.0 * focal)))))
    else:
        vfov = v_fov_from_hfov(hfov, w, h)
    fisheye_size = (w if size_explicit else 3600) if fisheye_xy else w
    fisheye_fov = (hfov if hfov_explicit else 180.0) if fisheye_xy else hfov

    count = int(args.count)
    if count <= 0:
        print("[ERR] --count must be >= 1", file=sys.stderr)
        sys.exit(1)
    yaw_step = 360.0 / count
    ext_dot = "." + args.ext.lower().lstrip(".")
    interp = map_interp_for_v360("bicubic")        # the reference hard-codes cubic (:730)

    jobs: List[Tuple[List[str], str, str]] = []
    specs: List[ViewSpec] = []
    taken: Set[str] = set()

    for img in files:
        stem = img.stem

        def out_path_for(view_id: str) -> pathlib.Path:
            pattern = "%s_%%07d_%s%s" if video_mode else "%s_%s%s"
            return out_dir / (pattern % (stem, view_id, ext_dot))

        def emit(view_id: str, yaw: float, pitch: float, *, fisheye: bool = False) -> None:
            path = out_path_for(view_id)
            if path.name in taken:
                return
            if fisheye:
                argv = build_ffmpeg_equisolid_cmd(args.ffmpeg, img, path, fisheye_size, fisheye_size, yaw, pitch,
                                                  fisheye_fov, interp, ext_dot, **job_kwargs)
                dims_fov = (fisheye_size, fisheye_size, fisheye_fov, fisheye_fov, "equisolid")
            else:
                argv = build_ffmpeg_cmd(args.ffmpeg, img, path, w, h, yaw, pitch, hfov, vfov, interp, ext_dot,
                                        **job_kwargs)
                dims_fov = (w, h, hfov, vfov, "perspective")
            jobs.append((argv, img.name, path.name))
            taken.add(path.name)
            specs.append(ViewSpec(source_path=img, output_name=path.name,
                                  view_id=_view_id_from_name(path.name, stem, video_mode),
                                  yaw_deg=yaw, pitch_deg=pitch, hfov_deg=dims_fov[2], vfov_deg=dims_fov[3],
                                  width=dims_fov[0], height=dims_fov[1], projection=dims_fov[4]))

        xy_pending: List[Tuple[str, float, float]] = []
        for slot in range(count):
            if stop_event.is_set():
                break
            idx1 = slot + 1
            tag = letter_tag(slot)
            yaw = normalize_angle_deg(slot * yaw_step)
            pitch = 0.0
            if idx1 % 2 == 0 and not fisheye_xy and even_pitch is not None:
                pitch += even_pitch
            pitch = clamp(rules.apply(idx1, pitch), -90.0, 90.0)
            if fisheye_xy:
                if idx1 in (1, 5):
                    xy_pending.append(("X" if idx1 == 1 else "Y", yaw, pitch))
                continue
            if idx1 not in del_set:
                emit(tag, yaw, pitch)
            for delta in add_map.get(idx1, ()):
                suffix = extra_suffix(delta, args.addcam_deg)
                extra_pitch = rules.apply(idx1, clamp(pitch + delta, -90.0, 90.0), suffix=suffix)
                emit(tag + suffix, yaw, extra_pitch)
        for tag, yaw, pitch in xy_pending:
            emit(tag, yaw, pitch, fisheye=True)

        pole_pitches = ([90.0] if add_top else []) + ([-90.0] if add_bottom else [])
        for n, pole_pitch in enumerate(pole_pitches):
            tag = letter_tag(count + n)
            emit(tag, 0.0, rules.apply(letter_to_index1(tag), pole_pitch))

    # ---- log lines (gs360_360PerspCut.py:919-967) ---------------------------------------------
    views_line = sensor_line = rs_line = ms_line = ""
    if jobs:
        first_src = jobs[0][1]
        ref_stem = pathlib.Path(first_src).stem
        seen: List[str] = []
        for _, src_name, dst_name in jobs:
            if src_name != first_src:
                break
            vid = _view_id_from_name(dst_name, ref_stem, video_mode)
            if vid and vid not in seen:
                seen.append(vid)
        if seen:
            views_line = "[INFO] View summary (%s): %d view%s - %s" % (
                first_src, len(seen), "s" if len(seen) != 1 else "", ", ".join(seen))
            if fisheye_xy:
                views_line += " | fisheye_fov=%.1fdeg | size=%dx%d" % (fisheye_fov, fisheye_size, fisheye_size)
            else:
                sensor_line = "[INFO] Sensor=%s mm | size=%dx%d" % (args.sensor_mm, w, h)
                focal_txt = "focal length=  %.3f mm" % focal
                if focal_35 is not None:
                    focal_txt += " (35mm eq=  %.3f mm)" % focal_35
                rs_line = "[INFO] For RealityScan: " + focal_txt
                if w > 0 and sensor_w / float(w) > 0:
                    px_mm = sensor_w / float(w)
                    ms_line = "[INFO] For Metashape: Precalibrated f=  %.5f  | pixel_size=  %.4f mm" % (
                        focal / px_mm, px_mm)

    return BuildResult(jobs=jobs, view_specs=specs, focal_used_mm=focal, focal_35mm_equiv=focal_35,
                       hfov_deg=hfov, vfov_deg=vfov, preview_views_line=views_line, sensor_line=sensor_line,
                       realityscan_line=rs_line, metashape_line=ms_line)


# ---- CLI (gs360_360PerspCut.py:983-1087) -----------------------------------------------------------------

def main(argv: Optional[Sequence[str]] = None) -> None:
    install_signal_handlers()
    args = create_arg_parser().parse_args(argv)
    for name in ("size", "hfov", "focal_mm"):
        setattr(args, name + "_explicit", getattr(args, name + "_explicit", False))

    input_path = pathlib.Path(args.input_dir).expanduser().resolve()
    if input_path.is_dir():
        args.input_is_video, args.video_bit_depth = False, 8
        out_dir = pathlib.Path(args.out_dir).resolve() if args.out_dir else input_path / "_geometry"
        out_dir.mkdir(parents=True, exist_ok=True)
        files = [p for p in sorted(input_path.iterdir()) if p.is_file() and p.suffix.lower() in EXTS]
        if not files:
            print("[WARN] No target images found (tif/jpg/png)", file=sys.stderr)
            sys.exit(0)
    elif input_path.is_file():
        args.input_is_video = True
        if args.fps is None or args.fps <= 0:
            print("[ERR] -f/--fps must be specified for video inputs", file=sys.stderr)
            sys.exit(1)
        out_dir = pathlib.Path(args.out_dir).resolve() if args.out_dir else (
            input_path.parent / (input_path.stem + "_geometry"))
        out_dir.mkdir(parents=True, exist_ok=True)
        args.video_bit_depth = detect_input_bit_depth(input_path)
        files = [input_path]
    else:
        print("[ERR] Input path not found:", input_path, file=sys.stderr)
        sys.exit(1)

    result = build_view_jobs(args, files, out_dir)
    total = result.total
    if args.dry_run:
        for cmd, _, _ in result.jobs:
            print("$ " + " ".join(shlex.quote(c) for c in cmd))
        print("\n[DRY] Exiting without execution (total %d commands)" % total)
        return
    if args.print_cmd == "all":
        for cmd, _, _ in result.jobs:
            print("$ " + " ".join(shlex.quote(c) for c in cmd))
    elif args.print_cmd == "once" and result.jobs:
        print("$ " + " ".join(shlex.quote(c) for c in result.jobs[0][0]))

    workers = parse_jobs(args.jobs)
    print("[INFO] parallel jobs: %d / total: %d" % (workers, total))
    if result.preview_views_line:
        print(result.preview_views_line)
        for line in (result.sensor_line, result.realityscan_line, result.metashape_line):
            if line:
                print(line)

    from . import multigpu
    ok = fail = done = 0
    last_pct = -1
    # Jobs of one source share one decode and one upload: the executor groups them, but results
    # are still reported per job like the reference's one-process-per-job pool.  With several GPUs visible the
    # sources (and a video's frame ranges) are dealt out to one worker process per device (remap360/multigpu.py).
    for (cmd, src, dst), (rc, err) in multigpu.run_jobs(result.jobs, stop_event, workers):
        done += 1
        if rc == 0:
            ok += 1
            last_pct = update_progress("Progress", done, total, last_pct)
            continue
        fail += 1
        if stop_event.is_set():
            continue
        last_pct = update_progress("Progress", done, total, last_pct)
        sys.stdout.write("\n")
        sys.stdout.flush()
        print("[%d/%d] %s %s" % (done, total, dst, "canceled" if rc == 130 else "failed"), file=sys.stderr)
        if err.strip():
            print(err.strip(), file=sys.stderr)
    if total and last_pct >= 0:
        sys.stdout.write("\n")
        sys.stdout.flush()
    if stop_event.is_set():
        print("[STOPPED] Interrupted: success=%d, failed=%d, total=%d" % (ok, fail, total))
        sys.exit(130)
    print("[OK] Completed: success=%d, failed=%d, total=%d" % (ok, fail, total))


if __name__ == "__main__":
    main()
