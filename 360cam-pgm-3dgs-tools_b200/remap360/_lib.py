"""ctypes binding of libremap360.so (the C ABI in include/remap360.h).

There is no CPU fallback: if the shared library is missing or fails to load, importing
anything that needs it raises."""

from __future__ import annotations

import ctypes
import os
import pathlib
from ctypes import POINTER, Structure, c_char_p, c_double, c_float, c_int, c_int32, c_int64, c_uint8, c_void_p

LIB_PATH = pathlib.Path(__file__).resolve().parent / "libremap360.so"

R360_OK = 0
DTYPE_U8, DTYPE_U16, DTYPE_F16, DTYPE_F32 = 0, 1, 2, 3
INTERP = {"nearest": 0, "linear": 1, "bilinear": 1, "cubic": 2, "bicubic": 2, "lanczos4": 3}
CONVENTION = {"halfpixel": 0, "v360": 1}
PATH = {"auto": 0, "direct": 1, "tiled": 2}
OUT_PROJECTION = {"rectilinear": 0, "fisheye": 1}
LENS_MODEL = {"equisolid": 0, "equidistant": 1}
TRC = {"bt709": 0, "smpte170m": 0, "iec61966-2-1": 1, "srgb": 1, "linear": 2}
MAX_LENSES = 4


class Images(Structure):
    _fields_ = [("data", c_void_p), ("width", c_int32), ("height", c_int32), ("channels", c_int32),
                ("dtype", c_int32), ("pitch_bytes", c_int64), ("image_stride_bytes", c_int64),
                ("count", c_int32), ("reserved", c_int32)]


class View(Structure):
    _fields_ = [("yaw_deg", c_double), ("pitch_deg", c_double), ("roll_deg", c_double),
                ("hfov_deg", c_double), ("vfov_deg", c_double), ("src_slot", c_int32), ("projection", c_int32)]


class FisheyeCalib(Structure):
    _fields_ = [(n, c_double) for n in ("width", "height", "f", "cx", "cy", "k1", "k2", "k3", "k4",
                                        "p1", "p2", "b1", "b2", "lens_fov_deg")] + \
               [("model", c_int32), ("reserved", c_int32)]


class Undistort(Structure):
    _fields_ = [("zoom", c_double), ("src_slot", c_int32), ("reserved", c_int32)]


class Lut3D(Structure):
    _fields_ = [("table_device", c_void_p), ("size", c_int32), ("reserved", c_int32),
                ("domain_min", c_float * 3), ("domain_max", c_float * 3)]


class ColorConvert(Structure):
    _fields_ = [("in_trc", c_int32), ("out_trc", c_int32), ("matrix", c_float * 9), ("reserved", c_int32)]


class Options(Structure):
    _fields_ = [("interp", c_int32), ("convention", c_int32), ("path", c_int32), ("fill_invalid", c_int32),
                ("border_value", c_double), ("out_dtype", c_int32), ("reserved", c_int32)]


class Remap360Error(RuntimeError):
    def __init__(self, code: int, text: str):
        super().__init__("remap360 error %d: %s" % (code, text))
        self.code = code


_lib = None

# every symbol include/remap360.h declares (tests check the .so exports exactly these)
EXPORTS = ("r360_abi_version", "r360_error_string", "r360_last_cuda_error", "r360_default_options",
           "r360_device_info", "r360_remap_erp", "r360_remap_fisheye", "r360_coords", "r360_launch_count",
           "r360_plan_workspace_bytes", "r360_plan_create_erp", "r360_plan_create_fisheye", "r360_plan_info",
           "r360_plan_info_maps", "r360_remap_planned", "r360_plan_coords", "r360_plan_destroy",
           "r360_remap_undistort", "r360_coords_undistort", "r360_plan_create_undistort", "r360_apply_lut",
           "r360_convert_color")


def load() -> ctypes.CDLL:
    global _lib
    if _lib is not None:
        return _lib
    lib_path = pathlib.Path(os.environ.get("R360_LIBRARY") or LIB_PATH)      # override: kernel experiments only
    if not lib_path.exists():
        raise ImportError(
            "%s is missing: build it with `python 360cam-pgm-3dgs-tools_b200/build.py` "
            "(there is no CPU fallback for the remap kernels)" % lib_path)
    lib = ctypes.CDLL(str(lib_path))
    lib.r360_abi_version.restype = c_int
    lib.r360_error_string.restype = c_char_p
    lib.r360_error_string.argtypes = [c_int]
    lib.r360_last_cuda_error.restype = c_char_p
    lib.r360_default_options.argtypes = [POINTER(Options)]
    lib.r360_default_options.restype = None
    lib.r360_device_info.argtypes = [POINTER(c_int), POINTER(c_int), POINTER(c_int)]
    lib.r360_remap_erp.argtypes = [POINTER(Images), POINTER(Images), POINTER(View), c_int32,
                                   POINTER(Options), c_void_p]
    lib.r360_remap_fisheye.argtypes = [POINTER(Images), POINTER(Images), POINTER(FisheyeCalib), c_int32,
                                       POINTER(View), c_int32, POINTER(Options), c_void_p]
    lib.r360_coords.argtypes = [c_int32, c_int32, POINTER(FisheyeCalib), c_int32, POINTER(View), c_int32,
                                c_int32, c_int32, POINTER(Options), c_void_p, c_void_p, c_void_p, c_void_p,
                                c_void_p, c_void_p]
    lib.r360_launch_count.restype = c_int64
    lib.r360_plan_workspace_bytes.restype = ctypes.c_size_t
    lib.r360_plan_workspace_bytes.argtypes = [c_int32, c_int32, c_int32]
    lib.r360_plan_create_erp.argtypes = [POINTER(Images), POINTER(Images), POINTER(View), c_int32, POINTER(Options),
                                         c_void_p, ctypes.c_size_t, c_void_p, POINTER(c_void_p)]
    lib.r360_plan_create_fisheye.argtypes = [POINTER(Images), POINTER(Images), POINTER(FisheyeCalib), c_int32,
                                             POINTER(View), c_int32, POINTER(Options), c_void_p, ctypes.c_size_t,
                                             c_void_p, POINTER(c_void_p)]
    lib.r360_remap_undistort.argtypes = [POINTER(Images), POINTER(Images), POINTER(FisheyeCalib), c_int32,
                                         POINTER(Undistort), c_int32, POINTER(Options), c_void_p]
    lib.r360_coords_undistort.argtypes = [POINTER(FisheyeCalib), c_int32, POINTER(Undistort), c_int32, c_int32,
                                          c_int32, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p]
    lib.r360_plan_create_undistort.argtypes = [POINTER(Images), POINTER(Images), POINTER(FisheyeCalib), c_int32,
                                               POINTER(Undistort), c_int32, POINTER(Options), c_void_p,
                                               ctypes.c_size_t, c_void_p, POINTER(c_void_p)]
    lib.r360_apply_lut.argtypes = [POINTER(Images), POINTER(Images), POINTER(Lut3D), c_int32, c_int32, c_void_p]
    lib.r360_convert_color.argtypes = [POINTER(Images), POINTER(Images), POINTER(ColorConvert), c_int32, c_void_p]
    lib.r360_plan_info.argtypes = [c_void_p, POINTER(c_int32), POINTER(c_int32)]
    lib.r360_plan_info_maps.argtypes = [c_void_p, POINTER(c_int32), POINTER(c_int32)]
    lib.r360_remap_planned.argtypes = [c_void_p, POINTER(Images), POINTER(Images), c_void_p]
    lib.r360_plan_coords.argtypes = [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p]
    lib.r360_plan_destroy.argtypes = [c_void_p]
    lib.r360_plan_destroy.restype = None
    lib.r360_debug_weight_tables.argtypes = [c_void_p, c_void_p]
    lib.r360_debug_weight_tables_lanczos4.argtypes = [c_void_p, c_void_p]
    if lib.r360_abi_version() != 2:
        raise ImportError("libremap360.so has ABI version %d, expected 2" % lib.r360_abi_version())
    _lib = lib
    return lib


def check(code: int) -> None:
    if code != R360_OK:
        lib = load()
        text = lib.r360_error_string(code).decode()
        if code == -3 or code == -4:
            detail = lib.r360_last_cuda_error().decode()
            if detail:
                text += " (" + detail + ")"
        raise Remap360Error(code, text)


def default_options() -> Options:
    opt = Options()
    load().r360_default_options(ctypes.byref(opt))
    return opt


def launch_count() -> int:
    return int(load().r360_launch_count())
