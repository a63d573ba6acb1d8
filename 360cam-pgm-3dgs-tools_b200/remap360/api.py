"""Tensor-level host API over the C ABI: torch provides device memory and streams, the
library provides the kernels.

    views = [PerspectiveView(yaw, pitch, hfov, vfov), ...]
    out = remap_erp(frames_u8_BHWC_cuda, views, (1600, 1600), interp="cubic")   # [B, V, h, w, C]

The two calls stand in for the reference's two remap back ends: one ffmpeg ``v360`` process per
(frame, view) (cli_tools/gs360_360PerspCut.py:286-349, :569-590) and NumPy map build +
``cv2.remap`` + mask fill (cli_tools/gs360_DualFisheyeDistortionCalibration.py:1759-1823,
:2001-2014)."""

from __future__ import annotations

import ctypes
from dataclasses import dataclass
from typing import Dict, Optional, Sequence, Tuple

import torch

from . import _lib
from ._lib import FisheyeCalib, Images, Options, View


@dataclass(frozen=True)
class PerspectiveView:
    """One rectilinear view: the yaw/pitch/roll/h_fov/v_fov options of the v360 filter string
    (gs360_360PerspCut.py:310-314) or one entry of build_sfm10_specs (DF:1258-1307)."""
    yaw_deg: float
    pitch_deg: float
    hfov_deg: float
    vfov_deg: float
    roll_deg: float = 0.0
    src_slot: int = 0          # dual fisheye: 0 = X lens image, 1 = Y lens image
    view_id: str = ""


@dataclass(frozen=True)
class FisheyeCalibration:
    """Fields of the reference's SensorCalibration (DF:67-85) used by the projection, plus the
    usable lens FOV (--lens-fov-deg, DF:395-402)."""
    width: float
    height: float
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
    lens_fov_deg: float = 190.0


_TORCH_TO_R360 = {torch.uint8: _lib.DTYPE_U8, torch.uint16: _lib.DTYPE_U16,
                  torch.float16: _lib.DTYPE_F16, torch.float32: _lib.DTYPE_F32}
_R360_TO_TORCH = {v: k for k, v in _TORCH_TO_R360.items()}


def _dtype_code(dt) -> int:
    if dt not in _TORCH_TO_R360:
        raise TypeError("unsupported tensor dtype %s (uint8, uint16, float16, float32)" % (dt,))
    return _TORCH_TO_R360[dt]


def _describe(t: torch.Tensor, what: str) -> Images:
    """[N, H, W, C] tensor (channel-contiguous, pixel-contiguous rows) -> r360_images."""
    if not t.is_cuda:
        raise ValueError("%s must be a CUDA tensor (there is no CPU path)" % what)
    if t.dim() != 4:
        raise ValueError("%s must be [N, H, W, C], got %s" % (what, tuple(t.shape)))
    n, h, w, c = t.shape
    es = t.element_size()
    if t.stride(3) != 1 or t.stride(2) != c:
        raise ValueError("%s must be interleaved HWC with contiguous pixels" % what)
    if n > 1 and t.stride(0) < t.stride(1) * h:
        raise ValueError("%s: overlapping images" % what)
    return Images(data=t.data_ptr(), width=w, height=h, channels=c, dtype=_dtype_code(t.dtype),
                  pitch_bytes=t.stride(1) * es, image_stride_bytes=t.stride(0) * es, count=n, reserved=0)


def _views_array(views: Sequence[PerspectiveView]):
    arr = (View * len(views))()
    for k, v in enumerate(views):
        arr[k] = View(float(v.yaw_deg), float(v.pitch_deg), float(v.roll_deg), float(v.hfov_deg),
                      float(v.vfov_deg), int(v.src_slot), 0)
    return arr


def _calib_array(calibs: Sequence[FisheyeCalibration]):
    arr = (FisheyeCalib * len(calibs))()
    for k, c in enumerate(calibs):
        arr[k] = FisheyeCalib(*(float(getattr(c, name)) for name, _ in FisheyeCalib._fields_))
    return arr


def _options(interp: str, convention: str = "halfpixel", path: str = "auto", fill_invalid: bool = True,
             border_value: float = 0.0, out_dtype: Optional[torch.dtype] = None) -> Options:
    opt = _lib.default_options()
    try:
        opt.interp = _lib.INTERP[interp]
        opt.convention = _lib.CONVENTION[convention]
        opt.path = _lib.PATH[path]
    except KeyError as exc:
        raise ValueError("unknown option value %s" % exc) from None
    opt.fill_invalid = 1 if fill_invalid else 0
    opt.border_value = float(border_value)
    opt.out_dtype = -1 if out_dtype is None else _dtype_code(out_dtype)
    return opt


def _stream_handle(stream: Optional[torch.cuda.Stream], device) -> int:
    return (stream or torch.cuda.current_stream(device)).cuda_stream


def remap_erp(frames: torch.Tensor, views: Sequence[PerspectiveView], size: Tuple[int, int], *,
              interp: str = "cubic", convention: str = "halfpixel", out: Optional[torch.Tensor] = None,
              out_dtype: Optional[torch.dtype] = None, path: str = "auto",
              stream: Optional[torch.cuda.Stream] = None) -> torch.Tensor:
    """ERP frames [B, H, W, C] -> views [B, V, h, w, C] (size = (w, h))."""
    lib = _lib.load()
    if frames.dim() == 3:
        frames = frames.unsqueeze(0)
    b, _, _, c = frames.shape
    w, h = int(size[0]), int(size[1])
    dt = out_dtype or frames.dtype
    if out is None:
        out = torch.empty((b, len(views), h, w, c), dtype=dt, device=frames.device)
    elif tuple(out.shape) != (b, len(views), h, w, c) or out.dtype != dt:
        raise ValueError("out must be %s %s" % ((b, len(views), h, w, c), dt))
    src = _describe(frames, "frames")
    dst = _describe(out.view(b * len(views), h, w, c), "out")
    opt = _options(interp, convention, path, out_dtype=None if dt == frames.dtype else dt)
    with torch.cuda.device(frames.device):
        _lib.check(lib.r360_remap_erp(ctypes.byref(src), ctypes.byref(dst), _views_array(views), len(views),
                                      ctypes.byref(opt), _stream_handle(stream, frames.device)))
    return out


def remap_fisheye(images: torch.Tensor, calibs: Sequence[FisheyeCalibration],
                  views: Sequence[PerspectiveView], size: Tuple[int, int], *, interp: str = "cubic",
                  border_value: float = 0.0, fill_invalid: bool = True, out: Optional[torch.Tensor] = None,
                  out_dtype: Optional[torch.dtype] = None, path: str = "auto",
                  stream: Optional[torch.cuda.Stream] = None) -> torch.Tensor:
    """Fisheye groups [G, L, H, W, C] (L lens images per group) -> views [G, V, h, w, C].
    View yaw is relative to the lens named by ``src_slot`` (DF:1883)."""
    lib = _lib.load()
    if images.dim() != 5:
        raise ValueError("images must be [G, L, H, W, C]")
    g, nl, hh, ww, c = images.shape
    if nl != len(calibs):
        raise ValueError("one calibration per lens image is required")
    w, h = int(size[0]), int(size[1])
    dt = out_dtype or images.dtype
    if out is None:
        out = torch.empty((g, len(views), h, w, c), dtype=dt, device=images.device)
    elif tuple(out.shape) != (g, len(views), h, w, c) or out.dtype != dt:
        raise ValueError("out must be %s %s" % ((g, len(views), h, w, c), dt))
    if not images.is_contiguous():
        raise ValueError("images must be contiguous")
    src = _describe(images.view(g * nl, hh, ww, c), "images")
    dst = _describe(out.view(g * len(views), h, w, c), "out")
    opt = _options(interp, "halfpixel", path, fill_invalid, border_value,
                   None if dt == images.dtype else dt)
    with torch.cuda.device(images.device):
        _lib.check(lib.r360_remap_fisheye(ctypes.byref(src), ctypes.byref(dst), _calib_array(calibs), nl,
                                          _views_array(views), len(views), ctypes.byref(opt),
                                          _stream_handle(stream, images.device)))
    return out


def sample_coordinates(views: Sequence[PerspectiveView], size: Tuple[int, int], *,
                       erp_size: Optional[Tuple[int, int]] = None,
                       calibs: Optional[Sequence[FisheyeCalibration]] = None,
                       convention: str = "halfpixel", path: str = "auto", device="cuda",
                       stream: Optional[torch.cuda.Stream] = None) -> Dict[str, torch.Tensor]:
    """The source coordinates the kernels sample at (test/debug): float32 maps as cv2.remap would
    receive them, their float64 pre-images, and (fisheye) the validity mask; each [V, h, w]."""
    lib = _lib.load()
    w, h = int(size[0]), int(size[1])
    device = torch.device(device)
    n = len(views)
    res = {"x32": torch.empty((n, h, w), dtype=torch.float32, device=device),
           "y32": torch.empty((n, h, w), dtype=torch.float32, device=device),
           "x64": torch.empty((n, h, w), dtype=torch.float64, device=device),
           "y64": torch.empty((n, h, w), dtype=torch.float64, device=device)}
    valid_ptr = None
    if calibs is not None:
        res["valid"] = torch.empty((n, h, w), dtype=torch.uint8, device=device)
        valid_ptr = res["valid"].data_ptr()
        cal, nl, sw, sh = _calib_array(calibs), len(calibs), 0, 0
    else:
        if erp_size is None:
            raise ValueError("erp_size or calibs is required")
        cal, nl, sw, sh = None, 1, int(erp_size[0]), int(erp_size[1])
    opt = _options("cubic", convention, path)
    with torch.cuda.device(device):
        _lib.check(lib.r360_coords(sw, sh, cal, nl, _views_array(views), n, w, h, ctypes.byref(opt),
                                   res["x32"].data_ptr(), res["y32"].data_ptr(), res["x64"].data_ptr(),
                                   res["y64"].data_ptr(), valid_ptr, _stream_handle(stream, device)))
    return res
