"""Tensor-level host API over the C ABI: torch provides device memory and streams, the
library provides the kernels.

    views = [PerspectiveView(yaw, pitch, hfov, vfov), ...]
    out = remap_erp(frames_u8_BHWC_cuda, views, (1600, 1600), interp="cubic")   # [B, V, h, w, C]

The two calls stand in for the reference's two remap back ends: one ffmpeg ``v360`` process per
(frame, view) (cli_tools/gs360_360PerspCut.py:286-349, :569-590) and NumPy map build +
``cv2.remap`` + mask fill (cli_tools/gs360_DualFisheyeDistortionCalibration.py:1759-1823,
:2001-2014).

``path="auto"`` (default) uses the tiled kernels through a cached *plan* -- the counterpart of
the reference building its maps once and applying them to every frame (DF:1857-1907 /
:1996-2014); ``path="direct"`` runs the per-pixel float64 path with no plan."""

from __future__ import annotations

import ctypes
import threading
from collections import OrderedDict
from dataclasses import dataclass
from typing import Dict, Optional, Sequence, Tuple

import torch

from . import _lib
from ._lib import FisheyeCalib, Images, Options, Undistort, View


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
    projection: str = "rectilinear"   # "fisheye": equidistant fisheye output, hfov/vfov as v360's
                                      # h_fov/v_fov of `output=fisheye` (PC:375-379, preset fisheyeXY)


def fisheye_fov_from_dfov(d_fov_deg: float, width: int, height: int) -> Tuple[float, float]:
    """(h_fov, v_fov) that v360 derives from ``d_fov`` for ``output=fisheye`` [upstream FFmpeg
    vf_v360.c fov_from_dfov, unverified here: no ffmpeg in this image]: the diagonal of the frame
    spans d_fov on an equidistant projection, so each side spans its share of the half-diagonal."""
    d = 0.5 * (float(width) ** 2 + float(height) ** 2) ** 0.5
    return d / float(width) * float(d_fov_deg), d / float(height) * float(d_fov_deg)


@dataclass(frozen=True)
class UndistortItem:
    """One output of the fisheye -> undistorted-fisheye remap: the lens image it reads and the zoom
    (RemapCache.undistort_zoom, DF:1136-1145)."""
    zoom: float
    src_slot: int = 0
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
    model: str = "equisolid"         # or "equidistant" (v360 input=fisheye, V2F:466-473)


_TORCH_TO_R360 = {torch.uint8: _lib.DTYPE_U8, torch.uint16: _lib.DTYPE_U16,
                  torch.float16: _lib.DTYPE_F16, torch.float32: _lib.DTYPE_F32}


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
        if v.projection not in _lib.OUT_PROJECTION:
            raise ValueError("unknown view projection %r" % (v.projection,))
        arr[k] = View(float(v.yaw_deg), float(v.pitch_deg), float(v.roll_deg), float(v.hfov_deg),
                      float(v.vfov_deg), int(v.src_slot), _lib.OUT_PROJECTION[v.projection])
    return arr


def _is_undistort(views) -> bool:
    return len(views) > 0 and isinstance(views[0], UndistortItem)


def _undistort_array(items: Sequence[UndistortItem]):
    arr = (Undistort * len(items))()
    for k, it in enumerate(items):
        arr[k] = Undistort(float(it.zoom), int(it.src_slot), 0)
    return arr


def _view_key(v):
    if isinstance(v, UndistortItem):
        return ("undistort", v.zoom, v.src_slot)
    return (v.yaw_deg, v.pitch_deg, v.roll_deg, v.hfov_deg, v.vfov_deg, v.src_slot, v.projection)


def _calib_array(calibs: Sequence[FisheyeCalibration]):
    arr = (FisheyeCalib * len(calibs))()
    for k, c in enumerate(calibs):
        if c.model not in _lib.LENS_MODEL:
            raise ValueError("unknown lens model %r" % (c.model,))
        arr[k] = FisheyeCalib(*(float(getattr(c, name)) for name, _ in FisheyeCalib._fields_[:14]),
                              _lib.LENS_MODEL[c.model], 0)
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


def alloc_views(n_groups: int, n_views: int, h: int, w: int, c: int, dtype, device) -> torch.Tensor:
    """Output tensor [G, V, h, w, C] whose rows start on 16-byte boundaries: when ``w * C * itemsize`` is
    not a multiple of 16 (e.g. the dual-fisheye default 1750 px x 3 bytes) the rows are padded and a
    strided view is returned, so that the kernels can use 16-byte vector stores.  ``.contiguous()``
    gives the packed image."""
    es = torch.empty((), dtype=dtype).element_size()
    row_bytes = w * c * es
    if row_bytes % 16 == 0:
        return torch.empty((n_groups, n_views, h, w, c), dtype=dtype, device=device)
    pitch = (row_bytes + 15) // 16 * 16 // es            # elements per padded row
    flat = torch.empty((n_groups, n_views, h, pitch), dtype=dtype, device=device)
    return flat.as_strided((n_groups, n_views, h, w, c), (n_views * h * pitch, h * pitch, pitch, c, 1))


def _as_image_batch(t: torch.Tensor) -> torch.Tensor:
    """[G, V, h, w, C] -> [G * V, h, w, C] without copying, also for row-padded tensors."""
    g, v, h, w, c = t.shape
    if t.is_contiguous():
        return t.view(g * v, h, w, c)
    if t.stride(0) != v * t.stride(1):
        raise ValueError("out: views of a group must be equally spaced and groups back to back")
    return t.as_strided((g * v, h, w, c), (t.stride(1), t.stride(2), t.stride(3), t.stride(4)))


# ---- plans --------------------------------------------------------------------------------------

class Plan:
    """A built tile plan (r360_plan) plus the device workspace it lives in."""

    def __init__(self, handle: int, workspace: torch.Tensor, tiles_per_view: int, n_fallback: int, n_views: int,
                 n_map_tiles: int = 0):
        self.handle, self.workspace = handle, workspace
        self.tiles_per_view, self.n_fallback, self.n_views = tiles_per_view, n_fallback, n_views
        self.n_map_tiles = n_map_tiles         # tiles that carry a per-pixel map (pole neighbourhoods)

    @property
    def fallback_fraction(self) -> float:
        return self.n_fallback / float(self.tiles_per_view * self.n_views)

    def __del__(self):
        try:
            if self.handle:
                _lib.load().r360_plan_destroy(self.handle)
                self.handle = 0
        except Exception:
            pass


_PLAN_CACHE: "OrderedDict[tuple, Plan]" = OrderedDict()
_PLAN_CACHE_SIZE = 32
_PLAN_LOCK = threading.Lock()       # the job runners call in from several host threads


def _layout_key(im: Images, aligned: bool):
    return (im.width, im.height, im.channels, im.dtype, im.pitch_bytes,
            im.image_stride_bytes if im.count > 1 else 0, aligned)


def get_plan(src: Images, dst: Images, views: Sequence[PerspectiveView], opt: Options, device,
             calibs: Optional[Sequence[FisheyeCalibration]] = None,
             stream: Optional[torch.cuda.Stream] = None) -> Plan:
    """Build (or fetch from the cache) the plan for this layout / view set / option set."""
    lib = _lib.load()
    device = torch.device(device)
    key = (device.index, _layout_key(src, src.data % 16 == 0 if src.data else True),
           _layout_key(dst, dst.data % 16 == 0 if dst.data else True),
           tuple(_view_key(v) for v in views),
           None if calibs is None else tuple(tuple(getattr(c, n) for n, _ in FisheyeCalib._fields_[:15]) for c in calibs),
           (opt.interp, opt.convention, opt.fill_invalid, opt.border_value, opt.out_dtype))
    with _PLAN_LOCK:
        plan = _PLAN_CACHE.get(key)
        if plan is not None:
            _PLAN_CACHE.move_to_end(key)
            return plan
    nbytes = lib.r360_plan_workspace_bytes(len(views), dst.width, dst.height)
    workspace = torch.empty(int(nbytes), dtype=torch.uint8, device=device)
    handle = ctypes.c_void_p()
    popt = Options.from_buffer_copy(opt)
    popt.path = _lib.PATH["tiled"]
    with torch.cuda.device(device):
        if calibs is None:
            rc = lib.r360_plan_create_erp(ctypes.byref(src), ctypes.byref(dst), _views_array(views), len(views),
                                          ctypes.byref(popt), workspace.data_ptr(), nbytes,
                                          _stream_handle(stream, device), ctypes.byref(handle))
        elif _is_undistort(views):
            rc = lib.r360_plan_create_undistort(ctypes.byref(src), ctypes.byref(dst), _calib_array(calibs),
                                                len(calibs), _undistort_array(views), len(views),
                                                ctypes.byref(popt), workspace.data_ptr(), nbytes,
                                                _stream_handle(stream, device), ctypes.byref(handle))
        else:
            rc = lib.r360_plan_create_fisheye(ctypes.byref(src), ctypes.byref(dst), _calib_array(calibs), len(calibs),
                                              _views_array(views), len(views), ctypes.byref(popt),
                                              workspace.data_ptr(), nbytes, _stream_handle(stream, device),
                                              ctypes.byref(handle))
    _lib.check(rc)
    tiles, nfb = ctypes.c_int32(), ctypes.c_int32()
    _lib.check(lib.r360_plan_info(handle, ctypes.byref(tiles), ctypes.byref(nfb)))
    nmap = ctypes.c_int32()
    _lib.check(lib.r360_plan_info_maps(handle, ctypes.byref(nmap), None))
    plan = Plan(handle.value, workspace, tiles.value, nfb.value, len(views), nmap.value)
    with _PLAN_LOCK:
        # two threads may have built the same plan at once: both are valid, the later one stays cached
        _PLAN_CACHE[key] = plan
        while len(_PLAN_CACHE) > _PLAN_CACHE_SIZE:
            _PLAN_CACHE.popitem(last=False)
    return plan


def clear_plan_cache() -> None:
    with _PLAN_LOCK:
        _PLAN_CACHE.clear()


def _run(src: Images, dst: Images, views, opt: Options, path: str, device, stream, calibs=None) -> None:
    lib = _lib.load()
    if path == "direct":
        with torch.cuda.device(device):
            if calibs is None:
                rc = lib.r360_remap_erp(ctypes.byref(src), ctypes.byref(dst), _views_array(views), len(views),
                                        ctypes.byref(opt), _stream_handle(stream, device))
            elif _is_undistort(views):
                rc = lib.r360_remap_undistort(ctypes.byref(src), ctypes.byref(dst), _calib_array(calibs),
                                              len(calibs), _undistort_array(views), len(views), ctypes.byref(opt),
                                              _stream_handle(stream, device))
            else:
                rc = lib.r360_remap_fisheye(ctypes.byref(src), ctypes.byref(dst), _calib_array(calibs), len(calibs),
                                            _views_array(views), len(views), ctypes.byref(opt),
                                            _stream_handle(stream, device))
        _lib.check(rc)
        return
    if len(views) == 0:
        raise _lib.Remap360Error(-1, "invalid argument (no views)")
    plan = get_plan(src, dst, views, opt, device, calibs, stream)
    if stream is not None:
        plan.workspace.record_stream(stream)       # the plan may be evicted from the cache while in use
    with torch.cuda.device(device):
        _lib.check(lib.r360_remap_planned(plan.handle, ctypes.byref(src), ctypes.byref(dst),
                                          _stream_handle(stream, device)))


def remap_erp(frames: torch.Tensor, views: Sequence[PerspectiveView], size: Tuple[int, int], *,
              interp: str = "cubic", convention: str = "halfpixel", out: Optional[torch.Tensor] = None,
              out_dtype: Optional[torch.dtype] = None, path: str = "auto",
              stream: Optional[torch.cuda.Stream] = None) -> torch.Tensor:
    """ERP frames [B, H, W, C] -> views [B, V, h, w, C] (size = (w, h))."""
    if frames.dim() == 3:
        frames = frames.unsqueeze(0)
    b, _, _, c = frames.shape
    w, h = int(size[0]), int(size[1])
    dt = out_dtype or frames.dtype
    if out is None:
        out = alloc_views(b, len(views), h, w, c, dt, frames.device)
    elif tuple(out.shape) != (b, len(views), h, w, c) or out.dtype != dt:
        raise ValueError("out must be %s %s" % ((b, len(views), h, w, c), dt))
    if len(views) == 0:
        raise _lib.Remap360Error(-1, "invalid argument (no views)")
    src = _describe(frames, "frames")
    dst = _describe(_as_image_batch(out), "out")
    opt = _options(interp, convention, path, out_dtype=None if dt == frames.dtype else dt)
    _run(src, dst, views, opt, path, frames.device, stream)
    return out


def remap_fisheye(images: torch.Tensor, calibs: Sequence[FisheyeCalibration],
                  views: Sequence[PerspectiveView], size: Tuple[int, int], *, interp: str = "cubic",
                  border_value: float = 0.0, fill_invalid: bool = True, out: Optional[torch.Tensor] = None,
                  out_dtype: Optional[torch.dtype] = None, path: str = "auto",
                  stream: Optional[torch.cuda.Stream] = None) -> torch.Tensor:
    """Fisheye groups [G, L, H, W, C] (L lens images per group) -> views [G, V, h, w, C].
    View yaw is relative to the lens named by ``src_slot`` (DF:1883)."""
    if images.dim() != 5:
        raise ValueError("images must be [G, L, H, W, C]")
    g, nl, hh, ww, c = images.shape
    if nl != len(calibs):
        raise ValueError("one calibration per lens image is required")
    w, h = int(size[0]), int(size[1])
    dt = out_dtype or images.dtype
    if out is None:
        out = alloc_views(g, len(views), h, w, c, dt, images.device)
    elif tuple(out.shape) != (g, len(views), h, w, c) or out.dtype != dt:
        raise ValueError("out must be %s %s" % ((g, len(views), h, w, c), dt))
    if not images.is_contiguous():
        raise ValueError("images must be contiguous")
    if len(views) == 0:
        raise _lib.Remap360Error(-1, "invalid argument (no views)")
    src = _describe(images.view(g * nl, hh, ww, c), "images")
    dst = _describe(_as_image_batch(out), "out")
    opt = _options(interp, "halfpixel", path, fill_invalid, border_value, None if dt == images.dtype else dt)
    _run(src, dst, views, opt, path, images.device, stream, calibs)
    return out


def undistort_fisheye(images: torch.Tensor, calibs: Sequence[FisheyeCalibration],
                      items: Sequence[UndistortItem], size: Optional[Tuple[int, int]] = None, *,
                      interp: str = "cubic", border_value: float = 0.0, fill_invalid: bool = True,
                      out: Optional[torch.Tensor] = None, out_dtype: Optional[torch.dtype] = None,
                      path: str = "auto", stream: Optional[torch.cuda.Stream] = None) -> torch.Tensor:
    """Fisheye groups [G, L, H, W, C] -> undistorted fisheye images [G, N, h, w, C], one per item
    (DF:1120-1217: build_remap_cache + cv2.remap + mask fill).  ``size`` defaults to the source size,
    which is what the reference writes."""
    if images.dim() != 5:
        raise ValueError("images must be [G, L, H, W, C]")
    g, nl, hh, ww, c = images.shape
    if nl != len(calibs):
        raise ValueError("one calibration per lens image is required")
    if len(items) == 0:
        raise _lib.Remap360Error(-1, "invalid argument (no items)")
    w, h = (ww, hh) if size is None else (int(size[0]), int(size[1]))
    dt = out_dtype or images.dtype
    if out is None:
        out = alloc_views(g, len(items), h, w, c, dt, images.device)
    elif tuple(out.shape) != (g, len(items), h, w, c) or out.dtype != dt:
        raise ValueError("out must be %s %s" % ((g, len(items), h, w, c), dt))
    if not images.is_contiguous():
        raise ValueError("images must be contiguous")
    src = _describe(images.view(g * nl, hh, ww, c), "images")
    dst = _describe(_as_image_batch(out), "out")
    opt = _options(interp, "halfpixel", path, fill_invalid, border_value, None if dt == images.dtype else dt)
    _run(src, dst, list(items), opt, path, images.device, stream, calibs)
    return out


def sample_coordinates(views: Sequence[PerspectiveView], size: Tuple[int, int], *,
                       erp_size: Optional[Tuple[int, int]] = None,
                       calibs: Optional[Sequence[FisheyeCalibration]] = None,
                       convention: str = "halfpixel", path: str = "auto", device="cuda",
                       stream: Optional[torch.cuda.Stream] = None) -> Dict[str, torch.Tensor]:
    """The source coordinates the kernels sample at (test/debug): float32 maps as cv2.remap would
    receive them, their float64 pre-images, and (fisheye) the validity mask; each [V, h, w].
    With a tiled path the plan is the one an 8-bit 3-channel contiguous source would get."""
    lib = _lib.load()
    w, h = int(size[0]), int(size[1])
    device = torch.device(device)
    if device.index is None:
        device = torch.device("cuda", torch.cuda.current_device())
    n = len(views)
    res = {"x32": torch.empty((n, h, w), dtype=torch.float32, device=device),
           "y32": torch.empty((n, h, w), dtype=torch.float32, device=device),
           "x64": torch.empty((n, h, w), dtype=torch.float64, device=device),
           "y64": torch.empty((n, h, w), dtype=torch.float64, device=device)}
    valid_ptr = None
    if calibs is not None:
        res["valid"] = torch.empty((n, h, w), dtype=torch.uint8, device=device)
        valid_ptr = res["valid"].data_ptr()
        cal, nl, sw, sh = _calib_array(calibs), len(calibs), int(calibs[0].width), int(calibs[0].height)
    else:
        if erp_size is None:
            raise ValueError("erp_size or calibs is required")
        cal, nl, sw, sh = None, 1, int(erp_size[0]), int(erp_size[1])
    opt = _options("cubic", convention, path)
    if path == "direct" and _is_undistort(views):
        with torch.cuda.device(device):
            _lib.check(lib.r360_coords_undistort(cal, nl, _undistort_array(views), n, w, h,
                                                 res["x32"].data_ptr(), res["y32"].data_ptr(),
                                                 res["x64"].data_ptr(), res["y64"].data_ptr(), valid_ptr,
                                                 _stream_handle(stream, device)))
        return res
    if path == "direct":
        with torch.cuda.device(device):
            _lib.check(lib.r360_coords(sw if calibs is None else 0, sh if calibs is None else 0, cal, nl,
                                       _views_array(views), n, w, h, ctypes.byref(opt),
                                       res["x32"].data_ptr(), res["y32"].data_ptr(), res["x64"].data_ptr(),
                                       res["y64"].data_ptr(), valid_ptr, _stream_handle(stream, device)))
        return res
    src = Images(data=None, width=sw, height=sh, channels=3, dtype=_lib.DTYPE_U8, pitch_bytes=sw * 3,
                 image_stride_bytes=sw * sh * 3, count=nl, reserved=0)
    dst = Images(data=None, width=w, height=h, channels=3, dtype=_lib.DTYPE_U8, pitch_bytes=w * 3,
                 image_stride_bytes=w * h * 3, count=n, reserved=0)
    plan = get_plan(src, dst, views, opt, device, calibs, stream)
    with torch.cuda.device(device):
        _lib.check(lib.r360_plan_coords(plan.handle, res["x32"].data_ptr(), res["y32"].data_ptr(),
                                        res["x64"].data_ptr(), res["y64"].data_ptr(), valid_ptr,
                                        _stream_handle(stream, device)))
    res["fallback_fraction"] = torch.tensor(plan.fallback_fraction)
    return res
