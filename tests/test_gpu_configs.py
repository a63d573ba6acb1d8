"""GPU parity at the sizes and view sets BASELINE.json names (configs 1, 3, 4; config 2 and 5 live in
tests/test_gpu_parity.py), through the planner's own presets and real FOVs.

  cfg 1  3840 x 1920 uint8 panorama, `default` preset (8 views, 112.62 deg = 12 mm) -> 1600^2, linear + cubic
  cfg 3  7680 x 3840 uint8, `fisheyelike` preset (10 views, 93.27 deg = 17 mm) -> 1600^2, frames in blocks
  cfg 4  7680 x 3840 uint16 (16-bit linear content), bicubic -> uint16 and -> float16, the 12 `full360coverage`
         views plus seam (180, +-179.9), poles (+-90) and +-60 degree views; seam columns and pole rows asserted
         on their own

Oracle: the float64 maps of oracle/geometry.py fed to the real cv2.remap (sampler.sample_cv2, the library call the
reference makes, DF:2001-2014) where cv2 has the type, the NumPy model (sampler.sample) on a subset of rows for
the float16 output cv2 cannot produce.  Bars: <= 1 LSB on >= 99.9 % of the pixels (BASELINE.json north_star).
"""

import numpy as np
import pytest

torch = pytest.importorskip("torch")

pytestmark = pytest.mark.gpu

from oracle import geometry as geo  # noqa: E402
from oracle import sampler  # noqa: E402

OK_FRACTION = 0.999


@pytest.fixture(scope="module")
def r360():
    import remap360
    if not torch.cuda.is_available():
        pytest.fail("-m gpu tests need a CUDA device")
    return remap360


def _preset(r360, name, size=1600):
    """The planner's own views for a preset (gs360_360PerspCut.py:593-980 through remap360.perspcut)."""
    from bench import preset_views
    return [r360.PerspectiveView(y, p, hf, vf, view_id=vid) for vid, y, p, hf, vf in preset_views(name, size)]


def _to_cuda(a):
    if a.dtype == np.uint16:
        return torch.from_numpy(a.view(np.int16)).cuda().view(torch.uint16)
    return torch.from_numpy(a).cuda()


def _to_numpy(t):
    if t.dtype == torch.uint16:
        return t.view(torch.int16).cpu().numpy().view(np.uint16)
    return t.cpu().numpy()


def _within1(got, want):
    d = np.abs(got.astype(np.int64) - want.astype(np.int64))
    return float((d <= 1).mean()), float((d == 0).mean()), d


def _seam_pole_masks(mx, my, W, H):
    ix, iy = np.floor(mx).astype(np.int64), np.floor(my).astype(np.int64)
    return (ix <= 1) | (ix >= W - 3), (iy <= 1) | (iy >= H - 3)


@pytest.mark.parametrize("interp", ["linear", "cubic"])
def test_cfg1_default_preset_3840_u8(r360, interp):
    pytest.importorskip("cv2")
    views = _preset(r360, "default")
    assert len(views) == 8 and abs(views[0].hfov_deg - 112.6199) < 1e-3          # 12 mm on a 36 mm sensor (PC:77-86)
    W, H, size = 3840, 1920, 1600
    src = np.random.default_rng(21).integers(0, 256, (H, W, 3), dtype=np.uint8)
    dev = _to_cuda(src)[None]
    for path in ("tiled", "direct"):
        got = _to_numpy(r360.remap_erp(dev, views, (size, size), interp=interp, path=path))[0]
        for k, v in enumerate(views):
            mx, my = geo.erp_map64(W, H, size, size, v.yaw_deg, v.pitch_deg, v.hfov_deg, v.vfov_deg)
            ok, exact, _ = _within1(got[k], sampler.sample_cv2(src, mx, my, interp, "erp"))
            assert ok >= OK_FRACTION, (path, v.view_id, ok, exact)


@pytest.mark.parametrize("interp", ["linear", "cubic"])
def test_cfg3_fisheyelike_preset_8k_u8_in_frame_blocks(r360, interp):
    """Config 3 is a video: three frames so that the batch holds one full block of two frames and a left-over one."""
    pytest.importorskip("cv2")
    views = _preset(r360, "fisheyelike")
    assert len(views) == 10 and abs(views[0].hfov_deg - 93.2732) < 1e-3          # 17 mm
    W, H, size = 7680, 3840, 1600
    rng = np.random.default_rng(31)
    frames = rng.integers(0, 256, (3, H, W, 3), dtype=np.uint8)
    got = _to_numpy(r360.remap_erp(_to_cuda(frames), views, (size, size), interp=interp, path="tiled"))
    for k, v in enumerate(views):
        mx, my = geo.erp_map64(W, H, size, size, v.yaw_deg, v.pitch_deg, v.hfov_deg, v.vfov_deg)
        for f in ((0, 1, 2) if k in (0, 5) else (k % 3,)):                         # every frame for two views, one for the rest
            ok, exact, _ = _within1(got[f, k], sampler.sample_cv2(frames[f], mx, my, interp, "erp"))
            assert ok >= OK_FRACTION, (v.view_id, f, ok, exact)


def _cfg4_views(r360):
    fov = 104.2500326978036
    extra = [(180.0, 0.0), (179.9, 0.0), (-179.9, 0.0), (0.0, 90.0), (0.0, -90.0), (40.0, 60.0), (-40.0, -60.0)]
    return _preset(r360, "full360coverage") + [r360.PerspectiveView(y, p, fov, fov, view_id="x%g_%g" % (y, p)) for y, p in extra]


@pytest.fixture(scope="module")
def cfg4_source():
    rng = np.random.default_rng(41)
    W, H = 7680, 3840
    # 16-bit "log" footage: a wide ramp (the legal range of a 10-bit signal scaled to 16 bits) plus sensor noise, and
    # a band of full-range noise so that saturation at both ends is exercised
    ramp = np.linspace(64 * 64, 940 * 64, W, dtype=np.float64)[None, :, None]
    src = np.clip(ramp + rng.normal(0, 900, (H, W, 3)), 0, 65535).astype(np.uint16)
    src[H // 2 - 64:H // 2 + 64] = rng.integers(0, 65536, (128, W, 3)).astype(np.uint16)
    return src


@pytest.mark.parametrize("path", ["tiled", "direct"])
def test_cfg4_8k_u16_bicubic_to_u16(r360, cfg4_source, path):
    pytest.importorskip("cv2")
    src = cfg4_source
    H, W = src.shape[:2]
    size = 1600
    views = _cfg4_views(r360)
    assert len(views) == 19
    dev = _to_cuda(src)[None]
    got = _to_numpy(r360.remap_erp(dev, views, (size, size), interp="cubic", path=path))[0]
    seam_px = pole_px = 0
    for k, v in enumerate(views):
        mx, my = geo.erp_map64(W, H, size, size, v.yaw_deg, v.pitch_deg, v.hfov_deg, v.vfov_deg)
        want = sampler.sample_cv2(src, mx, my, "cubic", "erp")
        ok, exact, d = _within1(got[k], want)
        assert ok >= OK_FRACTION, (path, v.view_id, ok, exact)
        seam, pole = _seam_pole_masks(mx, my, W, H)
        dmax = d.max(axis=2)
        for name, sel in (("seam", seam), ("pole", pole)):
            if sel.any():
                assert (dmax[sel] <= 1).mean() >= OK_FRACTION, (path, v.view_id, name, float((dmax[sel] <= 1).mean()))
        seam_px += int(seam.sum()); pole_px += int(pole.sum())
    # the view set does reach the seam columns and the pole rows (rows 0, 1, H - 2, H - 1 lie within 0.07 degrees of
    # a pole: a disc of about one output pixel radius in each of the two pole views)
    assert seam_px > 10000 and pole_px >= 8, (seam_px, pole_px)


@pytest.mark.parametrize("path", ["tiled", "direct"])
def test_cfg4_8k_u16_bicubic_to_f16(r360, cfg4_source, path):
    """float16 output = value / 65535 rounded to half: no cv2 counterpart, so the NumPy model on every 16th row."""
    src = cfg4_source
    H, W = src.shape[:2]
    size = 1600
    views = _cfg4_views(r360)
    dev = _to_cuda(src)[None]
    got = _to_numpy(r360.remap_erp(dev, views, (size, size), interp="cubic", out_dtype=torch.float16, path=path))[0]
    assert got.dtype == np.float16
    rows = np.arange(0, size, 16)
    for k, v in enumerate(views):
        mx, my = geo.erp_map64(W, H, size, size, v.yaw_deg, v.pitch_deg, v.hfov_deg, v.vfov_deg)
        want = sampler.sample(src, mx[rows], my[rows], "cubic", "erp", out_dtype=np.float16)
        g, w = got[k][rows].astype(np.float64), want.astype(np.float64)
        lsb = np.maximum(np.abs(w), 2.0 ** -14) * 2.0 ** -10
        ok = float((np.abs(g - w) <= lsb).mean())
        assert ok >= OK_FRACTION, (path, v.view_id, ok)
