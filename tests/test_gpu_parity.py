"""GPU parity: the CUDA path (through the C ABI) against the CPU oracle on the same seeded
inputs.  Run on the B200 box with ``pytest -m gpu``.

Bars (BASELINE.json north_star):
  * sample coordinates within 1e-3 px of the float64 oracle map (we also record how far
    inside that bar the kernels are);
  * pixels within 1 LSB of cv2.remap's arithmetic on >= 99.9 % of pixels -- integer paths are
    expected to be bit-exact, which is asserted where it holds by construction;
  * seam columns and pole rows are checked explicitly.
"""

import numpy as np
import pytest

torch = pytest.importorskip("torch")

pytestmark = pytest.mark.gpu

from oracle import geometry as geo  # noqa: E402
from oracle import sampler  # noqa: E402

COORD_TOL_PX = 1e-3        # north_star tolerance
PIXEL_OK_FRACTION = 0.999  # north_star tolerance: |diff| <= 1 LSB on this fraction of pixels


@pytest.fixture(scope="module")
def r360():
    import remap360
    if not torch.cuda.is_available():
        pytest.fail("-m gpu tests need a CUDA device")
    return remap360


def _views(r360, specs, fov=104.2500326978036, vfov=None):
    return [r360.PerspectiveView(yaw_deg=y, pitch_deg=p, hfov_deg=fov, vfov_deg=vfov or fov) for y, p in specs]


FULL360 = [(0, 0), (45, 30), (45, -30), (90, 0), (135, 30), (135, -30), (180, 0), (-135, 30),
           (-135, -30), (-90, 0), (-45, 30), (-45, -30)]
HARD = [(180, 0), (179.9, 0), (-179.9, 0), (0, 90), (0, -90), (30, 60), (-70, -60), (123.4, 89.0)]


def _noise(rng, shape, dtype):
    if dtype == np.float32:
        return rng.random(shape, dtype=np.float32)
    if dtype == np.float16:
        return rng.random(shape, dtype=np.float32).astype(np.float16)
    return rng.integers(0, np.iinfo(dtype).max + 1, shape).astype(dtype)


def _to_cuda(a):
    if a.dtype == np.uint16:
        return torch.from_numpy(a.view(np.int16)).cuda().view(torch.uint16)
    return torch.from_numpy(a).cuda()


def _to_numpy(t):
    if t.dtype == torch.uint16:
        return t.view(torch.int16).cpu().numpy().view(np.uint16)
    return t.cpu().numpy()


# ------------------------------------------------------------------------------------------
# coordinates
# ------------------------------------------------------------------------------------------

@pytest.mark.parametrize("path", ["direct", "tiled"])
@pytest.mark.parametrize("convention", ["halfpixel", "v360"])
@pytest.mark.parametrize("W,H,size", [(7680, 3840, 400), (3840, 1920, 256)])
def test_erp_coordinates_against_float64_oracle(r360, path, W, H, size, convention):
    views = _views(r360, FULL360 + HARD)
    got = r360.sample_coordinates(views, (size, size), erp_size=(W, H), convention=convention, path=path)
    x64, y64 = got["x64"].cpu().numpy(), got["y64"].cpu().numpy()
    x32, y32 = got["x32"].cpu().numpy(), got["y32"].cpu().numpy()
    worst = 0.0
    same32 = []
    for k, v in enumerate(views):
        mx, my = geo.erp_map64(W, H, size, size, v.yaw_deg, v.pitch_deg, v.hfov_deg, v.vfov_deg, convention)
        # longitude is only defined modulo the panorama width (and is arbitrary AT a pole)
        period = W if convention == "halfpixel" else W - 1
        dx = np.abs(x64[k] - mx)
        dx = np.minimum(dx, np.abs(dx - period))
        at_pole = (np.abs(my - (-0.5 if convention == "halfpixel" else 0)) < 1e-6) | \
                  (np.abs(my - ((H - 0.5) if convention == "halfpixel" else H - 1)) < 1e-6)
        dy = np.abs(y64[k] - my)
        worst = max(worst, dx[~at_pole].max(), dy.max())
        d32 = np.abs(x32[k].astype(np.float64) - mx)
        d32 = np.minimum(d32, np.abs(d32 - period))
        assert d32[~at_pole].max() <= COORD_TOL_PX
        assert np.abs(y32[k].astype(np.float64) - my).max() <= COORD_TOL_PX
        same32.append(np.mean((x32[k] == mx.astype(np.float32)) & (y32[k] == my.astype(np.float32))))
    assert worst <= COORD_TOL_PX
    # how far inside the bar: the float64 pre-image should agree to ~1e-5 px or better
    assert worst <= 2e-5, worst
    assert min(same32) > 0.98


@pytest.mark.parametrize("path", ["direct", "tiled"])
def test_fisheye_coordinates_against_oracle_and_reference_maps(r360, path, golden_df, golden_df_maps):
    cal = golden_df["sensors"]["0"]
    calib = r360.FisheyeCalibration(**{k: cal[k] for k in ("width", "height", "f", "cx", "cy", "k1", "k2", "k3",
                                                           "k4", "p1", "p2", "b1", "b2")}, lens_fov_deg=190.0)
    specs = golden_df["sfm10_default"]
    views, keys = [], []
    for spec in specs:
        for slot, (lens_key, lens_yaw) in enumerate((("X", 0.0), ("Y", 180.0))):
            views.append(r360.PerspectiveView(geo.wrap_angle_deg(spec["yaw_deg"] - lens_yaw), spec["pitch_deg"],
                                              spec["hfov_deg"], spec["vfov_deg"], src_slot=slot))
            keys.append("m96_%s_%s" % (spec["view_id"], lens_key))
    got = r360.sample_coordinates(views, (96, 96), calibs=[calib, calib], path=path)
    x64, y64, valid = got["x64"].cpu().numpy(), got["y64"].cpu().numpy(), got["valid"].cpu().numpy().astype(bool)
    for k, v in enumerate(views):
        mx, my, ok = geo.fisheye_map64(cal, v.yaw_deg, v.pitch_deg, v.hfov_deg, v.vfov_deg, 96, 96, 190.0)
        assert np.array_equal(valid[k], ok)
        front = mx == mx  # all finite
        sel = ok | (np.hypot(mx - 1920, my - 1920) < 4000)
        assert np.abs(x64[k] - mx)[sel & front].max() <= 2e-5
        assert np.abs(y64[k] - my)[sel & front].max() <= 2e-5
        # the reference's own float32 maps (recorded by tests/golden/make_golden.py)
        assert np.array_equal(valid[k], golden_df_maps[keys[k] + "_v"])
        assert np.abs(x64[k] - golden_df_maps[keys[k] + "_x"])[ok].max() < 0.01 if ok.any() else True
        assert np.abs(y64[k] - golden_df_maps[keys[k] + "_y"])[ok].max() < 0.01 if ok.any() else True


@pytest.mark.parametrize("path", ["direct", "tiled"])
def test_fisheye_coordinates_every_distortion_term(r360, path, golden_df, golden_df_maps):
    cal = golden_df["synthetic_calibration"]
    calib = r360.FisheyeCalibration(**{k: cal[k] for k in ("width", "height", "f", "cx", "cy", "k1", "k2", "k3",
                                                           "k4", "p1", "p2", "b1", "b2")}, lens_fov_deg=185.0)
    views = [r360.PerspectiveView(y, p, 100.0, 80.0) for y, p in golden_df["synthetic_views"]]
    got = r360.sample_coordinates(views, (80, 64), calibs=[calib], path=path)
    for n, v in enumerate(views):
        mx, my, ok = geo.fisheye_map64(cal, v.yaw_deg, v.pitch_deg, 100.0, 80.0, 80, 64, 185.0)
        assert np.array_equal(got["valid"][n].cpu().numpy().astype(bool), ok)
        assert np.abs(got["x64"][n].cpu().numpy() - mx)[ok].max() <= 2e-5
        assert np.abs(got["y64"][n].cpu().numpy() - my)[ok].max() <= 2e-5
        assert np.abs(got["x64"][n].cpu().numpy() - golden_df_maps["syn%d_x" % n])[ok].max() < 6e-3


# ------------------------------------------------------------------------------------------
# pixels, panorama source
# ------------------------------------------------------------------------------------------

def _oracle_erp_views(src, views, size, interp, convention="halfpixel", out_dtype=None):
    H, W = src.shape[:2]
    outs = []
    for v in views:
        mx, my = geo.erp_map64(W, H, size, size, v.yaw_deg, v.pitch_deg, v.hfov_deg, v.vfov_deg, convention)
        outs.append(sampler.sample(src, mx, my, interp, "erp", out_dtype=out_dtype))
    return np.stack(outs)


def _lsb_stats(got, want):
    if got.dtype == np.float16 or got.dtype == np.float32:
        g, w = got.astype(np.float64), want.astype(np.float64)
        lsb = np.maximum(np.abs(w), 2.0 ** -14) * (2.0 ** -10 if got.dtype == np.float16 else 2.0 ** -22)
        d = np.abs(g - w) / lsb
    else:
        d = np.abs(got.astype(np.int64) - want.astype(np.int64)).astype(np.float64)
    return float((d == 0).mean()), float((d <= 1).mean()), float(d.max())


@pytest.mark.parametrize("path", ["direct", "tiled"])
@pytest.mark.parametrize("dtype", [np.uint8, np.uint16, np.float16, np.float32])
@pytest.mark.parametrize("interp", ["nearest", "linear", "cubic", "lanczos4"])
@pytest.mark.parametrize("channels", [1, 3, 4])
def test_erp_pixels_small_noise_all_types(r360, path, dtype, interp, channels):
    rng = np.random.default_rng(1234)
    W, H, size = 768, 384, 128
    src = _noise(rng, (H, W, channels), dtype)
    views = _views(r360, [(0, 0), (45, 30), (180, 0), (-179.9, -20), (0, 90), (10, -90), (77, 60)], fov=100.0)
    got = _to_numpy(r360.remap_erp(_to_cuda(src)[None], views, (size, size), interp=interp, path=path))[0]
    want = _oracle_erp_views(src, views, size, interp)
    exact, within1, worst = _lsb_stats(got, want)
    assert within1 >= PIXEL_OK_FRACTION, (exact, within1, worst)
    if dtype in (np.uint8, np.uint16) or interp == "nearest":
        # integer / copy paths: a mismatch can only come from a coordinate landing in another 1/32 bin
        assert exact >= 0.9995, (exact, within1, worst)


@pytest.mark.parametrize("path", ["direct", "tiled"])
@pytest.mark.parametrize("interp", ["linear", "cubic"])
def test_erp_pixels_full_size_u8_against_cv2(r360, path, interp):
    """8K -> 1600^2, noise content (worst case), seam / pole / preset views, checked against the
    real cv2.remap fed with the float64 oracle map."""
    pytest.importorskip("cv2")
    rng = np.random.default_rng(1234)
    W, H, size = 7680, 3840, 1600
    src = _noise(rng, (H, W, 3), np.uint8)
    specs = [(0, 0), (45, 30), (180, 0), (0, 90), (-135, -30), (30, -60)]
    views = _views(r360, specs)
    got = _to_numpy(r360.remap_erp(_to_cuda(src)[None], views, (size, size), interp=interp, path=path))[0]
    for k, v in enumerate(views):
        mx, my = geo.erp_map64(W, H, size, size, v.yaw_deg, v.pitch_deg, v.hfov_deg, v.vfov_deg)
        want = sampler.sample_cv2(src, mx, my, interp, "erp")
        exact, within1, worst = _lsb_stats(got[k], want)
        assert within1 >= PIXEL_OK_FRACTION, (specs[k], exact, within1, worst)
        # seam columns and pole rows, explicitly
        ix = np.floor(mx).astype(np.int64)
        iy = np.floor(my).astype(np.int64)
        seam = (ix <= 1) | (ix >= W - 3)
        pole = (iy <= 1) | (iy >= H - 3)
        for name, sel in (("seam", seam), ("pole", pole)):
            if sel.any():
                d = np.abs(got[k].astype(np.int64) - want.astype(np.int64)).max(axis=2)
                assert (d[sel] <= 1).mean() >= PIXEL_OK_FRACTION, (specs[k], name, (d[sel] <= 1).mean())


@pytest.mark.parametrize("path", ["direct", "tiled"])
def test_erp_pixels_u16_to_f16_and_u16_cubic_seam_pole(r360, path):
    """BASELINE config 4: 16-bit linear frames, bicubic, uint16 and fp16 outputs."""
    rng = np.random.default_rng(99)
    W, H, size = 2048, 1024, 256
    ramp = np.linspace(64 * 256, 940 * 256, W, dtype=np.float64)[None, :, None]
    src = np.clip(ramp + rng.normal(0, 600, (H, W, 3)), 0, 65535).astype(np.uint16)
    views = _views(r360, [(180, 0), (179.9, 0), (-179.9, 0), (0, 90), (0, -90), (40, 60), (-40, -60), (0, 0)])
    dev = _to_cuda(src)[None]
    got16 = _to_numpy(r360.remap_erp(dev, views, (size, size), interp="cubic", path=path))[0]
    want16 = _oracle_erp_views(src, views, size, "cubic")
    exact, within1, worst = _lsb_stats(got16, want16)
    assert within1 >= PIXEL_OK_FRACTION and exact > 0.99, (exact, within1, worst)
    goth = _to_numpy(r360.remap_erp(dev, views, (size, size), interp="cubic", out_dtype=torch.float16, path=path))[0]
    wanth = _oracle_erp_views(src, views, size, "cubic", out_dtype=np.float16)
    exact, within1, worst = _lsb_stats(goth, wanth)
    assert goth.dtype == np.float16 and within1 >= PIXEL_OK_FRACTION, (exact, within1, worst)


@pytest.mark.parametrize("path", ["direct", "tiled"])
def test_erp_v360_convention_and_roll(r360, path):
    rng = np.random.default_rng(5)
    W, H, size = 1024, 512, 96
    src = _noise(rng, (H, W, 3), np.uint8)
    v = r360.PerspectiveView(33.0, -12.0, 95.0, 70.0, roll_deg=17.0)
    got = _to_numpy(r360.remap_erp(_to_cuda(src)[None], [v], (size, size), interp="linear",
                                   convention="v360", path=path))[0, 0]
    mx, my = geo.erp_map64(W, H, size, size, 33.0, -12.0, 95.0, 70.0, "v360", roll_deg=17.0)
    want = sampler.sample(src, mx, my, "linear", "erp")
    exact, within1, _ = _lsb_stats(got, want)
    assert within1 >= PIXEL_OK_FRACTION and exact > 0.999


@pytest.mark.parametrize("path", ["direct", "tiled"])
def test_batched_frames_equal_single_calls_and_ragged_sizes(r360, path):
    """Frame-major batching, more than 16 views (launch chunking), and an output size that is
    not a multiple of any tile size."""
    rng = np.random.default_rng(8)
    W, H = 640, 320
    frames = _noise(rng, (3, H, W, 3), np.uint8)
    views = _views(r360, [(k * 17.0 - 170, (k % 5) * 20.0 - 40) for k in range(19)], fov=90.0, vfov=60.0)
    dev = _to_cuda(frames)
    out = r360.remap_erp(dev, views, (77, 45), interp="cubic", path=path)
    assert tuple(out.shape) == (3, 19, 45, 77, 3)
    for f in range(3):
        single = r360.remap_erp(dev[f:f + 1], views, (77, 45), interp="cubic", path=path)
        assert torch.equal(single[0], out[f])
    mx, my = geo.erp_map64(W, H, 77, 45, views[18].yaw_deg, views[18].pitch_deg, 90.0, 60.0)
    ref = sampler.sample(frames[2], mx, my, "cubic", "erp")
    assert (np.abs(_to_numpy(out[2, 18]).astype(int) - ref.astype(int)) <= 1).mean() >= PIXEL_OK_FRACTION


@pytest.mark.parametrize("path", ["direct", "tiled"])
def test_row_pitch_and_strided_views_of_larger_buffers(r360, path):
    rng = np.random.default_rng(11)
    W, H, size = 500, 250, 64
    big = _to_cuda(_noise(rng, (2, H, W + 12, 3), np.uint8))
    frames = big[:, :, 5:5 + W, :]                      # row pitch > W * C, unaligned start
    views = _views(r360, [(20, 10), (-160, -30)], fov=80.0)
    outbig = torch.zeros((2, 2, size, size + 8, 3), dtype=torch.uint8, device="cuda")
    # write into a padded destination (row pitch > w * C) through the descriptor layer
    from remap360 import api
    src = api._describe(frames, "frames")
    dst = api._describe(outbig.view(4, size, size + 8, 3)[:, :, :size, :], "out")
    api._run(src, dst, views, api._options("linear", path=path), path, frames.device, None)
    ref = r360.remap_erp(frames.contiguous(), views, (size, size), interp="linear", path=path)
    assert torch.equal(outbig[:, :, :, :size, :], ref)
    assert int(outbig[:, :, :, size:, :].abs().sum()) == 0     # padding untouched


# ------------------------------------------------------------------------------------------
# pixels, dual-fisheye source
# ------------------------------------------------------------------------------------------

def _template_calib(r360, golden_df, fov=190.0):
    cal = golden_df["sensors"]["0"]
    return cal, r360.FisheyeCalibration(**{k: cal[k] for k in ("width", "height", "f", "cx", "cy", "k1", "k2",
                                                               "k3", "k4", "p1", "p2", "b1", "b2")},
                                        lens_fov_deg=fov)


@pytest.mark.parametrize("path", ["direct", "tiled"])
@pytest.mark.parametrize("dtype", [np.uint8, np.uint16])
@pytest.mark.parametrize("interp", ["nearest", "linear", "cubic", "lanczos4"])
def test_fisheye_pixels_with_invalid_fill(r360, path, dtype, interp, golden_df):
    """Scaled-down sensor so that sensor edges, the lens-FOV circle and fully invalid views all occur."""
    rng = np.random.default_rng(21)
    cal = dict(golden_df["synthetic_calibration"])
    cal.update(width=600, height=560, f=164.1, cx=2.45, cy=-1.5)
    calib = r360.FisheyeCalibration(**{k: cal[k] for k in ("width", "height", "f", "cx", "cy", "k1", "k2", "k3",
                                                           "k4", "p1", "p2", "b1", "b2")}, lens_fov_deg=185.0)
    pair = _noise(rng, (1, 2, 560, 600, 3), dtype)
    specs = [(0, 0, 0), (35, 20, 0), (-60, -35, 1), (95, 5, 0), (170, 0, 1), (0, 80, 1)]
    views = [r360.PerspectiveView(y, p, 100.0, 80.0, src_slot=s) for y, p, s in specs]
    for fill, bv in ((True, 0), (True, 37), (False, 9)):
        got = _to_numpy(r360.remap_fisheye(_to_cuda(pair), [calib, calib], views, (120, 90), interp=interp,
                                           border_value=bv, fill_invalid=fill, path=path))[0]
        for k, (y, p, s) in enumerate(specs):
            mx, my, ok = geo.fisheye_map64(cal, y, p, 100.0, 80.0, 120, 90, 185.0)
            want = sampler.sample(pair[0, s], mx, my, interp, "constant", bv)
            if fill:
                want = sampler.apply_invalid_fill(want, ok, bv)
            exact, within1, worst = _lsb_stats(got[k], want)
            assert within1 >= PIXEL_OK_FRACTION, (specs[k], fill, bv, exact, within1, worst)
            assert exact >= 0.999


@pytest.mark.parametrize("path", ["direct", "tiled"])
def test_fisheye_default_layout_full_size_against_cv2(r360, path, golden_df):
    """BASELINE config 5 at full size: 3840^2 lens pair, template calibration, 10 x 1750^2, cubic."""
    pytest.importorskip("cv2")
    rng = np.random.default_rng(1234)
    cal, calib = _template_calib(r360, golden_df)
    pair = _noise(rng, (1, 2, 3840, 3840, 3), np.uint8)
    specs = golden_df["sfm10_default"]
    lens_of = {vid: (0 if info["lens_key"] == "X" else 1) for vid, info in golden_df["maps_1750"]["views"].items()}
    views = [r360.PerspectiveView(geo.wrap_angle_deg(s["yaw_deg"] - (0.0, 180.0)[lens_of[s["view_id"]]]),
                                  s["pitch_deg"], s["hfov_deg"], s["vfov_deg"], src_slot=lens_of[s["view_id"]])
             for s in specs]
    got = _to_numpy(r360.remap_fisheye(_to_cuda(pair), [calib, calib], views, (1750, 1750), interp="cubic",
                                       path=path))[0]
    for k in (0, 1, 3, 5, 8):
        v = views[k]
        mx, my, ok = geo.fisheye_map64(cal, v.yaw_deg, v.pitch_deg, v.hfov_deg, v.vfov_deg, 1750, 1750, 190.0)
        assert ok.all()
        want = sampler.sample_cv2(pair[0, v.src_slot], mx, my, "cubic", "constant", 0)
        exact, within1, worst = _lsb_stats(got[k], want)
        assert within1 >= PIXEL_OK_FRACTION, (specs[k]["view_id"], exact, within1, worst)


# ------------------------------------------------------------------------------------------
# size-independent properties and error behaviour
# ------------------------------------------------------------------------------------------

@pytest.mark.parametrize("path", ["direct", "tiled"])
def test_constant_image_stays_constant_and_direct_equals_tiled(r360, path):
    W, H = 7680, 3840
    for dtype, val in ((torch.uint8, 201), (torch.float32, 0.625)):
        src = torch.full((1, H, W, 3), val, dtype=dtype, device="cuda")
        views = _views(r360, FULL360[:4] + HARD[:5])
        for interp in ("linear", "cubic"):
            out = r360.remap_erp(src, views, (512, 512), interp=interp, path=path)
            if dtype == torch.uint8:
                assert int(out.min()) == val and int(out.max()) == val
            else:
                assert float((out - val).abs().max()) < 1e-6


def test_direct_and_tiled_paths_agree_at_full_size(r360):
    rng = np.random.default_rng(3)
    src = _to_cuda(_noise(rng, (1, 3840, 7680, 3), np.uint8))
    views = _views(r360, FULL360)
    for interp in ("linear", "cubic"):
        a = r360.remap_erp(src, views, (1600, 1600), interp=interp, path="direct")
        b = r360.remap_erp(src, views, (1600, 1600), interp=interp, path="tiled")
        d = (a.to(torch.int16) - b.to(torch.int16)).abs()
        assert float((d <= 1).float().mean()) >= PIXEL_OK_FRACTION
        assert float((d == 0).float().mean()) >= 0.999


def test_bad_arguments_raise(r360):
    src = torch.zeros((1, 64, 128, 3), dtype=torch.uint8, device="cuda")
    v = _views(r360, [(0, 0)])
    with pytest.raises(ValueError):
        r360.remap_erp(src.cpu(), v, (32, 32))
    with pytest.raises(ValueError):
        r360.remap_erp(src, v, (32, 32), interp="lanczos9")
    with pytest.raises(r360.Remap360Error):
        r360.remap_erp(src, [], (32, 32))
    with pytest.raises(r360.Remap360Error):
        r360.remap_erp(torch.zeros((1, 64, 128, 5), dtype=torch.uint8, device="cuda"), v, (32, 32))
    with pytest.raises(r360.Remap360Error):
        r360.remap_erp(src, v, (32, 32), out_dtype=torch.float32)       # u8 -> f32 is not built
    with pytest.raises(r360.Remap360Error):
        r360.remap_erp(src, [r360.PerspectiveView(0, 0, 90, 90, src_slot=1)], (32, 32))
    with pytest.raises(TypeError):
        r360.remap_erp(src.to(torch.int32), v, (32, 32))


def test_launch_counter_moves(r360):
    src = torch.zeros((1, 64, 128, 3), dtype=torch.uint8, device="cuda")
    before = r360.launch_count()
    r360.remap_erp(src, _views(r360, [(0, 0)]), (32, 32))
    assert r360.launch_count() > before


# ------------------------------------------------------------------------------------------
# fisheye -> undistorted fisheye (DF:1008-1217)
# ------------------------------------------------------------------------------------------

def _calib_from(r360, cal, fov):
    return r360.FisheyeCalibration(**{k: cal[k] for k in ("width", "height", "f", "cx", "cy", "k1", "k2", "k3",
                                                           "k4", "p1", "p2", "b1", "b2")}, lens_fov_deg=fov)


@pytest.mark.parametrize("path", ["direct", "tiled"])
def test_undistort_coordinates_against_oracle_and_reference_maps(r360, path, golden_undistort):
    meta, maps = golden_undistort
    for name, case in meta["cases"].items():
        cal, st = case["calibration"], case["stride"]
        w, h = int(cal["width"]), int(cal["height"])
        items = [r360.UndistortItem(case["undistort_zoom"], 0)]
        got = r360.sample_coordinates(items, (w, h), calibs=[_calib_from(r360, cal, case["lens_fov_deg"])], path=path)
        mx, my, ok, _ = geo.undistort_map64(cal, case["undistort_zoom"], case["lens_fov_deg"])
        x64, y64 = got["x64"][0].cpu().numpy(), got["y64"][0].cpu().numpy()
        assert np.array_equal(got["valid"][0].cpu().numpy().astype(bool), ok), name
        assert np.abs(x64 - mx)[ok].max() <= 2e-5 and np.abs(y64 - my)[ok].max() <= 2e-5, name
        # the reference's own float32 maps
        assert np.abs(x64[::st, ::st] - maps[name + "_x"])[ok[::st, ::st]].max() < 2e-3, name
        assert np.abs(y64[::st, ::st] - maps[name + "_y"])[ok[::st, ::st]].max() < 2e-3, name
        assert np.array_equal(ok[::st, ::st], maps[name + "_v"])


@pytest.mark.parametrize("path", ["direct", "tiled"])
@pytest.mark.parametrize("dtype", [np.uint8, np.uint16])
@pytest.mark.parametrize("interp", ["nearest", "linear", "cubic", "lanczos4"])
def test_undistort_pixels(r360, path, dtype, interp, golden_undistort):
    """Two lens images, three outputs (auto zoom, zoom < 1, a narrower lens FOV), mask fill on and off."""
    meta, _ = golden_undistort
    rng = np.random.default_rng(5)
    cases = [("wide_auto", 0), ("tiny_z09_fov150", 1)]
    for name, slot in cases:
        case = meta["cases"][name]
        cal = case["calibration"]
        w, h = int(cal["width"]), int(cal["height"])
        pair = _noise(rng, (2, 2, h, w, 3), dtype)
        calib = _calib_from(r360, cal, case["lens_fov_deg"])
        items = [r360.UndistortItem(case["undistort_zoom"], slot), r360.UndistortItem(1.31, 1 - slot)]
        for fill, bv in ((True, 0), (False, 23)):
            got = _to_numpy(r360.undistort_fisheye(_to_cuda(pair), [calib, calib], items, interp=interp,
                                                   border_value=bv, fill_invalid=fill, path=path))
            assert got.shape == (2, 2, h, w, 3)
            for g in range(2):
                for k, it in enumerate(items):
                    mx, my, ok, _ = geo.undistort_map64(cal, it.zoom, case["lens_fov_deg"])
                    want = sampler.sample(pair[g, it.src_slot], mx, my, interp, "constant", bv)
                    if fill:
                        want = sampler.apply_invalid_fill(want, ok, bv)
                    exact, within1, worst = _lsb_stats(got[g, k], want)
                    assert within1 >= PIXEL_OK_FRACTION and exact >= 0.999, (name, g, k, fill, exact, within1, worst)


def test_undistort_full_size_against_cv2(r360, golden_undistort):
    """3840^2 template calibration, u8 cubic, tiled path, checked against the real cv2.remap + mask fill
    (what undistort_prepared_image does, DF:1198-1211) on the float64 oracle map."""
    cv2 = pytest.importorskip("cv2")
    meta, _ = golden_undistort
    case = meta["cases"]["tmpl_auto"]
    cal = case["calibration"]
    w, h = int(cal["width"]), int(cal["height"])
    rng = np.random.default_rng(77)
    img = _noise(rng, (h, w, 3), np.uint8)
    got = _to_numpy(r360.undistort_fisheye(_to_cuda(img)[None, None], [_calib_from(r360, cal, 190.0)],
                                           [r360.UndistortItem(case["undistort_zoom"], 0)], interp="cubic"))[0, 0]
    mx, my, ok, _ = geo.undistort_map64(cal, case["undistort_zoom"], 190.0)
    want = cv2.remap(img, mx.astype(np.float32), my.astype(np.float32), cv2.INTER_CUBIC,
                     borderMode=cv2.BORDER_CONSTANT, borderValue=0.0)
    want[~ok] = 0
    exact, within1, worst = _lsb_stats(got, want)
    assert within1 >= PIXEL_OK_FRACTION and exact >= 0.999, (exact, within1, worst)


def test_undistort_argument_errors(r360):
    dev = torch.zeros((1, 1, 16, 16, 3), dtype=torch.uint8, device="cuda")
    good = r360.FisheyeCalibration(16, 16, 10.0)
    with pytest.raises(r360.Remap360Error):
        r360.undistort_fisheye(dev, [r360.FisheyeCalibration(16, 16, 0.0)], [r360.UndistortItem(1.0, 0)], path="direct")
    with pytest.raises(r360.Remap360Error):
        r360.undistort_fisheye(dev, [good], [r360.UndistortItem(1.0, 3)], path="direct")
    with pytest.raises(r360.Remap360Error):
        r360.undistort_fisheye(dev, [good], [])


# ------------------------------------------------------------------------------------------
# panorama -> equidistant fisheye views (v360 output=fisheye; preset fisheyeXY, PC:871-887)
# ------------------------------------------------------------------------------------------

@pytest.mark.parametrize("path", ["direct", "tiled"])
def test_fisheye_output_coordinates_and_pixels(r360, path):
    rng = np.random.default_rng(31)
    W, H, size = 2048, 1024, 360
    src = _noise(rng, (H, W, 3), np.uint8)
    hf, vf = geo.fisheye_fov_from_dfov(180.0, size, size)
    assert r360.api.fisheye_fov_from_dfov(180.0, size, size) == (hf, vf)
    specs = [(0.0, 0.0), (180.0, 0.0), (40.0, -25.0)]
    views = [r360.PerspectiveView(y, p, hf, vf, projection="fisheye") for y, p in specs]
    views.append(r360.PerspectiveView(10.0, 5.0, 100.0, 100.0))              # mixed with a rectilinear view
    got = r360.sample_coordinates(views, (size, size), erp_size=(W, H), path=path)
    out = _to_numpy(r360.remap_erp(_to_cuda(src)[None], views, (size, size), interp="cubic", path=path))[0]
    for k, v in enumerate(views):
        mx, my = geo.erp_map64(W, H, size, size, v.yaw_deg, v.pitch_deg, v.hfov_deg, v.vfov_deg,
                               projection=v.projection)
        dx = np.abs(got["x64"][k].cpu().numpy() - mx)
        dx = np.minimum(dx, np.abs(dx - W))                                   # same point across the seam
        assert dx.max() <= 2e-5 and np.abs(got["y64"][k].cpu().numpy() - my).max() <= 2e-5, (k, dx.max())
        want = sampler.sample(src, mx, my, "cubic", "erp")
        exact, within1, worst = _lsb_stats(out[k], want)
        assert within1 >= PIXEL_OK_FRACTION and exact >= 0.999, (k, exact, within1, worst)
    with pytest.raises(ValueError):
        r360.remap_erp(_to_cuda(src)[None], [r360.PerspectiveView(0, 0, 90, 90, projection="cube")], (8, 8))


def test_concurrent_callers_with_different_layouts(r360):
    """Host threads driving the same kernel instantiation with different shared-memory needs (1- and
    3-channel bicubic) on their own streams: results equal the single-threaded ones."""
    import threading
    rng = np.random.default_rng(8)
    W, H, size = 1024, 512, 160
    views = _views(r360, [(0, 0), (120, 20), (180, 0)], fov=100.0)
    srcs = {c: _noise(rng, (H, W, c), np.uint8) for c in (1, 3, 4)}
    want = {c: _to_numpy(r360.remap_erp(_to_cuda(srcs[c])[None], views, (size, size), interp="cubic")) for c in srcs}
    errors = []

    def work(c):
        try:
            stream = torch.cuda.Stream()
            dev = _to_cuda(srcs[c])[None]
            for _ in range(6):
                with torch.cuda.stream(stream):
                    got = r360.remap_erp(dev, views, (size, size), interp="cubic", stream=stream)
                stream.synchronize()
                if not np.array_equal(_to_numpy(got), want[c]):
                    errors.append("mismatch for %d channels" % c)
        except Exception as exc:   # noqa: BLE001
            errors.append("%d channels: %r" % (c, exc))

    threads = [threading.Thread(target=work, args=(c,)) for c in (1, 3, 4, 3, 1)]
    for t in threads:
        t.start()
    for t in threads:
        t.join()
    assert not errors, errors
