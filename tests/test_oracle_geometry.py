"""Float64 coordinate oracle (oracle/geometry.py) against reference-derived fixtures
and against a literal scalar restatement of gs360_GUI.py:342-424."""

import math

import numpy as np
import pytest

from oracle import geometry as g


# --- scalar restatement of the GUI functions (gs360_GUI.py:342-395, :419-424) -------------

def _gui_lonlat(u, v, hfov, vfov, yaw, pitch):
    x, y, z = math.tan(hfov / 2.0) * u, math.tan(vfov / 2.0) * (-v), 1.0
    n = math.sqrt(x * x + y * y + z * z)
    x, y, z = x / n, y / n, z / n
    cp, sp = math.cos(pitch), math.sin(pitch)
    x, y, z = x, cp * y + sp * z, -sp * y + cp * z
    cy, sy = math.cos(yaw), math.sin(yaw)
    x, y, z = cy * x + sy * z, y, -sy * x + cy * z
    return math.atan2(x, z), math.asin(max(-1.0, min(1.0, y)))


@pytest.mark.parametrize("yaw,pitch", [(0, 0), (45, 30), (180, 0), (-135, -30), (0, 90), (0, -90), (179.9, 60)])
def test_erp_map_matches_gui_scalar_math(yaw, pitch):
    W, H, w, h, fov = 7680, 3840, 64, 48, 104.2500326978036
    vf = 90.0
    mx, my = g.erp_map64(W, H, w, h, yaw, pitch, fov, vf, "halfpixel")
    rng = np.random.default_rng(0)
    for _ in range(200):
        i, j = int(rng.integers(0, w)), int(rng.integers(0, h))
        u, v = (i + 0.5) / w * 2 - 1, (j + 0.5) / h * 2 - 1
        lon, lat = _gui_lonlat(u, v, math.radians(fov), math.radians(vf), math.radians(yaw), math.radians(pitch))
        # gs360_GUI.py:419-424 is edge-origin; the halfpixel convention is that minus 0.5
        ex = (lon / (2 * math.pi) + 0.5) * W - 0.5
        ey = (0.5 - lat / math.pi) * H - 0.5
        assert abs(mx[j, i] - ex) < 1e-8 and abs(my[j, i] - ey) < 1e-8


def test_view_centre_lands_on_yaw_pitch():
    W, H = 3840, 1920
    for yaw, pitch in [(0, 0), (90, 0), (-45, 30), (135, -30)]:
        mx, my = g.erp_map64(W, H, 2, 2, yaw, pitch, 1e-3, 1e-3)   # vanishing FOV -> centre ray
        assert abs(mx.mean() - ((yaw / 360 + 0.5) * W - 0.5)) < 1e-3
        assert abs(my.mean() - ((0.5 - pitch / 180) * H - 0.5)) < 1e-3


def test_conventions_differ_by_at_most_half_a_pixel():
    a = g.erp_map64(7680, 3840, 40, 40, 30, 10, 100, 100, "halfpixel")
    b = g.erp_map64(7680, 3840, 40, 40, 30, 10, 100, 100, "v360")
    assert np.abs(a[0] - b[0]).max() <= 0.5 + 1e-9 and np.abs(a[1] - b[1]).max() <= 0.5 + 1e-9
    with pytest.raises(ValueError):
        g.erp_map64(8, 4, 2, 2, 0, 0, 90, 90, "nope")


def test_seam_and_pole_ranges():
    W, H = 7680, 3840
    mx, my = g.erp_map64(W, H, 200, 200, 180.0, 0.0, 112.6, 112.6)
    assert mx.min() >= -0.5 and mx.max() < W - 0.5
    assert (mx[:, :100] > W / 2).all() and (mx[:, 100:] < W / 2).all()      # seam splits the view
    mx, my = g.erp_map64(W, H, 200, 200, 0.0, 90.0, 112.6, 112.6)
    assert my.min() >= -0.5 and my.min() < 20                               # reaches the pole rows
    mx, my = g.erp_map64(W, H, 200, 200, 0.0, -90.0, 112.6, 112.6)
    assert my.max() <= H - 0.5 and my.max() > H - 20


def test_rotation_describes_the_camera_the_pose_exporter_writes():
    """cli_tools/gs360_MS360xmlToPersCams.py:292-353: R_gl = Ry(-yaw) . Rx(pitch) in GL axes
    (x right, y up, z backwards).  Flipping z on both sides gives the y-up / z-forward
    rotation the remap uses."""
    def rot_x(d):
        c, s = math.cos(math.radians(d)), math.sin(math.radians(d))
        return np.array([[1, 0, 0], [0, c, -s], [0, s, c]])

    def rot_y(d):
        c, s = math.cos(math.radians(d)), math.sin(math.radians(d))
        return np.array([[c, 0, s], [0, 1, 0], [-s, 0, c]])

    flip = np.diag([1.0, 1.0, -1.0])
    for yaw, pitch in [(0, 0), (45, 30), (-135, -30), (180, 0), (10, 90)]:
        r_gl = rot_y(-yaw) @ rot_x(pitch)
        assert np.allclose(flip @ r_gl @ flip, g.view_rotation(yaw, pitch), atol=1e-12)


def test_roll_is_a_rotation_about_the_view_axis():
    r = g.view_rotation(20, 10, 33)
    assert np.allclose(r @ r.T, np.eye(3), atol=1e-12)
    assert np.allclose(r[:, 2], g.view_rotation(20, 10, 0)[:, 2], atol=1e-12)
    a = g.camera_rays(8, 8, 90, 90, 20, 10, 33)
    b = (g.camera_rays(8, 8, 90, 90, 0, 0, 0) @ r.T)
    assert np.allclose(a, b, atol=1e-12)


# --- dual fisheye against maps produced by the reference ----------------------------------

def test_fisheye_maps_match_reference_float32_maps(golden_df, golden_df_maps):
    calib = golden_df["sensors"]["0"]
    worst_valid = 0.0
    for spec in golden_df["sfm10_default"]:
        for lens_key, lens_yaw in (("X", 0.0), ("Y", 180.0)):
            yaw_rel = g.wrap_angle_deg(spec["yaw_deg"] - lens_yaw)
            mx, my, valid = g.fisheye_map64(calib, yaw_rel, spec["pitch_deg"], spec["hfov_deg"],
                                            spec["vfov_deg"], 96, 96, 190.0)
            key = "m96_%s_%s" % (spec["view_id"], lens_key)
            assert np.array_equal(valid, golden_df_maps[key + "_v"])
            ref_x, ref_y = golden_df_maps[key + "_x"], golden_df_maps[key + "_y"]
            # the reference computes in float32 (arccos near the axis costs it ~1e-2 px)
            assert np.abs(mx - ref_x).max() < 0.03 and np.abs(my - ref_y).max() < 0.03
            if valid.any():
                worst_valid = max(worst_valid, np.abs(mx - ref_x)[valid].max(), np.abs(my - ref_y)[valid].max())
    assert worst_valid < 0.01


def test_lens_choice_and_full_size_samples(golden_df, golden_df_maps):
    calib = golden_df["sensors"]["0"]
    maps = g.dualfisheye_view_maps(calib, calib, golden_df["sfm10_default"])
    info = golden_df["maps_1750"]
    s = info["stride"]
    for vid, m in maps.items():
        assert m["lens_key"] == info["views"][vid]["lens_key"]
        assert float(np.mean(m["valid"])) == info["views"][vid]["valid_ratio"] == 1.0
        assert np.abs(m["map_x"][::s, ::s] - golden_df_maps["s1750_%s_x" % vid]).max() < 0.06
        assert np.abs(m["map_y"][::s, ::s] - golden_df_maps["s1750_%s_y" % vid]).max() < 0.06
        # ... and tight away from the optical axis, where float32 arccos is well conditioned
        far = np.hypot(m["map_x"][::s, ::s] - 1920, m["map_y"][::s, ::s] - 1920) > 300
        assert np.abs(m["map_x"][::s, ::s] - golden_df_maps["s1750_%s_x" % vid])[far].max() < 2e-3


def test_every_distortion_term(golden_df, golden_df_maps):
    calib = golden_df["synthetic_calibration"]
    for n, (yaw, pitch) in enumerate(golden_df["synthetic_views"]):
        mx, my, valid = g.fisheye_map64(calib, yaw, pitch, 100.0, 80.0, 80, 64, 185.0)
        assert np.array_equal(valid, golden_df_maps["syn%d_v" % n])
        assert np.abs(mx - golden_df_maps["syn%d_x" % n])[valid].max() < 6e-3
        assert np.abs(my - golden_df_maps["syn%d_y" % n])[valid].max() < 6e-3


def test_wrap_angle(golden_df):
    for a, want in golden_df["wrap_angle_deg"]:
        assert g.wrap_angle_deg(a) == want


# --- fisheye -> undistorted fisheye (DF:1008-1170) -------------------------------------------

def test_undistort_maps_match_reference_float32_maps(golden_undistort):
    meta, maps = golden_undistort
    for name, case in meta["cases"].items():
        st = case["stride"]
        mx, my, valid, _ = g.undistort_map64(case["calibration"], case["undistort_zoom"], case["lens_fov_deg"])
        mx, my, valid = mx[::st, ::st], my[::st, ::st], valid[::st, ::st]
        assert np.array_equal(valid, maps[name + "_v"]), name
        assert abs(valid.mean() - case["valid_ratio"]) < 0.02
        # the reference evaluates in float32 on coordinates up to 3840: ~1e-3 px of rounding
        assert np.abs(mx - maps[name + "_x"]).max() < 2e-3, name
        assert np.abs(my - maps[name + "_y"]).max() < 2e-3, name


def test_auto_undistort_zoom_matches_reference(golden_undistort):
    meta, _ = golden_undistort
    cals = {"tmpl": "tmpl_auto", "syn": "syn_auto", "tiny": "tiny_auto", "wide": "wide_auto"}
    for key, want in meta["auto_zoom"].items():
        parts = key.split("_")
        cal = meta["cases"][cals[parts[0]]]["calibration"]
        n = int(parts[2][1:]) if len(parts) > 2 else 192
        got = g.auto_undistort_zoom(cal, float(parts[1]), n)
        assert abs(got - want) <= 2e-6 * want, (key, got, want)
    assert meta["auto_zoom"]["wide_190"] > 1.05       # the search really ran


def test_undistort_rejects_degenerate_focal():
    with pytest.raises(ValueError):
        g.undistort_map64(dict(width=8, height=8, f=0.0, cx=0, cy=0, k1=0, k2=0, k3=0, k4=0, p1=0, p2=0, b1=0, b2=0),
                          1.0, 190.0)


# --- v360 output=fisheye (preset fisheyeXY, PC:375-379) -------------------------------------------

def test_fisheye_output_is_equidistant_about_the_view_axis():
    w = h = 101
    hf, vf = g.fisheye_fov_from_dfov(180.0, w, h)
    assert abs(hf - 180.0 / math.sqrt(2.0)) < 1e-9 and hf == vf
    rays = g.camera_rays(w, h, hf, vf, 0.0, 0.0, projection="fisheye")
    assert np.allclose(np.linalg.norm(rays, axis=2), 1.0)
    c = w // 2
    assert np.allclose(rays[c, c], (0.0, 0.0, 1.0), atol=1e-12)
    # angle from the axis grows linearly with the distance from the centre: edge centre = h_fov / 2,
    # corner pixel centres = d_fov / 2 scaled by (w - 1) / w
    ang = np.degrees(np.arccos(np.clip(rays[..., 2], -1, 1)))
    assert abs(ang[c, w - 1] - hf / 2 * (w - 1) / w) < 1e-9
    assert abs(ang[0, 0] - 90.0 * (w - 1) / w) < 1e-9
    k = np.arange(c, w)
    assert np.allclose(np.diff(ang[c, k]), hf / w, atol=1e-9)
    # +u looks right, +v (down in the image) looks down
    assert rays[c, w - 1, 0] > 0 and rays[h - 1, c, 1] < 0
    # yaw / pitch move the axis exactly as for rectilinear views
    r2 = g.camera_rays(w, h, hf, vf, 90.0, 30.0, projection="fisheye")
    r1 = g.camera_rays(3, 3, 10.0, 10.0, 90.0, 30.0)
    assert np.allclose(r2[c, c], r1[1, 1], atol=1e-12)
    with pytest.raises(ValueError):
        g.camera_rays(4, 4, 90, 90, 0, 0, projection="nope")


# --- the reference's own functions (gs360_GUI.py:342-424), recorded by tests/golden/make_golden.py ----------

def _wrap_diff(a, b, period):
    d = a - b
    return d - period * np.rint(d / period)


def test_erp_map_equals_the_reference_gui_functions(golden_dir):
    """33 x 33 pixel centres of ten views (presets, seam, poles, odd angles): the oracle's halfpixel map is the
    reference's edge-origin pixel coordinate minus 0.5; longitude modulo the panorama width, and not compared
    where the ray points at a pole (there it is arbitrary)."""
    gold = np.load(golden_dir / "gui_geometry.npz")
    n = len(gold["uv"])
    for W, H in ((7680, 3840), (3840, 1920)):
        for k, (yaw, pitch, hf, vf) in enumerate(gold["views"]):
            mx, my = g.erp_map64(W, H, n, n, yaw, pitch, hf, vf, "halfpixel")
            ex, ey = gold["x_%d" % W][k] - 0.5, gold["y_%d" % W][k] - 0.5
            assert np.abs(my - ey).max() < 1e-7, (W, yaw, pitch)
            off_pole = np.abs(np.abs(gold["lat"][k]) - math.pi / 2) > 1e-6
            assert off_pole.mean() > 0.99
            assert np.abs(_wrap_diff(mx, ex, W))[off_pole].max() < 1e-6, (W, yaw, pitch)


@pytest.mark.gpu
@pytest.mark.parametrize("path", ["direct", "tiled"])
def test_device_coordinates_equal_the_reference_gui_functions(golden_dir, path):
    """The kernels' float64 coordinates against the same recorded reference values (bar: 1e-3 px)."""
    torch = pytest.importorskip("torch")
    import remap360
    gold = np.load(golden_dir / "gui_geometry.npz")
    n = len(gold["uv"])
    views = [remap360.PerspectiveView(y, p, hf, vf) for y, p, hf, vf in gold["views"]]
    for W, H in ((7680, 3840), (3840, 1920)):
        got = remap360.sample_coordinates(views, (n, n), erp_size=(W, H), convention="halfpixel", path=path)
        x64, y64 = got["x64"].cpu().numpy(), got["y64"].cpu().numpy()
        for k in range(len(views)):
            ex, ey = gold["x_%d" % W][k] - 0.5, gold["y_%d" % W][k] - 0.5
            # at a pole the row clamps and the longitude is arbitrary: compare away from it
            off_pole = np.abs(np.abs(gold["lat"][k]) - math.pi / 2) > 1e-3
            assert np.abs(np.clip(y64[k], 0, H - 1) - np.clip(ey, 0, H - 1)).max() < 1e-3
            assert np.abs(_wrap_diff(x64[k], ex, W))[off_pole].max() < 1e-3, (W, k)
