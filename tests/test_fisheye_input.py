"""Fisheye frames -> perspective view (gs360_Video2Frames.py --fisheye-perspective, V2F:467-493).

CPU part: flags -> FOVs -> filter string against values recorded by running the reference's own code
(tests/golden/v2f_fisheye.json), the lens record against an independent v360-style statement of the
mapping.  GPU part (``-m gpu``): both device paths against the float64 oracle and cv2's arithmetic."""

import json

import numpy as np
import pytest

from oracle import geometry as geo
from oracle import sampler

COORD_TOL_PX = 1e-3
PIXEL_OK_FRACTION = 0.999


@pytest.fixture(scope="module")
def golden_v2f(golden_dir):
    return json.loads((golden_dir / "v2f_fisheye.json").read_text())["cases"]


def test_filter_string_and_fov_equal_reference(golden_v2f):
    pytest.importorskip("torch")
    from remap360 import fisheye_input as fi
    assert len(golden_v2f) == 15
    for c in golden_v2f:
        assert fi.v360_filter(c["projection"], c["input_fov"], c["focal_mm"], c["size"]) == c["filter"]
        hf, vf = fi.view_fov_deg(c["focal_mm"], c["size"])
        assert hf == c["hfov_deg"] and vf == c["vfov_deg"]
    assert fi.FISHEYE_SENSOR_WIDTH_MM == 36.0 and fi.FISHEYE_INPUT_FOV_DEG == 190.0


def test_validation_messages():
    pytest.importorskip("torch")
    from remap360 import fisheye_input as fi
    assert fi.validate_args(8.0, 1600, 190.0) is None
    assert fi.validate_args(0.0, 1600, 190.0).startswith("Focal length must be greater than zero")
    assert fi.validate_args(8.0, 0, 190.0).startswith("Output size must be greater than zero")
    assert fi.validate_args(8.0, 1600, -1.0).startswith("Input fisheye FOV must be greater than zero")


@pytest.mark.parametrize("projection", ["equidistant", "equisolid"])
@pytest.mark.parametrize("convention", ["halfpixel", "v360"])
def test_ideal_lens_record_equals_v360_style_mapping(projection, convention):
    """f / b1 / cx / cy of the lens record reproduce (uf, vf) = r(theta) / r(fov / 2) scaled to the image."""
    pytest.importorskip("torch")
    from remap360 import fisheye_input as fi
    cal = fi.ideal_calibration(3840, 2880, projection, 190.0, 150.0, convention)
    ocal = geo.ideal_fisheye_calib(3840, 2880, projection, 190.0, 150.0, convention)
    for k in ("f", "b1", "cx", "cy"):
        assert getattr(cal, k) == ocal[k]
    assert cal.model == projection and cal.lens_fov_deg == 360.0
    mx, my, valid = geo.fisheye_map64(ocal, 12.0, -7.0, 100.0, 80.0, 160, 120, 360.0)
    ux, uy = geo.v360_fisheye_input_map64(3840, 2880, projection, 190.0, 150.0, 100.0, 80.0, 160, 120, 12.0, -7.0,
                                          convention)
    assert np.abs(mx - ux).max() < 1e-8 and np.abs(my - uy).max() < 1e-8
    # the edge of the field lands on the edge of the image (halfpixel) / the last pixel centre (v360)
    half = np.radians(95.0)
    r_edge = half if projection == "equidistant" else 2.0 * np.sin(half / 2.0)
    x_edge = 3840 * 0.5 + ocal["cx"] + r_edge * (ocal["f"] + ocal["b1"])
    assert abs(x_edge - (3839.5 if convention == "halfpixel" else 3839.0)) < 1e-9


# ---------------------------------------------------------------------------------------------------------
# GPU
# ---------------------------------------------------------------------------------------------------------

def _lsb(got, want):
    d = np.abs(got.astype(np.int64) - want.astype(np.int64))
    return float((d == 0).mean()), float((d <= 1).mean())


@pytest.mark.gpu
@pytest.mark.parametrize("path", ["direct", "tiled"])
@pytest.mark.parametrize("projection", ["equidistant", "equisolid"])
def test_coordinates_against_float64_oracle(path, projection):
    torch = pytest.importorskip("torch")
    import remap360
    from remap360 import fisheye_input as fi
    W, H, size = 3840, 3840, 512
    for convention, fov, focal, yaw, pitch in (("halfpixel", 190.0, 8.0, 0.0, 0.0), ("v360", 220.0, 14.0, 25.0, -10.0),
                                               ("halfpixel", 150.0, 6.0, 0.0, 0.0)):
        cal = fi.ideal_calibration(W, H, projection, fov, fov, convention)
        hf, vf = fi.view_fov_deg(focal, size)
        got = remap360.sample_coordinates([remap360.PerspectiveView(yaw, pitch, hf, vf)], (size, size), calibs=[cal],
                                          path=path)
        ocal = geo.ideal_fisheye_calib(W, H, projection, fov, fov, convention)
        mx, my, valid = geo.fisheye_map64(ocal, yaw, pitch, hf, vf, size, size, 360.0)
        ok = got["valid"][0].cpu().numpy().astype(bool)
        # validity may differ only where the coordinate is within rounding of the image border
        edge = (np.abs(mx) < 1e-6) | (np.abs(mx - (W - 1)) < 1e-6) | (np.abs(my) < 1e-6) | (np.abs(my - (H - 1)) < 1e-6)
        assert np.all((ok == valid) | edge)
        both = ok & valid
        assert both.mean() > 0.3
        ex = np.abs(got["x64"][0].cpu().numpy() - mx)[both].max()
        ey = np.abs(got["y64"][0].cpu().numpy() - my)[both].max()
        assert max(ex, ey) < COORD_TOL_PX, (convention, fov, ex, ey)
        assert max(ex, ey) < 5e-5            # how far inside the bar the kernels are


@pytest.mark.gpu
@pytest.mark.parametrize("path", ["direct", "tiled"])
@pytest.mark.parametrize("projection", ["equidistant", "equisolid"])
@pytest.mark.parametrize("interp", ["linear", "cubic"])
def test_pixels_against_cv2_arithmetic(path, projection, interp):
    """V2F's defaults (190 degree input, 8 mm, cubic) on a noise frame; a 150 degree lens so that part of the
    view leaves the image and is filled with black."""
    torch = pytest.importorskip("torch")
    from remap360 import fisheye_input as fi
    rng = np.random.default_rng(7)
    W, H, size = 1280, 960, 384
    frames = rng.integers(0, 256, (2, H, W, 3), dtype=np.uint8)
    for fov, focal, pitch in ((190.0, 8.0, 0.0), (150.0, 5.0, 12.0)):
        got = fi.fisheye_to_perspective(torch.from_numpy(frames).cuda(), projection=projection, input_fov_deg=fov,
                                        focal_mm=focal, size_px=size, pitch_deg=pitch, interp=interp,
                                        path=path).cpu().numpy()
        hf, vf = fi.view_fov_deg(focal, size)
        ocal = geo.ideal_fisheye_calib(W, H, projection, fov, fov, "halfpixel")
        mx, my, valid = geo.fisheye_map64(ocal, 0.0, pitch, hf, vf, size, size, 360.0)
        if fov == 150.0:
            assert 0.05 < 1.0 - valid.mean() < 0.95
        for b in range(2):
            want = sampler.apply_invalid_fill(sampler.sample(frames[b], mx, my, interp, "constant", 0), valid, 0)
            exact, within1 = _lsb(got[b], want)
            assert within1 >= PIXEL_OK_FRACTION, (fov, b, exact, within1)
            assert exact >= 0.998
