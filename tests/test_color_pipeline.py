"""Input colour pipeline (.cube LUT + Rec.709 -> sRGB), DF:494-725: the oracle against outputs recorded from
the reference, the host-side .cube parser, and (``-m gpu``) the CUDA kernel against the oracle."""

import pathlib

import numpy as np
import pytest

from oracle import color as ocolor

GOLDEN = pathlib.Path(__file__).resolve().parent / "golden"
CASES = [(dt, ch, lut, space) for dt, ch in (("uint8", 3), ("uint16", 3), ("uint8", 4))
         for lut in ("s5", "s17") for space in ("passthrough", "srgb")]


@pytest.fixture(scope="module")
def gold():
    return np.load(GOLDEN / "color_pipeline.npz")


def _lut(gold, name):
    return gold["cube_%s_table" % name], gold["cube_%s_min" % name], gold["cube_%s_max" % name]


@pytest.mark.parametrize("dt,ch,lut,space", CASES)
def test_oracle_matches_reference_outputs(gold, dt, ch, lut, space):
    table, dmin, dmax = _lut(gold, lut)
    img = gold["img_%s_c%d_%s" % (dt, ch, lut)]
    want = gold["out_%s_c%d_%s_%s" % (dt, ch, lut, space)]
    got = ocolor.apply_pipeline(img, table, dmin, dmax, space)
    assert got.dtype == want.dtype and np.array_equal(got, want)


def test_transfer_functions_match_reference(gold):
    got = ocolor.rec709_to_srgb(gold["rec709_to_srgb_in"])
    assert np.array_equal(got, gold["rec709_to_srgb_out"])
    assert got[0] == 0.0 and abs(float(got[-1]) - 1.0) < 1e-6


def test_cube_parser_matches_reference(gold, tmp_path):
    torch = pytest.importorskip("torch")  # noqa: F841  (remap360 imports torch)
    from remap360 import color
    for name in ("s5", "s17"):
        path = tmp_path / (name + ".cube")
        path.write_text(str(gold["cube_%s_text" % name]))
        lut = color.load_cube_lut(path)
        table, dmin, dmax = _lut(gold, name)
        assert lut.size == table.shape[0] and np.array_equal(lut.table, table)
        assert np.array_equal(np.float32(lut.domain_min), dmin) and np.array_equal(np.float32(lut.domain_max), dmax)
    (tmp_path / "bad.cube").write_text("LUT_3D_SIZE 2\n0 0 0\n")
    with pytest.raises(ValueError, match="row count mismatch"):
        color.load_cube_lut(tmp_path / "bad.cube")
    (tmp_path / "nosize.cube").write_text("0 0 0\n")
    with pytest.raises(ValueError, match="LUT_3D_SIZE is missing"):
        color.load_cube_lut(tmp_path / "nosize.cube")
    with pytest.raises(FileNotFoundError):
        color.load_cube_lut(tmp_path / "absent.cube")
    assert color.normalize_lut_output_color_space("native") == "passthrough"
    with pytest.raises(ValueError):
        color.normalize_lut_output_color_space("p3")


@pytest.mark.gpu
@pytest.mark.parametrize("dt,ch,lut,space", CASES)
def test_cuda_kernel_matches_reference_outputs(gold, dt, ch, lut, space):
    torch = pytest.importorskip("torch")
    from remap360 import color
    table, dmin, dmax = _lut(gold, lut)
    cube = color.CubeLUT(int(table.shape[0]), table, tuple(float(v) for v in dmin), tuple(float(v) for v in dmax))
    img = gold["img_%s_c%d_%s" % (dt, ch, lut)]
    want = gold["out_%s_c%d_%s_%s" % (dt, ch, lut, space)]
    dev = torch.from_numpy(img.view(np.int16) if dt == "uint16" else img).cuda()
    dev = dev.view(torch.uint16) if dt == "uint16" else dev
    got = color.apply_input_color_pipeline(dev[None], cube, space)[0]
    got = got.view(torch.int16).cpu().numpy().view(np.uint16) if dt == "uint16" else got.cpu().numpy()
    diff = np.abs(got.astype(np.int64) - want.astype(np.int64))
    # float32 powf on the device vs NumPy's: identical up to results that sit on a rounding boundary
    assert diff.max() <= 1 and (diff == 0).mean() >= (0.9999 if dt == "uint8" else 0.995), (diff.max(), (diff == 0).mean())
    if space == "passthrough":
        assert diff.max() == 0                                  # no transcendental involved: bit-exact


@pytest.mark.gpu
def test_cuda_kernel_in_place_rgb_order_and_errors(gold):
    torch = pytest.importorskip("torch")
    import remap360
    from remap360 import color
    table, dmin, dmax = _lut(gold, "s5")
    cube = color.CubeLUT(5, table, tuple(float(v) for v in dmin), tuple(float(v) for v in dmax))
    img = gold["img_uint8_c3_s5"]
    want = ocolor.apply_pipeline(np.ascontiguousarray(img[..., ::-1]), table, dmin, dmax, "passthrough", "rgb")
    dev = torch.from_numpy(np.ascontiguousarray(img[..., ::-1])).cuda()[None, None]     # [1, 1, H, W, C] batch
    out = color.apply_input_color_pipeline(dev, cube, "passthrough", channel_order="rgb", out=dev)
    assert out.data_ptr() == dev.data_ptr() and np.array_equal(out[0, 0].cpu().numpy(), want)
    assert color.apply_input_color_pipeline(dev, None) is dev
    with pytest.raises(ValueError, match="at least 3-channel"):
        color.apply_input_color_pipeline(torch.zeros((1, 4, 4, 1), dtype=torch.uint8, device="cuda"), cube)
    with pytest.raises(remap360.Remap360Error):
        color.apply_input_color_pipeline(torch.zeros((1, 4, 4, 3), dtype=torch.float16, device="cuda"), cube)
