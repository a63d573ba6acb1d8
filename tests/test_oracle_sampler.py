"""The NumPy model of cv2.remap (oracle/sampler.py) against (a) fixtures recorded
from cv2.remap by tests/golden/make_golden.py and (b) cv2.remap itself, live.

Bar: bit-exact for every dtype / interpolation / border / channel count."""

import numpy as np
import pytest

from oracle import sampler

DTYPES = ("uint8", "uint16", "float32")


@pytest.mark.parametrize("dt", DTYPES)
@pytest.mark.parametrize("ch", (1, 3))
@pytest.mark.parametrize("interp", sampler.INTERPS)
@pytest.mark.parametrize("bv", (0, 37))
def test_model_matches_recorded_cv2(golden_cv2, dt, ch, interp, bv):
    src = golden_cv2["src_%s_c%d" % (dt, ch)]
    want = golden_cv2["out_%s_c%d_%s_b%d" % (dt, ch, interp, bv)]
    got = sampler.sample(src, golden_cv2["map_x"], golden_cv2["map_y"], interp, "constant", bv)
    assert got.dtype == want.dtype and got.shape == want.shape
    assert np.array_equal(got, want)


def _random_case(rng, dt, ch, h=90, w=130, n=160, margin=5.0):
    if dt == "float32":
        src = rng.random((h, w, ch), dtype=np.float32)
    else:
        src = rng.integers(0, np.iinfo(dt).max + 1, (h, w, ch)).astype(dt)
    if ch == 1:
        src = src[..., 0]
    mx = (rng.random((n, n)) * (w + 2 * margin) - margin).astype(np.float32)
    my = (rng.random((n, n)) * (h + 2 * margin) - margin).astype(np.float32)
    return src, mx, my


@pytest.mark.parametrize("dt", DTYPES)
@pytest.mark.parametrize("ch", (1, 3, 4))
@pytest.mark.parametrize("interp", sampler.INTERPS)
def test_model_matches_live_cv2_constant_border(dt, ch, interp):
    pytest.importorskip("cv2")
    rng = np.random.default_rng(hash((dt, ch, interp)) % (2 ** 32))
    src, mx, my = _random_case(rng, dt, ch)
    for bv in (0, 200):
        want = sampler.sample_cv2(src, mx, my, interp, "constant", bv)
        got = sampler.sample(src, mx, my, interp, "constant", bv)
        assert np.array_equal(got, want)


@pytest.mark.parametrize("dt", DTYPES)
@pytest.mark.parametrize("interp", sampler.INTERPS)
def test_model_matches_live_cv2_panorama_border(dt, interp):
    """x wraps, y clamps: cv2 sees a source padded by wrap / replicate."""
    pytest.importorskip("cv2")
    rng = np.random.default_rng(7)
    src, _, _ = _random_case(rng, dt, 3)
    h, w = src.shape[:2]
    n = 200
    mx = (rng.random((n, n)) * w - 0.5).astype(np.float32)
    my = (rng.random((n, n)) * h - 0.5).astype(np.float32)
    mx[0, :4] = (-0.5, -0.49, w - 0.51, w - 1.0)          # seam taps on both sides
    my[1, :4] = (-0.5, 0.0, h - 1.0, h - 0.5)             # pole rows
    want = sampler.sample_cv2(src, mx, my, interp, "erp")
    got = sampler.sample(src, mx, my, interp, "erp")
    assert np.array_equal(got, want)


def test_fraction_quantisation_is_round_half_even_at_one_64th():
    # 1/64 is the midpoint between fraction bins 0 and 1 -> rounds to the even bin 0;
    # 3/64 is the midpoint between 1 and 2 -> rounds to 2.
    ix, fx = sampler.quantise(np.array([10 + 1 / 64, 10 + 3 / 64, -0.5, -1 / 64], dtype=np.float32))
    assert fx.tolist() == [0, 2, 16, 0]
    assert ix.tolist() == [10, 10, -1, 0]


def test_cubic_table_sums_and_known_entries():
    t1, tab, itab = sampler.tables("cubic")
    assert itab.reshape(32, 32, 16).sum(axis=2).tolist() == [[32768] * 32] * 32
    # fraction 0: identity tap (saturated 32767 + the folded unit)
    assert itab[0, 0, 1, 1] == 32767 and itab[0, 0].sum() == 32768
    # a = -0.75 at t = 0.5: (-3/32, 19/32, 19/32, -3/32)
    assert np.allclose(t1[16], [-0.09375, 0.59375, 0.59375, -0.09375], atol=1e-7)


def test_float16_output_paths():
    rng = np.random.default_rng(3)
    src16 = rng.integers(0, 65536, (32, 48, 3)).astype(np.uint16)
    mx = (rng.random((20, 20)) * 40 + 2).astype(np.float32)
    my = (rng.random((20, 20)) * 24 + 2).astype(np.float32)
    as_u16 = sampler.sample(src16, mx, my, "cubic", "erp")
    as_f16 = sampler.sample(src16, mx, my, "cubic", "erp", out_dtype=np.float16)
    assert as_f16.dtype == np.float16
    # half precision of value/65535, unclamped -> compare away from saturation
    mid = (as_u16 > 0) & (as_u16 < 65535)
    assert np.abs(as_f16.astype(np.float64) * 65535.0 - as_u16)[mid].max() <= 65535 * 2.0 ** -11 + 1
    srch = rng.random((32, 48, 3)).astype(np.float16)
    out = sampler.sample(srch, mx, my, "linear", "erp")
    ref = sampler.sample(srch.astype(np.float32), mx, my, "linear", "erp").astype(np.float16)
    assert out.dtype == np.float16 and np.array_equal(out, ref)


def test_invalid_fill():
    img = np.full((4, 5, 3), 9, np.uint8)
    valid = np.ones((4, 5), bool)
    valid[1, 2] = False
    out = sampler.apply_invalid_fill(img, valid, 0)
    assert out[1, 2].tolist() == [0, 0, 0] and out.sum() == 9 * 3 * 19
