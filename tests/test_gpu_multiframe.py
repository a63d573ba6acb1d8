"""GPU: frame blocks of the tiled kernel.

A batch (video frames, DF pairs) shares its map between frames, so the tiled kernel samples several frames per
work item with one set of coordinates / tap addresses / weights.  The arithmetic per frame is unchanged: whatever
the block size (``R360_FRAMES`` = 1, 2, 4) and however a batch splits into full blocks, left-over frames and
one-frame slots of patches too large for the ring, every frame must come out BIT-IDENTICAL to the same frame
remapped alone.  Parity with the oracle is then inherited from tests/test_gpu_parity.py (single frames).
"""

import os

import numpy as np
import pytest

torch = pytest.importorskip("torch")

pytestmark = pytest.mark.gpu

FOV = 104.2500326978036
SPECS = [(0, 0), (45, 30), (180, 0), (179.9, 0), (0, 90), (0, -90), (-70, -60), (123.4, 89.0)]


@pytest.fixture(scope="module")
def r360():
    import remap360
    if not torch.cuda.is_available():
        pytest.fail("-m gpu tests need a CUDA device")
    return remap360


@pytest.fixture()
def frames_env():
    """Sets R360_FRAMES / R360_MULTI_PCT for the duration of a test (read by the library at every launch)."""
    saved = {k: os.environ.get(k) for k in ("R360_FRAMES", "R360_MULTI_PCT")}

    def set_(frames, pct=None):
        os.environ["R360_FRAMES"] = str(frames)
        if pct is None:
            os.environ.pop("R360_MULTI_PCT", None)
        else:
            os.environ["R360_MULTI_PCT"] = str(pct)
    yield set_
    for k, v in saved.items():
        if v is None:
            os.environ.pop(k, None)
        else:
            os.environ[k] = v


def _rand(shape, dtype, seed):
    g = torch.Generator(device="cuda")
    g.manual_seed(seed)
    if dtype == torch.uint8:
        return torch.randint(0, 256, shape, dtype=torch.uint8, device="cuda", generator=g)
    if dtype == torch.uint16:
        return torch.randint(0, 65536, shape, dtype=torch.int32, device="cuda", generator=g).to(torch.uint16)
    t = torch.rand(shape, dtype=torch.float32, device="cuda", generator=g)
    return t.to(dtype)


def _same(a, b):
    if a.dtype == torch.uint16:
        a, b = a.view(torch.int16), b.view(torch.int16)
    if a.dtype in (torch.float16, torch.float32):
        return bool(torch.equal(a.view(torch.int16 if a.dtype == torch.float16 else torch.int32),
                                b.view(torch.int16 if b.dtype == torch.float16 else torch.int32)))
    return bool(torch.equal(a, b))


@pytest.mark.parametrize("fr", [2, 4])
@pytest.mark.parametrize("n_frames", [2, 3, 4, 5, 9])
@pytest.mark.parametrize("interp", ["linear", "cubic"])
def test_erp_u8_blocks_equal_single_frames(r360, frames_env, fr, n_frames, interp):
    W, H, size = 2048, 1024, 224            # 224 = 7 tiles: the last column / row of tiles is full, sizes below are ragged
    views = [r360.PerspectiveView(y, p, FOV, FOV) for y, p in SPECS]
    src = _rand((n_frames, H, W, 3), torch.uint8, 11 + n_frames)
    frames_env(1)
    want = torch.stack([r360.remap_erp(src[k:k + 1], views, (size, size), interp=interp, path="tiled")[0]
                        for k in range(n_frames)])
    for pct in (50, 100, 10):               # 10 %: nearly every item splits into one-frame slots
        frames_env(fr, pct)
        got = r360.remap_erp(src, views, (size, size), interp=interp, path="tiled")
        assert _same(got, want), (fr, n_frames, interp, pct)


@pytest.mark.parametrize("fr", [2, 4])
@pytest.mark.parametrize("dtype,out_dtype", [(torch.uint8, None), (torch.uint16, None), (torch.uint16, torch.float16),
                                             (torch.float16, None), (torch.float32, None)])
@pytest.mark.parametrize("interp", ["nearest", "linear", "cubic", "lanczos4"])
@pytest.mark.parametrize("channels", [1, 3, 4])
def test_erp_every_sampler_in_blocks(r360, frames_env, fr, dtype, out_dtype, interp, channels):
    W, H, size = 1024, 512, 100             # ragged: 100 = 3 tiles + 4 pixels
    views = [r360.PerspectiveView(y, p, FOV, FOV) for y, p in SPECS[:6]]
    src = _rand((5, H, W, channels), dtype, 5)
    frames_env(1)
    want = torch.stack([r360.remap_erp(src[k:k + 1], views, (size, size), interp=interp, path="tiled",
                                       out_dtype=out_dtype)[0] for k in range(5)])
    frames_env(fr)
    got = r360.remap_erp(src, views, (size, size), interp=interp, path="tiled", out_dtype=out_dtype)
    assert _same(got, want)


@pytest.mark.parametrize("fr", [2, 4])
@pytest.mark.parametrize("interp", ["nearest", "linear", "cubic"])
def test_dual_fisheye_pairs_in_blocks(r360, frames_env, fr, interp):
    """Two lens images per group: the frames of a block are every second image of the batch."""
    cal = [r360.FisheyeCalibration(width=960, height=960, f=300.0, cx=3.0, cy=-2.0, k1=0.02, k2=-0.004, p1=1e-4, b1=0.3),
           r360.FisheyeCalibration(width=960, height=960, f=301.0, cx=-4.0, cy=1.0, k1=0.018, k3=1e-4, p2=-1e-4, b2=0.1)]
    views = [r360.PerspectiveView(y, p, 90.0, 90.0, src_slot=s) for y, p, s in
             [(0, 0, 0), (40, 20, 0), (-60, -35, 1), (0, 0, 1), (85, 0, 0), (0, 80, 1)]]
    pairs = _rand((5, 2, 960, 960, 3), torch.uint8, 3)
    frames_env(1)
    want = torch.stack([r360.remap_fisheye(pairs[k:k + 1], cal, views, (250, 250), interp=interp, path="tiled")[0]
                        for k in range(5)])
    frames_env(fr)
    got = r360.remap_fisheye(pairs, cal, views, (250, 250), interp=interp, path="tiled")
    assert _same(got.contiguous(), want.contiguous())


@pytest.mark.parametrize("interp", ["linear", "cubic"])
def test_full_size_blocks_equal_single_frames(r360, frames_env, interp):
    """BASELINE config 2 shapes: 8K frames, 1600 x 1600 views incl. seam and pole views."""
    views = [r360.PerspectiveView(y, p, FOV, FOV) for y, p in [(0, 0), (45, 30), (180, 0), (0, 90)]]
    src = _rand((4, 3840, 7680, 3), torch.uint8, 77)
    frames_env(1)
    want = r360.remap_erp(src, views, (1600, 1600), interp=interp, path="tiled")
    for fr in (2, 4):
        frames_env(fr)
        got = r360.remap_erp(src, views, (1600, 1600), interp=interp, path="tiled")
        assert _same(got, want), fr
