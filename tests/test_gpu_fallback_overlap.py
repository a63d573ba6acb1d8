"""GPU: the fallback tiles (float64 direct path on the tiles the plan rejects).

The fallback kernel projects a pixel once and samples it from four frames of the batch at a time; with
``R360_FALLBACK_OVERLAP=1`` (an experiment, off by default) it runs on a side stream beside the tiled kernel.  Neither
may change a single output value: every frame of a batch must be BIT-IDENTICAL to the same frame remapped alone, the
overlapped launch to the serial one, the fork / join must hold when one plan is launched from several host threads on
different streams at once, and inside a CUDA graph capture.
"""

import os
import threading

import pytest

torch = pytest.importorskip("torch")

pytestmark = pytest.mark.gpu

FOV = 104.2500326978036
POLE_VIEWS = [(0, 90), (0, -90), (40, 60), (-40, -60), (45, 30), (180, 0)]


@pytest.fixture(scope="module")
def r360():
    import remap360
    if not torch.cuda.is_available():
        pytest.fail("-m gpu tests need a CUDA device")
    return remap360


@pytest.fixture()
def overlap_env():
    saved = os.environ.get("R360_FALLBACK_OVERLAP")

    def set_(on):
        if on:
            os.environ["R360_FALLBACK_OVERLAP"] = "1"
        else:
            os.environ.pop("R360_FALLBACK_OVERLAP", None)
    yield set_
    if saved is None:
        os.environ.pop("R360_FALLBACK_OVERLAP", None)
    else:
        os.environ["R360_FALLBACK_OVERLAP"] = saved


def _rand(shape, dtype, seed):
    g = torch.Generator(device="cuda")
    g.manual_seed(seed)
    if dtype == torch.uint16:
        return torch.randint(0, 65536, shape, dtype=torch.int32, device="cuda", generator=g).to(torch.uint16)
    return torch.randint(0, 256, shape, dtype=torch.uint8, device="cuda", generator=g)


def _same(a, b):
    if a.dtype == torch.uint16:
        a, b = a.view(torch.int16), b.view(torch.int16)
    return bool(torch.equal(a, b))


@pytest.mark.parametrize("dtype", [torch.uint8, torch.uint16])
@pytest.mark.parametrize("interp", ["nearest", "linear", "cubic"])
@pytest.mark.parametrize("n_frames", [1, 3, 4, 9])
def test_batched_and_overlapped_fallback_equal_single_frames(r360, overlap_env, dtype, interp, n_frames):
    W, H, size = 4096, 2048, 480
    views = [r360.PerspectiveView(y, p, FOV, FOV) for y, p in POLE_VIEWS]
    src = _rand((n_frames, H, W, 3), dtype, 5 + n_frames)
    overlap_env(False)
    want = torch.stack([r360.remap_erp(src[k:k + 1], views, (size, size), interp=interp, path="tiled")[0] for k in range(n_frames)])
    from remap360 import api
    plan = list(api._PLAN_CACHE.values())[-1]
    assert plan.n_fallback > 0, "the pole views must leave tiles to the fallback kernel for this test to mean anything"
    assert plan.n_map_tiles > 0, "and tiles with a per-pixel map"
    got = r360.remap_erp(src, views, (size, size), interp=interp, path="tiled")
    assert _same(got, want)
    overlap_env(True)
    for _ in range(2):
        got = r360.remap_erp(src, views, (size, size), interp=interp, path="tiled")
        assert _same(got, want)


@pytest.mark.parametrize("dtype", [torch.uint8, torch.uint16])
@pytest.mark.parametrize("interp", ["nearest", "linear", "cubic", "lanczos4"])
def test_per_pixel_map_tiles_equal_the_direct_path(r360, dtype, interp):
    """Tiles next to a pole carry an explicit float64 map and run through the tiled kernel's staging and samplers, in
    a second pass with the largest ring when their patch is large; with ``R360_COORD_TILES=0`` /
    ``R360_LARGE_PATCH_PASS=0`` the same tiles go to the float64 fallback kernel.  Same coordinates (to ~1e-12 px), same
    arithmetic: the outputs must agree everywhere (a differing pixel would need a coordinate within 1e-12 px of a
    1/32-px rounding tie)."""
    from remap360 import api
    if interp == "lanczos4" and dtype != torch.uint8:
        pytest.skip("lanczos4 is covered for 8-bit sources")
    W, H, size = 4096, 2048, 480
    views = [r360.PerspectiveView(y, p, FOV, FOV) for y, p in POLE_VIEWS + [(179.0, 75.0), (-178.0, -80.0)]]
    src = _rand((2, H, W, 3), dtype, 41)
    saved = {k: os.environ.get(k) for k in ("R360_COORD_TILES", "R360_LARGE_PATCH_PASS")}
    try:
        os.environ["R360_COORD_TILES"] = "0"
        os.environ["R360_LARGE_PATCH_PASS"] = "0"
        api.clear_plan_cache()
        want = r360.remap_erp(src, views, (size, size), interp=interp, path="tiled").clone()
        plan0 = list(api._PLAN_CACHE.values())[-1]
        os.environ.pop("R360_LARGE_PATCH_PASS")
        api.clear_plan_cache()
        got_large = r360.remap_erp(src, views, (size, size), interp=interp, path="tiled").clone()
        plan_large = list(api._PLAN_CACHE.values())[-1]
        os.environ.pop("R360_COORD_TILES")
        api.clear_plan_cache()
        got = r360.remap_erp(src, views, (size, size), interp=interp, path="tiled")
        plan1 = list(api._PLAN_CACHE.values())[-1]
    finally:
        for k, v in saved.items():
            if v is None:
                os.environ.pop(k, None)
            else:
                os.environ[k] = v
        api.clear_plan_cache()
    assert plan0.n_map_tiles == 0 and plan_large.n_map_tiles == 0 and plan1.n_map_tiles > 0
    assert plan_large.n_fallback <= plan0.n_fallback          # polynomial tiles with large patches left the list
    assert plan1.n_fallback < plan_large.n_fallback
    # (1) the large-patch pass moves POLYNOMIAL tiles from the direct path to the tiled one: same results up to the
    # 1/32-px bin flips that ~1e-5 px of coordinate difference causes on a few pixels in a million (include/remap360.h)
    a, b = got_large.to(torch.int32), want.to(torch.int32)
    flips = int((a != b).any(dim=-1).sum())
    assert flips <= 2e-5 * a[..., 0].numel(), "large-patch pass: %d pixels differ from the direct path" % flips
    # (2) tiles with a per-pixel map sample at the direct path's own coordinates: 8-bit results are identical
    # (integer arithmetic on both sides); 16-bit ones may differ by single LSBs -- the packed sampler of the tiled
    # kernel and the generic one of the fallback kernel both round like cv2 (each is within the parity bar of the
    # oracle, tests/test_gpu_parity.py) but associate the float32 products differently
    a, b = got.to(torch.int32), got_large.to(torch.int32)
    differ, worst = int((a != b).sum()), int((a - b).abs().max())
    if dtype == torch.uint8:
        assert differ == 0, "per-pixel maps: %d samples differ (worst %d)" % (differ, worst)
    else:
        assert worst <= 1 and differ <= 1e-4 * a.numel(), "per-pixel maps: %d of %d samples differ (worst %d)" % (differ, a.numel(), worst)


@pytest.mark.parametrize("n_pairs", [1, 5])
def test_batched_fallback_on_fisheye_pairs(r360, n_pairs):
    """Constant-border taps and the invalid-pixel fill of the dual-fisheye projection, four pairs at a time."""
    import json
    import pathlib
    meta = json.loads((pathlib.Path(__file__).parent / "golden" / "dualfisheye.json").read_text())
    cal = meta["sensors"]["0"]
    scaled = {k: cal[k] / 4.0 if k in ("f", "cx", "cy") else cal[k] for k in ("f", "cx", "cy", "k1", "k2", "k3", "k4", "p1", "p2", "b1", "b2")}
    calib = r360.FisheyeCalibration(**scaled, width=int(cal["width"]) // 4, height=int(cal["height"]) // 4)
    views = [r360.PerspectiveView(y, p, 100.0, 100.0, src_slot=s) for y, p, s in ((0, 0, 0), (80, 10, 0), (-85, -20, 1), (30, 70, 1))]
    H = W = int(cal["height"]) // 4
    src = _rand((n_pairs, 2, H, W, 3), torch.uint8, 31 + n_pairs)
    want = torch.stack([r360.remap_fisheye(src[k:k + 1], [calib, calib], views, (352, 352), interp="cubic", path="tiled")[0]
                        for k in range(n_pairs)])
    from remap360 import api
    assert list(api._PLAN_CACHE.values())[-1].n_fallback > 0
    got = r360.remap_fisheye(src, [calib, calib], views, (352, 352), interp="cubic", path="tiled")
    assert _same(got, want)


def test_one_plan_from_two_host_threads_on_two_streams(r360, overlap_env):
    """The fork / join events are made per call: two threads launching the same cached plan at once must each get
    their own frames' result (a shared event would let one thread's fallback tiles start before its input exists)."""
    overlap_env(True)
    W, H, size = 4096, 2048, 480
    views = [r360.PerspectiveView(y, p, FOV, FOV) for y, p in POLE_VIEWS]
    rounds = 12
    srcs = [[_rand((2, H, W, 3), torch.uint8, 100 * t + k) for k in range(rounds)] for t in range(2)]
    want = [[r360.remap_erp(s, views, (size, size), interp="cubic", path="tiled").clone() for s in row] for row in srcs]
    torch.cuda.synchronize()
    errors = []

    def work(t):
        try:
            stream = torch.cuda.Stream()
            with torch.cuda.stream(stream):
                for k in range(rounds):
                    # the input is produced on this thread's stream right before the launch: the fallback tiles must
                    # wait for it through the fork event
                    fresh = torch.empty_like(srcs[t][k])
                    fresh.copy_(srcs[t][k], non_blocking=True)
                    got = r360.remap_erp(fresh, views, (size, size), interp="cubic", path="tiled", stream=stream)
                    stream.synchronize()
                    if not _same(got, want[t][k]):
                        errors.append((t, k))
        except Exception as exc:                                   # surfaced by the assertion below
            errors.append((t, repr(exc)))

    threads = [threading.Thread(target=work, args=(t,)) for t in range(2)]
    for th in threads:
        th.start()
    for th in threads:
        th.join()
    assert not errors, errors


def test_fork_and_join_inside_a_graph_capture(r360, overlap_env):
    overlap_env(True)
    W, H, size = 2048, 1024, 320
    views = [r360.PerspectiveView(y, p, FOV, FOV) for y, p in POLE_VIEWS]
    src = _rand((2, H, W, 3), torch.uint8, 77)
    want = r360.remap_erp(src, views, (size, size), interp="linear", path="tiled").clone()
    out = torch.zeros_like(want)
    stream = torch.cuda.Stream()
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.stream(stream):
        r360.remap_erp(src, views, (size, size), interp="linear", path="tiled", out=out, stream=stream)   # plan + descriptors built
        stream.synchronize()
        out.zero_()
        with torch.cuda.graph(graph, stream=stream):
            r360.remap_erp(src, views, (size, size), interp="linear", path="tiled", out=out, stream=stream)
    graph.replay()
    torch.cuda.synchronize()
    assert _same(out, want)
    src2 = _rand((2, H, W, 3), torch.uint8, 78)
    want2 = r360.remap_erp(src2, views, (size, size), interp="linear", path="tiled").clone()
    src.copy_(src2)
    graph.replay()
    torch.cuda.synchronize()
    assert _same(out, want2)
