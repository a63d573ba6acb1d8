"""Diagnose a failing seed of tests/test_gpu_fuzz.py::test_random_erp_configurations: for every pixel where the direct
path differs from the oracle, how far the float64 coordinate is from a 1/32-px rounding boundary."""
import sys, pathlib
import numpy as np, torch
ROOT = pathlib.Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT / "360cam-pgm-3dgs-tools_b200")); sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "tests"))
import remap360 as r360
from oracle import geometry as geo, sampler
import test_gpu_fuzz as tf

for seed in [int(a) for a in sys.argv[1:]]:
    rng = np.random.default_rng(1000 + seed)
    channels = int(rng.choice([1, 3, 3, 3, 4]))
    dtype = [np.uint8, np.uint8, np.uint16, np.float32, np.float16][int(rng.integers(0, 5))]
    W = int(rng.choice([96, 250, 512, 777, 1024, 2048]))
    H = max(8, W // 2 + int(rng.integers(-3, 4)))
    ow, oh = int(rng.integers(1, 200)), int(rng.integers(1, 200))
    interp = ["nearest", "linear", "cubic", "lanczos4"][int(rng.integers(0, 4))]
    convention = "halfpixel" if rng.random() < 0.8 else "v360"
    if dtype in (np.float32, np.float16):
        src = rng.random((H, W, channels), dtype=np.float32).astype(dtype)
    else:
        src = rng.integers(0, np.iinfo(dtype).max + 1, (H, W, channels)).astype(dtype)
    views = []
    for _ in range(int(rng.integers(1, 5))):
        proj = "fisheye" if rng.random() < 0.2 else "rectilinear"
        fov_hi = 300.0 if proj == "fisheye" else 175.0
        views.append(r360.PerspectiveView(float(rng.uniform(-200, 200)), float(rng.choice([rng.uniform(-90, 90), 90.0, -90.0, 0.0])),
                                          float(rng.uniform(5, fov_hi)), float(rng.uniform(5, fov_hi)),
                                          roll_deg=float(rng.choice([0.0, rng.uniform(-180, 180)])), projection=proj))
    print("seed", seed, dtype.__name__, interp, convention, "src", W, H, channels, "out", ow, oh)
    dev = tf._cuda(src)[None]
    direct = tf._host(r360.remap_erp(dev, views, (ow, oh), interp=interp, convention=convention, path="direct"))[0]
    coords = r360.sample_coordinates(views, (ow, oh), erp_size=(W, H), convention=convention, path="direct")
    for k, v in enumerate(views):
        mx, my = geo.erp_map64(W, H, ow, oh, v.yaw_deg, v.pitch_deg, v.hfov_deg, v.vfov_deg, convention, v.roll_deg, v.projection)
        want = sampler.sample(src, mx, my, interp, "erp")
        if dtype in (np.float32, np.float16):
            bad = np.abs(direct[k].astype(np.float64) - want.astype(np.float64)).max(axis=-1) > 1e-3
        else:
            bad = np.abs(direct[k].astype(np.int64) - want.astype(np.int64)).max(axis=-1) > 1
        x64, y64 = coords["x64"][k].cpu().numpy(), coords["y64"][k].cpu().numpy()
        print(" view", k, v, "bad", int(bad.sum()), "of", bad.size, "coord err", float(np.abs(np.mod(x64 - mx + 0.5 * W, W) - 0.5 * W).max()),
              float(np.abs(y64 - my).max()))
        for (j, i) in list(zip(*np.nonzero(bad)))[:6]:
            fx = np.float32(mx[j, i]) * 32.0; fy = np.float32(my[j, i]) * 32.0
            print("   px", j, i, "oracle xy", mx[j, i], my[j, i], "dist to rounding boundary (1/32 px units)",
                  abs((fx % 1.0) - 0.5), abs((fy % 1.0) - 0.5), "device xy", x64[j, i], y64[j, i],
                  "lat row", my[j, i], "got", direct[k][j, i], "want", want[j, i])
