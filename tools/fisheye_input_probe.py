"""Throughput and fallback-tile count of the Video2Frames fisheye -> perspective path (V2F defaults: 190 degree lens,
8 mm, square output) on 3840^2 frames, both lens laws."""
import sys, json, pathlib, torch
ROOT = pathlib.Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT / "360cam-pgm-3dgs-tools_b200")); sys.path.insert(0, str(ROOT))
import remap360
from remap360 import api, fisheye_input as fi

frames = torch.randint(0, 256, (8, 3840, 3840, 3), dtype=torch.uint8, device="cuda")
for projection in ("equidistant", "equisolid"):
    for size in (1600, 3000):
        out = torch.empty((8, size, size, 3), dtype=torch.uint8, device="cuda")
        fn = lambda: fi.fisheye_to_perspective(frames, projection=projection, size_px=size, out=out)
        for _ in range(3): fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10): fn()
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 10
        plans = list(api._PLAN_CACHE.values())
        print(json.dumps({"projection": projection, "size": size, "ms": ms, "Gpix_per_s": 8 * size * size / ms / 1e6,
                          "fallback_tiles": plans[-1].n_fallback, "tiles": plans[-1].tiles_per_view}))
