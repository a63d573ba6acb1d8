"""Kernel-only throughput of the other BASELINE.json configurations (the bench line covers configs[1]).
Writes one JSON object per line: config 3 (fisheyelike), config 4 (16-bit bicubic, u16 and fp16 out, seam and
pole views), config 5 (dual fisheye -> 10 x 1750^2), plus nearest / direct-path reference points."""
import json
import pathlib
import sys

import torch

ROOT = pathlib.Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT / "360cam-pgm-3dgs-tools_b200"))
sys.path.insert(0, str(ROOT))
import remap360  # noqa: E402
from bench import preset_views  # noqa: E402


def timeit(fn, steps=10, warmup=3):
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / steps


def main():
    dev = torch.device("cuda")
    W, H = 7680, 3840
    rows = []

    def erp_case(name, preset, dtype, interp, frames, out_dtype=None, extra_views=(), path="auto"):
        views = [remap360.PerspectiveView(y, p, hf, vf) for _, y, p, hf, vf in preset_views(preset, 1600)]
        views += [remap360.PerspectiveView(y, p, views[0].hfov_deg, views[0].vfov_deg) for y, p in extra_views]
        if dtype == torch.uint16:
            src = torch.randint(0, 32767, (frames, H, W, 3), dtype=torch.int16, device=dev).view(torch.uint16)
        else:
            src = torch.randint(0, 256, (frames, H, W, 3), dtype=torch.uint8, device=dev)
        out = torch.empty((frames, len(views), 1600, 1600, 3), dtype=out_dtype or dtype, device=dev)
        ms = timeit(lambda: remap360.remap_erp(src, views, (1600, 1600), interp=interp, out=out,
                                               out_dtype=out_dtype, path=path))
        pix = frames * len(views) * 1600 * 1600
        rows.append({"case": name, "ms": ms, "Mpix_per_s": pix / ms / 1e3, "views": len(views), "frames": frames,
                     "dtype": str(dtype), "out": str(out_dtype or dtype), "interp": interp, "path": path})
        del src, out
        torch.cuda.empty_cache()

    erp_case("cfg3 fisheyelike u8 cubic", "fisheyelike", torch.uint8, "cubic", 16)
    erp_case("cfg3 fisheyelike u8 linear", "fisheyelike", torch.uint8, "linear", 16)
    hard = [(180, 0), (179.9, 0), (-179.9, 0), (0, 90), (0, -90), (0, 60), (0, -60)]
    erp_case("cfg4 u16 cubic -> u16 (12 preset + seam/pole views)", "full360coverage", torch.uint16, "cubic", 4, None, hard)
    erp_case("cfg4 u16 cubic -> f16 (12 preset + seam/pole views)", "full360coverage", torch.uint16, "cubic", 4, torch.float16, hard)
    erp_case("cfg2 u8 cubic + seam/pole views", "full360coverage", torch.uint8, "cubic", 8, None, hard)
    erp_case("cfg2 u8 nearest", "full360coverage", torch.uint8, "nearest", 8)
    erp_case("cfg2 u8 cubic, direct path", "full360coverage", torch.uint8, "cubic", 2, None, (), "direct")
    erp_case("cfg2 u8 linear, direct path", "full360coverage", torch.uint8, "linear", 2, None, (), "direct")

    # config 5: dual fisheye, template calibration values (tests/golden/dualfisheye.json)
    meta = json.loads((ROOT / "tests" / "golden" / "dualfisheye.json").read_text())
    cal = meta["sensors"]["0"]
    calib = remap360.FisheyeCalibration(**{k: cal[k] for k in ("width", "height", "f", "cx", "cy", "k1", "k2", "k3",
                                                               "k4", "p1", "p2", "b1", "b2")}, lens_fov_deg=190.0)
    lens_of = {vid: (0 if i["lens_key"] == "X" else 1) for vid, i in meta["maps_1750"]["views"].items()}
    views = []
    for s in meta["sfm10_default"]:
        slot = lens_of[s["view_id"]]
        yaw_rel = ((s["yaw_deg"] - (0.0, 180.0)[slot] + 180.0) % 360.0) - 180.0
        views.append(remap360.PerspectiveView(yaw_rel, s["pitch_deg"], s["hfov_deg"], s["vfov_deg"], src_slot=slot))
    pairs = torch.randint(0, 256, (8, 2, 3840, 3840, 3), dtype=torch.uint8, device=dev)
    out = remap360.alloc_views(8, 10, 1750, 1750, 3, torch.uint8, dev)       # rows padded to 16 bytes
    for interp in ("cubic", "linear"):
        ms = timeit(lambda: remap360.remap_fisheye(pairs, [calib, calib], views, (1750, 1750), interp=interp, out=out))
        from remap360 import api
        plan = list(api._PLAN_CACHE.values())[-1]
        rows.append({"case": "cfg5 dual fisheye u8 %s, 8 pairs x 10 views 1750^2" % interp, "ms": ms,
                     "fallback_tiles": plan.n_fallback, "tiles": plan.tiles_per_view * plan.n_views,
                     "Mpix_per_s": 8 * 10 * 1750 * 1750 / ms / 1e3, "pairs_per_s": 8 / ms * 1e3, "interp": interp})
    # a12: both lenses 3840^2 -> undistorted 3840^2 (same template calibration)
    items = [remap360.UndistortItem(1.0, 0), remap360.UndistortItem(1.0, 1)]
    out_u = torch.empty((8, 2, 3840, 3840, 3), dtype=torch.uint8, device=dev)
    for interp in ("cubic", "linear"):
        ms = timeit(lambda: remap360.undistort_fisheye(pairs, [calib, calib], items, interp=interp, out=out_u))
        plan = list(api._PLAN_CACHE.values())[-1]
        rows.append({"case": "a12 undistort u8 %s, 8 pairs x 2 lenses 3840^2" % interp, "ms": ms,
                     "fallback_tiles": plan.n_fallback, "tiles": plan.tiles_per_view * plan.n_views,
                     "Mpix_per_s": 8 * 2 * 3840 * 3840 / ms / 1e3, "pairs_per_s": 8 / ms * 1e3, "interp": interp})
    del pairs, out, out_u
    torch.cuda.empty_cache()
    erp_case("cfg2 u8 lanczos4 (packed sampler, table in shared memory)", "full360coverage", torch.uint8, "lanczos4", 2)
    # preset fisheyeXY: 2 equidistant fisheye views 3600^2, d_fov 180
    hf, vf = remap360.api.fisheye_fov_from_dfov(180.0, 3600, 3600)
    fviews = [remap360.PerspectiveView(y, 0.0, hf, vf, projection="fisheye") for y in (0.0, 180.0)]
    src = torch.randint(0, 256, (8, H, W, 3), dtype=torch.uint8, device=dev)
    out_f = torch.empty((8, 2, 3600, 3600, 3), dtype=torch.uint8, device=dev)
    ms = timeit(lambda: remap360.remap_erp(src, fviews, (3600, 3600), interp="cubic", out=out_f))
    plan = list(api._PLAN_CACHE.values())[-1]
    rows.append({"case": "fisheyeXY u8 cubic, 8 frames x 2 views 3600^2", "ms": ms, "fallback_tiles": plan.n_fallback,
                 "tiles": plan.tiles_per_view * plan.n_views, "Mpix_per_s": 8 * 2 * 3600 * 3600 / ms / 1e3})
    for r in rows:
        print(json.dumps(r))


if __name__ == "__main__":
    main()
