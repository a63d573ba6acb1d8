"""Kernel experiments: launch shapes of the tiled kernel on the bench workload (cfg 2: 8K u8 RGB frames ->
full360coverage, 12 x 1600^2 views).  For the library named by R360_LIBRARY (default: the in-tree one) every
combination of frames per work item / blocks per SM / multi-frame budget is timed kernel-only (CUDA events) and a
digest of the first two frames' views is printed, so that shapes and variants can be compared bit for bit.

    R360_LIBRARY=tools/variants/lib_x.so python tools/shape_sweep.py [--frames 16] [--interp cubic linear]
        [--fr 1 2 4] [--ctas 0 1 2 3 4] [--pct 50 100]          (ctas 0 = the library's default)"""
import argparse
import hashlib
import json
import os
import pathlib
import sys

import torch

ROOT = pathlib.Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT / "360cam-pgm-3dgs-tools_b200"))
sys.path.insert(0, str(ROOT))
import remap360  # noqa: E402
from bench import preset_views  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--frames", type=int, default=16)
    ap.add_argument("--interp", nargs="+", default=["cubic", "linear"])
    ap.add_argument("--fr", nargs="+", type=int, default=[1, 2, 4], help="frames per item (0 = the library's choice)")
    ap.add_argument("--ctas", nargs="+", type=int, default=[0])
    ap.add_argument("--pct", nargs="+", type=int, default=[0], help="share of the ring a multi-frame item may take (0 = the library's choice)")
    ap.add_argument("--teams", nargs="+", type=int, default=[0], help="consumer teams for four-frame items (0 = the library's choice)")
    ap.add_argument("--iters", type=int, default=10)
    ap.add_argument("--dtype", default="u8")
    ns = ap.parse_args()
    dev = torch.device("cuda")
    g = torch.Generator(device=dev)
    g.manual_seed(99)
    if ns.dtype == "u16":
        src = torch.randint(0, 65536, (ns.frames, 3840, 7680, 3), dtype=torch.int32, device=dev, generator=g).to(torch.uint16)
    else:
        src = torch.randint(0, 256, (ns.frames, 3840, 7680, 3), dtype=torch.uint8, device=dev, generator=g)
    views = [remap360.PerspectiveView(y, p, hf, vf) for _, y, p, hf, vf in preset_views("full360coverage", 1600)]
    out = torch.empty((ns.frames, len(views), 1600, 1600, 3), dtype=src.dtype, device=dev)
    lib = os.environ.get("R360_LIBRARY", "shipped")
    for interp in ns.interp:
        for fr in ns.fr:
            for ctas in ns.ctas:
              for teams in (ns.teams if fr >= 2 else ns.teams[:1]):
                for pct in (ns.pct if fr > 1 else ns.pct[:1]):
                    if fr > 0:
                        os.environ["R360_FRAMES"] = str(fr)
                    else:
                        os.environ.pop("R360_FRAMES", None)
                    if teams > 0:
                        os.environ["R360_TEAMS"] = str(teams)
                    else:
                        os.environ.pop("R360_TEAMS", None)
                    if pct > 0:
                        os.environ["R360_MULTI_PCT"] = str(pct)
                    else:
                        os.environ.pop("R360_MULTI_PCT", None)
                    if ctas > 0:
                        os.environ["R360_TILED_CTAS_PER_SM"] = str(ctas)
                    else:
                        os.environ.pop("R360_TILED_CTAS_PER_SM", None)

                    from remap360 import api
                    api.clear_plan_cache()            # the blocks-per-SM variable also shapes the plan (ring, patch budget)

                    def fn():
                        remap360.remap_erp(src, views, (1600, 1600), interp=interp, out=out)
                    try:
                        for _ in range(3):
                            fn()
                        torch.cuda.synchronize()
                        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                        e0.record()
                        for _ in range(ns.iters):
                            fn()
                        e1.record()
                        torch.cuda.synchronize()
                    except Exception as exc:                      # a shape the library refuses
                        print(json.dumps({"library": lib, "interp": interp, "fr": fr, "ctas": ctas, "pct": pct, "error": str(exc)}), flush=True)
                        continue
                    ms = e0.elapsed_time(e1) / ns.iters
                    sha = hashlib.sha256(out[:2].view(torch.uint8).cpu().numpy().tobytes()).hexdigest()[:12]
                    print(json.dumps({"library": os.path.basename(lib), "interp": interp, "dtype": ns.dtype, "fr": fr, "ctas": ctas, "teams": teams,
                                      "pct": pct, "ms": round(ms, 4), "Gpix_per_s": round(out.numel() / 3 / ms / 1e6, 1),
                                      "sha": sha}), flush=True)


if __name__ == "__main__":
    main()
