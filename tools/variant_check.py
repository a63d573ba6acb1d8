"""Kernel experiments: run the bench workload (cfg 2) under the library named by R360_LIBRARY, print the
kernel-only throughput of both interpolations and a digest of the outputs so that variants can be
compared bit for bit with the shipped library.

    R360_LIBRARY=tools/variants/lib_x.so python tools/variant_check.py [frames]"""
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
    frames = int(sys.argv[1]) if len(sys.argv) > 1 else 8
    dev = torch.device("cuda")
    g = torch.Generator(device=dev)
    g.manual_seed(99)
    src = torch.randint(0, 256, (frames, 3840, 7680, 3), dtype=torch.uint8, device=dev, generator=g)
    views = [remap360.PerspectiveView(y, p, hf, vf) for _, y, p, hf, vf in preset_views("full360coverage", 1600)]
    views += [remap360.PerspectiveView(180.0, 0.0, views[0].hfov_deg, views[0].vfov_deg),
              remap360.PerspectiveView(0.0, 90.0, views[0].hfov_deg, views[0].vfov_deg)]
    out = torch.empty((frames, len(views), 1600, 1600, 3), dtype=torch.uint8, device=dev)
    row = {"library": os.environ.get("R360_LIBRARY", "shipped")}
    for interp in ("cubic", "linear"):
        def fn():
            remap360.remap_erp(src, views, (1600, 1600), interp=interp, out=out)
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(20):
            fn()
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 20
        row[interp] = {"ms": ms, "Gpix_per_s": out.numel() / 3 / ms / 1e6,
                       "sha": hashlib.sha256(out[:2].cpu().numpy().tobytes()).hexdigest()[:16]}
    print(json.dumps(row))


if __name__ == "__main__":
    main()
