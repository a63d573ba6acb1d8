"""File-to-file throughput of the PerspCut drop-in at the bench size: N synthetic 8K JPEG panoramas ->
full360coverage (12 x 1600^2 JPEG views each), GPU codec vs OpenCV codec.  Prints one JSON line per mode."""
import json
import os
import pathlib
import subprocess
import sys
import tempfile
import time

import cv2
import numpy as np

ROOT = pathlib.Path(__file__).resolve().parent.parent


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 8
    with tempfile.TemporaryDirectory() as tmp:
        tmp = pathlib.Path(tmp)
        (tmp / "in").mkdir()
        yy, xx = np.mgrid[0:3840, 0:7680].astype(np.float32)
        for k in range(n):
            img = np.stack([127 + 100 * np.sin(xx / 7680 * 6.2832 * (c + 1 + k)) * np.cos(yy / 3840 * 3.1416 * (c + 2))
                            for c in range(3)], axis=-1).astype(np.uint8)
            img += np.random.default_rng(k).integers(0, 12, img.shape, dtype=np.uint8)
            cv2.imwrite(str(tmp / "in" / ("pano%04d.jpg" % k)), img, [cv2.IMWRITE_JPEG_QUALITY, 92])
        sys.path.insert(0, str(ROOT / "360cam-pgm-3dgs-tools_b200"))
        from remap360 import executor, perspcut as pc
        files = sorted((tmp / "in").glob("*.jpg"))
        for mode in ("gpu", "cpu", "gpu"):
            os.environ.pop("R360_CPU_CODEC", None)
            if mode == "cpu":
                os.environ["R360_CPU_CODEC"] = "1"
            out = tmp / ("out_" + mode)
            args = pc.create_arg_parser().parse_args(["-i", str(tmp / "in"), "-o", str(out), "--preset", "full360coverage"])
            args.size_explicit = args.hfov_explicit = args.focal_mm_explicit = False
            args.input_is_video, args.video_bit_depth = False, 8
            warm = pc.build_view_jobs(args, files[:1], out)
            list(executor.run_jobs(warm.jobs, pc.stop_event, workers=1))           # plan, codec, CUDA context
            res = pc.build_view_jobs(args, files[1:], out)
            for workers in (1, 4):
                t0 = time.time()
                done = list(executor.run_jobs(res.jobs, pc.stop_event, workers=workers))
                dt = time.time() - t0
                ok = sum(1 for _j, (rc, _e) in done if rc == 0)
                print(json.dumps({"codec": mode, "workers": workers, "panoramas": len(files) - 1, "views_ok": ok,
                                  "seconds": round(dt, 3), "panoramas_per_s": round((len(files) - 1) / dt, 2),
                                  "views_per_s": round(ok / dt, 1)}))


if __name__ == "__main__":
    main()
