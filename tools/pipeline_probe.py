"""File-to-file throughput of the PerspCut drop-in at the bench size: N synthetic 8K JPEG panoramas ->
full360coverage (12 x 1600^2 JPEG views each), GPU codec vs OpenCV codec, 1 .. 16 host threads, with the seconds
each stage of the runner took (summed over threads); then an 8K Motion-JPEG clip through the video branch with the
nvJPEG and the OpenCV decoder.  Prints one JSON line per run.

    python tools/pipeline_probe.py [panoramas] [video frames]"""
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
    n_video = int(sys.argv[2]) if len(sys.argv) > 2 else 24
    os.environ["R360_TIMING"] = "1"
    # files in memory-backed storage when the box has it: the probe times the pipeline, not the container's overlay disk
    # (R360_PROBE_DIR overrides)
    scratch = os.environ.get("R360_PROBE_DIR") or ("/dev/shm" if os.path.isdir("/dev/shm") and os.access("/dev/shm", os.W_OK) else None)
    with tempfile.TemporaryDirectory(dir=scratch) as tmp:
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
        gpu_workers = tuple(int(t) for t in os.environ.get("R360_PROBE_WORKERS", "1,4,8,12,16").split(","))
        for mode in (("gpu",) if os.environ.get("R360_PROBE_SKIP_CPU") else ("gpu", "cpu")):
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
            for workers in (gpu_workers if mode == "gpu" else (4, 16)):
                list(executor.run_jobs(res.jobs, pc.stop_event, workers=workers))   # new threads build their codecs here
                executor.STAGE_SECONDS.clear()
                t0 = time.time()
                done = list(executor.run_jobs(res.jobs, pc.stop_event, workers=workers))
                dt = time.time() - t0
                ok = sum(1 for _j, (rc, _e) in done if rc == 0)
                print(json.dumps({"codec": mode, "scratch": scratch or "tmp", "jpeg_backend": os.environ.get("R360_JPEG_BACKEND", "gpu_hybrid"),
                                  "huffman": os.environ.get("R360_JPEG_HUFFMAN", "optimised"), "workers": workers,
                                  "panoramas": len(files) - 1, "views_ok": ok,
                                  "seconds": round(dt, 3), "panoramas_per_s": round((len(files) - 1) / dt, 2),
                                  "views_per_s": round(ok / dt, 1),
                                  "stage_seconds": {k: round(v, 3) for k, v in executor.STAGE_SECONDS.items()}}), flush=True)
        os.environ.pop("R360_CPU_CODEC", None)
        # ---- video branch: an 8K Motion-JPEG clip, 2 views per frame as PNG-free JPEG views
        clip = tmp / "clip.avi"
        wr = cv2.VideoWriter(str(clip), cv2.VideoWriter_fourcc(*"MJPG"), 30.0, (7680, 3840))
        if wr.isOpened() and n_video > 0:
            base = cv2.imread(str(files[0]))
            for k in range(n_video):
                wr.write(np.roll(base, 64 * k, axis=1))
            wr.release()
            for decoder in ("auto", "opencv"):
                os.environ["R360_VIDEO_DECODER"] = decoder
                out = tmp / ("vout_" + decoder)
                args = pc.create_arg_parser().parse_args(["-i", str(clip), "-o", str(out), "--preset", "full360coverage", "-f", "30"])
                args.size_explicit = args.hfov_explicit = args.focal_mm_explicit = False
                args.input_is_video, args.video_bit_depth = True, 8
                res = pc.build_view_jobs(args, [clip], out)
                for attempt in range(2):                       # the first pass pays for plans and codecs
                    executor.STAGE_SECONDS.clear()
                    t0 = time.time()
                    done = list(executor.run_jobs(res.jobs, pc.stop_event, workers=1))
                    dt = time.time() - t0
                ok = sum(1 for _j, (rc, _e) in done if rc == 0)
                n_files = len(list(out.glob("*.jpg")))
                print(json.dumps({"video": "8K MJPG clip, %d frames -> 12 views" % n_video, "decoder": decoder, "jobs_ok": ok,
                                  "files": n_files, "seconds": round(dt, 3), "frames_per_s": round(n_video / dt, 2),
                                  "views_per_s": round(n_files / dt, 1),
                                  "stage_seconds": {k: round(v, 3) for k, v in executor.STAGE_SECONDS.items()}, "errors": sorted({e for _j, (rc, e) in done if rc})[:2]}), flush=True)


if __name__ == "__main__":
    main()
