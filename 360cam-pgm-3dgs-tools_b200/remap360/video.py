"""Video input for the drop-in cutter: every view job of one video shares ONE decode.

The reference starts one ffmpeg process per view, each decoding the whole video again
(gs360_360PerspCut.py:746-749 output pattern, :569-590 execution).  Here the frames are decoded
once with OpenCV, pushed through the streaming remapper (pinned ring, H2D / kernel / D2H
overlapped) and every view of a frame is written as ``<stem>_%07d_<view>.<ext>`` numbered from 0.

The jobs' ``colorspace=iall=bt709:all=smpte170m[:trc=iec61966-2-1]`` step (gs360_360PerspCut.py:299-309)
runs on the device, in place on the uploaded frame, right before the remap -- the position it has in the
reference's filter chain (``r360_convert_color``; ``R360_VIDEO_COLOR=0`` leaves frames in the decoder's
BGR output)."""

from __future__ import annotations

import os
from typing import List, Optional, Sequence, Tuple


def _select_frames(n_in: int, in_fps: float, out_fps: float, start: Optional[float], end: Optional[float]):
    """ffmpeg ``fps=`` semantics (round=near): input frame i lands in output slot round(t_i * out_fps);
    the last frame landing in a slot wins, empty slots repeat the previous frame."""
    t0 = max(0.0, start or 0.0)
    slots = {}
    for i in range(n_in):
        t = i / in_fps
        if t < t0 or (end is not None and t > end):
            continue
        slots[int(round((t - t0) * out_fps))] = i
    if not slots:
        return []
    out, last = [], None
    for k in range(max(slots) + 1):
        last = slots.get(k, last)
        if last is not None:
            out.append(last)
    return out


def run_video_jobs(source, jobs: Sequence, stop_event=None) -> List[Tuple[int, str]]:
    import cv2
    import numpy as np
    import torch
    from . import api
    from .executor import _INTERP, _job_view, _write_image
    from .stream import StreamingRemapper

    results: List[Optional[Tuple[int, str]]] = [None] * len(jobs)
    usable = []
    for k, job in enumerate(jobs):
        if job.projection not in ("rectilinear", "fisheye"):
            results[k] = (1, "v360 output=%s is not available in the CUDA backend yet" % job.projection)
        elif job.interp not in _INTERP:
            results[k] = (1, "interp=%s is not available in the CUDA backend" % job.interp)
        else:
            usable.append(k)
    if usable:
        cap = cv2.VideoCapture(str(source))
        if not cap.isOpened():
            return [(1, "failed to open %s" % source)] * len(jobs)
        try:
            in_fps = cap.get(cv2.CAP_PROP_FPS) or 30.0
            n_in = int(cap.get(cv2.CAP_PROP_FRAME_COUNT))
            first = jobs[usable[0]]
            wanted = _select_frames(n_in, in_fps, float(first.fps), first.start, first.end)
            size = (first.width, first.height)
            views = [_job_view(jobs[k]) for k in usable]
            ok, frame = cap.read()
            if not ok:
                return [(1, "no frames in %s" % source)] * len(jobs)
            frame_filter = None
            if first.colorspace and os.environ.get("R360_VIDEO_COLOR", "1") != "0":
                from .color import convert_video_color

                def frame_filter(dev_frame, stream, _text=first.colorspace):
                    convert_video_color(dev_frame, filter_text=_text, channel_order="bgr", out=dev_frame, stream=stream)
            remapper = StreamingRemapper(views, size, frame.shape, torch.uint8, interp=_INTERP[first.interp],
                                         frame_filter=frame_filter)

            def frames():
                nonlocal frame
                pos, cur = 0, frame
                for want in wanted:
                    while pos < want:
                        okk, nxt = cap.read()
                        if not okk:
                            return
                        cur, pos = nxt, pos + 1
                    if stop_event is not None and stop_event.is_set():
                        return
                    yield torch.from_numpy(np.ascontiguousarray(cur))

            for n, out in enumerate(remapper.run(frames())):
                views_host = out.numpy()
                for col, k in enumerate(usable):
                    path = str(jobs[k].output) % n if "%" in str(jobs[k].output) else str(jobs[k].output)
                    _write_image(__import__("pathlib").Path(path), views_host[col], jobs[k].jpeg_quality, jobs[k].pix_fmt)
            cancelled = stop_event is not None and stop_event.is_set()
            for k in usable:
                results[k] = (130, "") if cancelled else (0, "")
        except Exception as exc:
            for k in usable:
                results[k] = (1, "%s: %s" % (type(exc).__name__, exc))
        finally:
            cap.release()
    return [r if r is not None else (1, "job skipped") for r in results]
