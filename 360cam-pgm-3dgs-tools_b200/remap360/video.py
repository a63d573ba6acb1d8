"""Video input for the drop-in cutter: every view job of one video shares ONE decode.

The reference starts one ffmpeg process per view, each decoding the whole video again
(gs360_360PerspCut.py:746-749 output pattern, :569-590 execution).  Here the frames are decoded
once, pushed through the streaming remapper (pinned ring, H2D / kernel / D2H overlapped, two frames
per launch) and every view of a frame is written as ``<stem>_%07d_<view>.<ext>`` numbered from 0 by a
pool of writer threads, so that encoding and disk I/O stay off the frame loop.

Decoders
  * Motion-JPEG streams (AVI / MOV ``MJPG``): the container is read in raw-packet mode and every packet is
    decoded on the GPU by nvJPEG (remap360/codec.py) on a small pool of host threads -- only compressed bytes
    cross the host link;
  * everything else: OpenCV's FFmpeg reader (multi-threaded decode, BGR frames on the host).
  ``R360_VIDEO_DECODER=opencv`` forces the second path.

Colour
  The jobs' ``colorspace=iall=bt709:all=smpte170m[:trc=iec61966-2-1]`` step (gs360_360PerspCut.py:299-309) runs
  on the device, in place on the uploaded frames, right before the remap -- the position it has in the
  reference's filter chain (``r360_convert_color``; ``R360_VIDEO_COLOR=0`` leaves frames in the decoder's BGR
  output).  Both decoders turn Y'CbCr into R'G'B' with the BT.601 matrix (swscale's default when the caller does not
  set colourspace details; JFIF for nvJPEG) while the filter declares the input to be BT.709 (``iall=bt709``), so the
  frames are first re-matrixed 601 -> 709 on the device (``color.decoder_matrix_correction``): what ffmpeg's
  own chain computes from the same Y'CbCr samples.  ``R360_VIDEO_DECODE_MATRIX=bt709`` says the decoder already
  used BT.709.

Frame selection follows ffmpeg: ``-ss S`` in front of ``-i`` restarts the timestamps at 0, so the output-side
``-to E`` keeps source times [S, S + E]; ``fps=`` puts input frame i into output slot round(t_i * fps), the last
frame landing in a slot wins, empty slots repeat the previous frame.

Multi-GPU: ``shard=(rank, world)`` restricts a call to a contiguous range of OUTPUT frame numbers
(remap360/sharding.py); numbering stays global, so the union over ranks is the single-GPU result."""

from __future__ import annotations

import os
import pathlib
import threading
from collections import OrderedDict, deque
from concurrent.futures import ThreadPoolExecutor
from typing import Iterator, List, Optional, Sequence, Tuple


def _select_frames(n_in: int, in_fps: float, out_fps: float, start: Optional[float], end: Optional[float]) -> List[int]:
    """Source frame index of every output frame (ffmpeg ``-ss S -i in -to E -vf fps=F``, round=near)."""
    t0 = max(0.0, start or 0.0)
    slots = {}
    for i in range(n_in):
        t = i / in_fps
        if t < t0 or (end is not None and (t - t0) > end):
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


def _count_frames(cap, cv2) -> int:
    """Frames in the stream: the container's count when it has one, else a grab() pass (decode without conversion)."""
    n = int(cap.get(cv2.CAP_PROP_FRAME_COUNT))
    if n > 0:
        return n
    n = 0
    while cap.grab():
        n += 1
    return n


def _bucket_key(job):
    return (job.width, job.height, job.interp, job.fps, job.start, job.end, job.colorspace)


class _OpenCvFrames:
    """BGR frames of the listed source indices (non-decreasing), decoded sequentially by OpenCV."""

    def __init__(self, source, wanted: Sequence[int]):
        import cv2
        self.cv2, self.source, self.wanted = cv2, str(source), list(wanted)

    def __iter__(self):
        import numpy as np
        import torch
        cv2 = self.cv2
        cap = cv2.VideoCapture(self.source)
        if not cap.isOpened():
            raise RuntimeError("failed to open %s" % self.source)
        try:
            pos, cur = -1, None                 # index of the frame held in `cur`
            if self.wanted and self.wanted[0] > 64:
                cap.set(cv2.CAP_PROP_POS_FRAMES, self.wanted[0])      # a shard in the middle of the file
                pos = self.wanted[0] - 1
            for want in self.wanted:
                while pos < want:
                    if pos + 1 < want:
                        if not cap.grab():                            # skipped frame: decode only
                            return
                    else:
                        ok, cur = cap.read()
                        if not ok:
                            return
                    pos += 1
                yield torch.from_numpy(np.ascontiguousarray(cur))
        finally:
            cap.release()


class _MjpegGpuFrames:
    """Frames of a Motion-JPEG stream decoded on the device: raw packets from the container, nvJPEG on `workers`
    host threads (one codec + one CUDA stream each), results handed over in order as CUDA tensors [H, W, 3] BGR."""

    def __init__(self, source, wanted: Sequence[int], device, workers: int = 6, ahead: int = 10):
        import cv2
        self.cv2, self.source, self.wanted, self.device = cv2, str(source), list(wanted), device
        self.workers, self.ahead = max(1, workers), max(2, ahead)

    @staticmethod
    def usable(source) -> bool:
        """True when the first packet of the stream is a baseline JPEG nvJPEG can decode on this device."""
        if os.environ.get("R360_VIDEO_DECODER", "").lower() == "opencv" or os.environ.get("R360_CPU_CODEC"):
            return False
        try:
            import cv2
            import torch
            from . import codec
            cap = cv2.VideoCapture(str(source))
            cap.set(cv2.CAP_PROP_FORMAT, -1)
            ok, pkt = cap.read()
            cap.release()
            if not ok or pkt is None or pkt.size < 4 or bytes(pkt.ravel()[:2]) != b"\xff\xd8":
                return False
            jc = codec.JpegCodec.for_thread()
            frame = jc.decode(bytes(pkt.ravel()))
            torch.cuda.current_stream().synchronize()
            return frame.dim() == 3 and frame.shape[2] == 3
        except Exception:
            return False

    def __iter__(self):
        import torch
        from . import codec
        cv2 = self.cv2
        cap = cv2.VideoCapture(self.source)
        cap.set(cv2.CAP_PROP_FORMAT, -1)                       # raw packets: one JPEG per frame
        if not cap.isOpened():
            raise RuntimeError("failed to open %s" % self.source)
        tls = threading.local()
        dev = self.device

        def decode(data: bytes):
            if getattr(tls, "jc", None) is None:
                with torch.cuda.device(dev):
                    tls.jc, tls.stream = codec.JpegCodec(dev), torch.cuda.Stream(dev)
            with torch.cuda.device(dev), torch.cuda.stream(tls.stream):
                frame = tls.jc.decode(data, stream=tls.stream)
            tls.stream.synchronize()
            return frame

        pool = ThreadPoolExecutor(max_workers=self.workers)
        try:
            pending: deque = deque()
            pos, data = -1, None
            it = iter(self.wanted)
            done_reading = False
            last_want, last_future = None, None
            while True:
                while not done_reading and len(pending) < self.ahead:
                    try:
                        want = next(it)
                    except StopIteration:
                        done_reading = True
                        break
                    if want == last_want:                          # a repeated output frame: decode once
                        pending.append(last_future)
                        continue
                    while pos < want:
                        ok, pkt = cap.read()
                        if not ok:
                            done_reading = True
                            break
                        data, pos = bytes(pkt.ravel()), pos + 1
                    if done_reading and pos < want:
                        break
                    last_want, last_future = want, pool.submit(decode, data)
                    pending.append(last_future)
                if not pending:
                    break
                yield pending.popleft().result()
        finally:
            pool.shutdown(wait=True)
            cap.release()


def run_video_jobs(source, jobs: Sequence, stop_event=None, shard: Tuple[int, int] = (0, 1), device=None,
                   writers: Optional[int] = None) -> List[Tuple[int, str]]:
    """Run the view jobs of one video source; returns (rc, err) per job.  Jobs that differ in size, interpolation,
    frame rate, time window or colour filter are rendered in separate passes (each pass decodes the frames it
    needs once)."""
    import cv2
    import torch
    from .executor import _INTERP

    device = torch.device(device) if device is not None else torch.device("cuda", torch.cuda.current_device())
    results: List[Optional[Tuple[int, str]]] = [None] * len(jobs)
    buckets: "OrderedDict[tuple, List[int]]" = OrderedDict()
    for k, job in enumerate(jobs):
        if job.projection not in ("rectilinear", "fisheye"):
            results[k] = (1, "v360 output=%s is not available in the CUDA backend yet" % job.projection)
        elif job.interp not in _INTERP:
            results[k] = (1, "interp=%s is not available in the CUDA backend" % job.interp)
        elif not job.fps or job.fps <= 0:
            results[k] = (1, "video job without a frame rate")
        else:
            buckets.setdefault(_bucket_key(job), []).append(k)
    if not buckets:
        return [r if r is not None else (1, "job skipped") for r in results]
    cap = cv2.VideoCapture(str(source))
    if not cap.isOpened():
        return [(1, "failed to open %s" % source)] * len(jobs)
    try:
        in_fps = cap.get(cv2.CAP_PROP_FPS) or 30.0
        n_in = _count_frames(cap, cv2)
    finally:
        cap.release()
    if n_in <= 0:
        return [(1, "no frames in %s" % source) if r is None else r for r in results]
    mjpeg_gpu = _MjpegGpuFrames.usable(source)
    pool = ThreadPoolExecutor(max_workers=writers or max(2, min(16, (os.cpu_count() or 4) - 2)))
    try:
        for idxs in buckets.values():
            try:
                rc = _run_bucket(source, [jobs[k] for k in idxs], n_in, in_fps, stop_event, shard, device, pool, mjpeg_gpu)
                for k in idxs:
                    results[k] = rc
            except Exception as exc:
                for k in idxs:
                    results[k] = (1, "%s: %s" % (type(exc).__name__, exc))
    finally:
        pool.shutdown(wait=True)
    return [r if r is not None else (1, "job skipped") for r in results]


def _run_bucket(source, jobs, n_in, in_fps, stop_event, shard, device, pool, mjpeg_gpu) -> Tuple[int, str]:
    import cv2
    import torch
    from .executor import _INTERP, _job_view, _write_image
    from .sharding import shard_range
    from .stream import StreamingRemapper

    first = jobs[0]
    wanted_all = _select_frames(n_in, in_fps, float(first.fps), first.start, first.end)
    if not wanted_all:
        return 1, "no frame of %s falls into the requested time window" % source
    lo, hi = shard_range(len(wanted_all), shard[1], shard[0])
    wanted = wanted_all[lo:hi]
    if not wanted:
        return 0, ""                                       # this rank owns no frame of the source
    views = [_job_view(j) for j in jobs]
    cap = cv2.VideoCapture(str(source))
    w_in, h_in = int(cap.get(cv2.CAP_PROP_FRAME_WIDTH)), int(cap.get(cv2.CAP_PROP_FRAME_HEIGHT))
    cap.release()
    frame_filter = None
    if first.colorspace and os.environ.get("R360_VIDEO_COLOR", "1") != "0":
        from .color import convert_video_color, correct_decoder_matrix
        fix_matrix = os.environ.get("R360_VIDEO_DECODE_MATRIX", "bt601").lower() != "bt709"

        def frame_filter(dev_frames, stream, _text=first.colorspace):
            if fix_matrix:
                correct_decoder_matrix(dev_frames, channel_order="bgr", out=dev_frames, stream=stream)
            convert_video_color(dev_frames, filter_text=_text, channel_order="bgr", out=dev_frames, stream=stream)
    with torch.cuda.device(device):
        from .executor import job_convention
        remapper = StreamingRemapper(views, (first.width, first.height), (h_in, w_in, 3), torch.uint8,
                                     interp=_INTERP[first.interp], convention=job_convention(), frame_filter=frame_filter,
                                     device=device, depth=6, batch=2, hold=2)
        source_frames = (_MjpegGpuFrames(source, wanted, device) if mjpeg_gpu else _OpenCvFrames(source, wanted))

        from .executor import _stage

        def frames() -> Iterator:
            it = iter(source_frames)
            while True:
                with _stage("video_decode_wait"):
                    fr = next(it, None)
                if fr is None or (stop_event is not None and stop_event.is_set()):
                    return
                yield fr

        # A result is a view of a pinned ring slot.  The remapper keeps a handed-out batch untouched until two more
        # batches have been handed out (hold=2), so the loop does not copy anything itself: a few dedicated threads
        # copy the views of a frame out of the ring and pass each copy on to the writer pool (encoding, disk I/O);
        # the loop only makes sure, before it advances, that the copies of the frame before last are done, and
        # bounds the number of views waiting for a writer.
        import numpy as np
        futures, copying, produced = deque(), deque(), 0
        copiers = ThreadPoolExecutor(max_workers=4)

        # JPEG views are encoded by the writer threads' cv2.imwrite -- and, while one of a few permits is free, by
        # nvJPEG on the device instead (the view goes back up, 7.7 MB; the encoder is ~5 x a host thread).  The two
        # encoders write the same format (4:4:4, the job's quality); R360_VIDEO_GPU_ENCODERS=0 leaves all to the host.
        gpu_permits = threading.BoundedSemaphore(max(1, int(os.environ.get("R360_VIDEO_GPU_ENCODERS", "3"))))
        gpu_state = {"on": int(os.environ.get("R360_VIDEO_GPU_ENCODERS", "3")) > 0 and not os.environ.get("R360_CPU_CODEC")}
        tls = threading.local()

        def encode_on_device(path, local, quality) -> bool:
            try:
                from . import codec
                if getattr(tls, "jc", None) is None:
                    with torch.cuda.device(device):
                        tls.jc, tls.stream = codec.JpegCodec(device), torch.cuda.Stream(device)
                with torch.cuda.device(device), torch.cuda.stream(tls.stream):
                    dev_view = torch.from_numpy(local).to(device, non_blocking=False)
                    data = tls.jc.encode(dev_view, quality, stream=tls.stream)
                path.parent.mkdir(parents=True, exist_ok=True)
                path.write_bytes(data)
                return True
            except Exception:
                gpu_state["on"] = False                    # no codec library / not this kind of image: host encoder from now on
                return False

        def write_view(path, local, quality, pix_fmt):
            with _stage("video_write"):
                if (gpu_state["on"] and path.suffix.lower() in (".jpg", ".jpeg") and local.dtype == np.uint8
                        and local.ndim == 3 and local.shape[2] == 3 and gpu_permits.acquire(blocking=False)):
                    try:
                        if encode_on_device(path, local, quality):
                            return
                    finally:
                        gpu_permits.release()
                _write_image(path, local, quality, pix_fmt)

        def copy_view(path, pinned_view, quality, pix_fmt):
            return pool.submit(write_view, path, np.copy(pinned_view), quality, pix_fmt)

        try:
            for n, out in enumerate(remapper.run(frames())):
                views_host = out.numpy()
                batch = []
                for col, job in enumerate(jobs):
                    path = str(job.output) % (lo + n) if "%" in str(job.output) else str(job.output)
                    batch.append(copiers.submit(copy_view, pathlib.Path(path), views_host[col], job.jpeg_quality, job.pix_fmt))
                copying.append(batch)
                produced += 1
                with _stage("video_copy_wait"):
                    while len(copying) > 2:                 # frames n - 2 and older: out of the ring before it turns
                        futures.extend(c.result() for c in copying.popleft())
                with _stage("video_writer_backpressure"):
                    while len(futures) > 96:                # bound the views waiting for a writer
                        futures.popleft().result()
            for batch in copying:
                futures.extend(c.result() for c in batch)
            for f in futures:
                f.result()
        finally:
            copiers.shutdown(wait=True)
    if stop_event is not None and stop_event.is_set():
        return 130, ""
    if produced == 0:
        return 1, "no frame of %s could be decoded" % source
    return 0, ""
