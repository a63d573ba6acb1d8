"""Runs planner jobs on the GPU.

A job is the ffmpeg-style argv that ``perspcut.build_view_jobs`` emits (and that the GUI may
have edited, gs360_GUI.py:19092-19147).  The reference hands each argv to an ffmpeg process
(gs360_360PerspCut.py:569-590); here the argv is parsed back into (source, view, output) and
executed with the CUDA remap.  Jobs that share a source are grouped so the source is decoded and
uploaded once."""

from __future__ import annotations

import os
import pathlib
import threading
import traceback
from collections import OrderedDict
from concurrent.futures import ThreadPoolExecutor
from dataclasses import dataclass
from typing import Dict, Iterator, List, Optional, Sequence, Tuple


@dataclass
class ParsedJob:
    source: pathlib.Path
    output: pathlib.Path
    projection: str            # "rectilinear" | "fisheye"
    width: int
    height: int
    yaw: float
    pitch: float
    roll: float
    hfov: float
    vfov: float
    interp: str
    video: bool
    fps: Optional[float]
    start: Optional[float]
    end: Optional[float]
    jpeg_quality: int
    pix_fmt: Optional[str]
    colorspace: Optional[str] = None     # the job's `colorspace=...` filter (video sources, PC:299-309)


class JobError(RuntimeError):
    pass


def parse_job_argv(argv: Sequence[str]) -> ParsedJob:
    """Recover the job from the argv of gs360_360PerspCut.py:286-414."""
    argv = list(argv)

    def opt(flag: str) -> Optional[str]:
        return argv[argv.index(flag) + 1] if flag in argv and argv.index(flag) + 1 < len(argv) else None

    source, vf = opt("-i"), opt("-vf")
    if source is None or vf is None or len(argv) < 2:
        raise JobError("not a view job: missing -i / -vf")
    v360 = next((f for f in vf.split(",") if f.startswith("v360=")), None)
    if v360 is None:
        raise JobError("not a view job: no v360 filter in -vf")
    kv = dict(item.split("=", 1) for item in v360[len("v360="):].split(":") if "=" in item)
    if kv.get("input") != "equirect":
        raise JobError("unsupported v360 input: %s" % kv.get("input"))
    fps = next((float(f.split("=", 1)[1]) for f in vf.split(",") if f.startswith("fps=")), None)
    colorspace = next((f for f in vf.split(",") if f.startswith("colorspace=")), None)
    q = opt("-q:v")
    proj = kv.get("output", "rectilinear")
    if proj == "fisheye":
        hfov = vfov = float(kv["d_fov"])
    else:
        hfov, vfov = float(kv["h_fov"]), float(kv["v_fov"])
    return ParsedJob(source=pathlib.Path(source), output=pathlib.Path(argv[-1]), projection=proj,
                     width=int(kv["w"]), height=int(kv["h"]), yaw=float(kv.get("yaw", 0.0)),
                     pitch=float(kv.get("pitch", 0.0)), roll=float(kv.get("roll", 0.0)), hfov=hfov, vfov=vfov,
                     interp=kv.get("interp", "cubic"), video=fps is not None, fps=fps,
                     start=float(opt("-ss")) if opt("-ss") else None, end=float(opt("-to")) if opt("-to") else None,
                     jpeg_quality=95 if q == "2" else 100, pix_fmt=opt("-pix_fmt"), colorspace=colorspace)


_INTERP = {"cubic": "cubic", "linear": "linear", "nearest": "nearest", "near": "nearest",
           "lanczos": "lanczos4"}          # v360 `lanczos` -> the cv2-compatible 8x8 Lanczos kernel
_gpu_slots = threading.BoundedSemaphore(max(4, min(32, os.cpu_count() or 4)))   # source groups resident on the device at once (an 8K group is ~0.2 GB of 180)
_pool_lock = threading.Lock()
_idle_workers: dict = {}          # device index -> [(stream, codec or None)], reused across run_jobs calls


# Wall-clock seconds per stage of the still-image runner, summed over host threads (R360_TIMING=1; read by
# tools/pipeline_probe.py to see what a file-to-file run waits for).
STAGE_SECONDS: Dict[str, float] = {}
_stage_lock = threading.Lock()


class _stage:
    def __init__(self, name: str, stream=None):
        self.name, self.stream, self.on = name, stream, bool(os.environ.get("R360_TIMING"))

    def __enter__(self):
        if self.on:
            import time
            self.t0 = time.perf_counter()
        return self

    def __exit__(self, *exc):
        if self.on:
            import time
            if self.stream is not None:
                self.stream.synchronize()
            with _stage_lock:
                STAGE_SECONDS[self.name] = STAGE_SECONDS.get(self.name, 0.0) + time.perf_counter() - self.t0
        return False


class _DeviceWorker:
    """A CUDA stream plus an nvJPEG codec, borrowed by one host thread at a time: decode, remap and encode
    of different sources overlap, and the (expensive) codec objects outlive the thread pools."""

    def __enter__(self):
        import torch
        self._dev = torch.cuda.current_device()
        with _pool_lock:
            idle = _idle_workers.setdefault(self._dev, [])
            item = idle.pop() if idle else None
        if item is None:
            item = (torch.cuda.Stream(), _gpu_codec(), {})
        self._item = item
        self.stream = item[0]
        self.codec = None if os.environ.get("R360_CPU_CODEC") else item[1]
        self._pinned = item[2]                 # (slot, shape) -> pinned staging buffer for views encoded on the host
        return self

    def pinned_view(self, slot: int, shape):
        """A pinned uint8 buffer of `shape`, kept with the worker (views that leave for the host encoders)."""
        import torch
        key = (slot, tuple(shape))
        buf = self._pinned.get(key)
        if buf is None:
            if len(self._pinned) >= 16:
                self._pinned.clear()
            buf = self._pinned[key] = torch.empty(tuple(shape), dtype=torch.uint8).pin_memory()
        return buf

    def __exit__(self, *exc):
        with _pool_lock:
            _idle_workers.setdefault(self._dev, []).append(self._item)
        return False


def job_convention() -> str:
    """Pixel convention of the ERP jobs.  The default, "halfpixel", is the geometry the reference states in its own
    code (pixel centres, gs360_GUI.py:419-424) and what every parity test pins; R360_JOB_CONVENTION=v360 switches
    the job runners to FFmpeg's `(W - 1)` scaling (up to half a pixel of difference across the
    panorama) for outputs that must line up with views an ffmpeg run produced earlier.  Note also that v360's
    `lanczos` is a 4 x 4 kernel where the CUDA backend (like cv2) uses 8 x 8 taps."""
    conv = os.environ.get("R360_JOB_CONVENTION", "halfpixel").lower()
    return conv if conv in ("halfpixel", "v360") else "halfpixel"


def _job_view(job: "ParsedJob"):
    """The device view of a parsed job.  ``output=fisheye`` jobs carry ``d_fov`` (PC:375-379); v360
    turns it into per-axis FOVs before building its map."""
    from . import api
    if job.projection == "fisheye":
        hfov, vfov = api.fisheye_fov_from_dfov(job.hfov, job.width, job.height)
        return api.PerspectiveView(job.yaw, job.pitch, hfov, vfov, roll_deg=job.roll, projection="fisheye")
    return api.PerspectiveView(job.yaw, job.pitch, job.hfov, job.vfov, roll_deg=job.roll)


def widen_for_pix_fmt(image, pix_fmt: Optional[str]):
    """``-pix_fmt rgb48le`` (PC:346: PNG / TIFF views of a video whose source has more than 8 bits): the reference's
    chain computes in 8-bit yuv444p (PC:306-309) and lets swscale widen the result, which replicates the byte
    (v -> v * 257).  8-bit images for any other pixel format are written as they are."""
    if pix_fmt == "rgb48le" and image.dtype.itemsize == 1:
        import numpy as np
        return image.astype(np.uint16) * np.uint16(257)
    return image


def narrow_to_8bit(image):
    """16-bit samples -> 8 bits for a JPEG view, full range to full range with rounding: round(v * 255 / 65535)
    = (v * 255 + 32767) // 65535 -- what swscale's rgb48 -> 8-bit conversion amounts to (the reference's
    `-pix_fmt yuvj444p` after a 16-bit still, PC:317-339).  cv2.imwrite would SATURATE instead (every sample
    >= 255 becomes 255).  Accepts a NumPy array or a CUDA / CPU tensor; 8-bit input is returned as it is."""
    import numpy as np
    try:
        import torch
        if isinstance(image, torch.Tensor):
            if image.dtype != torch.uint16:
                return image
            return ((image.to(torch.int32) * 255 + 32767) // 65535).to(torch.uint8)
    except ImportError:
        pass
    if image.dtype != np.uint16:
        return image
    return ((image.astype(np.uint32) * 255 + 32767) // 65535).astype(np.uint8)


def _write_image(path: pathlib.Path, image, jpeg_quality: int, pix_fmt: Optional[str] = None) -> None:
    import cv2
    path.parent.mkdir(parents=True, exist_ok=True)
    if path.suffix.lower() in (".png", ".tif", ".tiff"):
        image = widen_for_pix_fmt(image, pix_fmt)
    if path.suffix.lower() in (".jpg", ".jpeg"):
        image = narrow_to_8bit(image)
    params: List[int] = []
    if path.suffix.lower() in (".jpg", ".jpeg"):
        params = [int(cv2.IMWRITE_JPEG_QUALITY), int(jpeg_quality)]
        if hasattr(cv2, "IMWRITE_JPEG_SAMPLING_FACTOR"):      # yuvj444p, as the reference asks of ffmpeg
            params += [int(cv2.IMWRITE_JPEG_SAMPLING_FACTOR), int(cv2.IMWRITE_JPEG_SAMPLING_FACTOR_444)]
    if not cv2.imwrite(str(path), image, params):
        raise JobError("failed to write %s" % path)


def _gpu_codec():
    """The calling thread's nvJPEG codec, or None when R360_CPU_CODEC is set or the library is unusable
    (OpenCV's codec writes the same file format)."""
    if os.environ.get("R360_CPU_CODEC"):
        return None
    try:
        from . import codec
        return codec.JpegCodec()
    except Exception:
        return None


_host_writer_pool = None
_host_writer_lock = threading.Lock()


def _host_writers() -> ThreadPoolExecutor:
    """Host threads that encode and write the views the device encoder does not take."""
    global _host_writer_pool
    with _host_writer_lock:
        if _host_writer_pool is None:
            _host_writer_pool = ThreadPoolExecutor(max_workers=max(2, (os.cpu_count() or 4) * 3 // 4))
        return _host_writer_pool


def _host_encode_share() -> int:
    """Per cent of a panorama's JPEG views encoded on the host instead of by nvJPEG (R360_STILL_HOST_ENCODE_PCT)."""
    try:
        return min(100, max(0, int(os.environ.get("R360_STILL_HOST_ENCODE_PCT", "40"))))
    except ValueError:
        return 40


def _is_jpeg(path: pathlib.Path) -> bool:
    return path.suffix.lower() in (".jpg", ".jpeg")


def _run_still_group(source: pathlib.Path, jobs: List[ParsedJob], stop_event) -> List[Tuple[int, str]]:
    """All views of one still image: one decode, one upload, one batched launch per output size.
    JPEG sources are decoded and JPEG views encoded on the GPU (remap360/codec.py) when possible."""
    import cv2
    import numpy as np
    import torch
    from . import api

    with _DeviceWorker() as worker:
        return _run_still_group_on(worker, source, jobs, stop_event)


def _run_still_group_on(worker, source: pathlib.Path, jobs: List[ParsedJob], stop_event) -> List[Tuple[int, str]]:
    import cv2
    import numpy as np
    import torch
    from . import api

    jc, stream = worker.codec, worker.stream
    image = dev_image = None
    if jc is not None and _is_jpeg(source):
        try:
            with _stage("read"):
                data = source.read_bytes()
            with _stage("decode", stream), torch.cuda.stream(stream):
                dev_image = jc.decode(data, stream=stream)   # [H, W, C] uint8, BGR like cv2
            image = np.empty(tuple(dev_image.shape), dtype=np.uint8)   # shape / dtype carrier only
        except Exception:
            dev_image = None                                         # e.g. progressive JPEG: OpenCV reads it
    if dev_image is None:
        with _stage("decode"):
            image = cv2.imread(str(source), cv2.IMREAD_UNCHANGED)
    if image is None:
        return [(1, "failed to read %s" % source)] * len(jobs)
    if image.ndim == 2:
        image = image[..., None]
    if image.dtype not in (np.uint8, np.uint16):
        return [(1, "unsupported sample type %s in %s" % (image.dtype, source))] * len(jobs)
    results: List[Optional[Tuple[int, str]]] = [None] * len(jobs)
    buckets: Dict[Tuple[int, int, str], List[int]] = OrderedDict()
    for k, job in enumerate(jobs):
        if job.projection not in ("rectilinear", "fisheye"):
            results[k] = (1, "v360 output=%s is not available in the CUDA backend yet" % job.projection)
            continue
        if job.interp not in _INTERP:
            results[k] = (1, "interp=%s is not available in the CUDA backend" % job.interp)
            continue
        buckets.setdefault((job.width, job.height, _INTERP[job.interp]), []).append(k)
    host = np.ascontiguousarray(image)
    for (w, h, interp), idxs in buckets.items():
        if stop_event is not None and stop_event.is_set():
            for k in idxs:
                results[k] = (130, "")
            continue
        views = [_job_view(jobs[k]) for k in idxs]
        try:
            with _gpu_slots, torch.cuda.stream(stream):
                if dev_image is not None:
                    dev = dev_image
                elif host.dtype == np.uint16:
                    dev = torch.from_numpy(host.view(np.int16)).cuda().view(torch.uint16)
                else:
                    dev = torch.from_numpy(host).cuda()
                with _stage("remap", stream):
                    out = api.remap_erp(dev[None], views, (w, h), interp=interp, convention=job_convention(), stream=stream)[0]
                encoded = {}
                host_side = {}                                   # view index -> future of a host-encoded file
                if jc is not None and out.dtype in (torch.uint8, torch.uint16) and out.shape[-1] in (1, 3):
                    # a share of the JPEG views goes to host threads (cv2.imwrite) while this thread's nvJPEG
                    # encoder takes the rest: the device encoder saturates at ~900 views/s whatever the number of
                    # threads feeding it, and the host has idle cores beside the decoders
                    share = _host_encode_share()
                    if share > 0 and out.dtype == torch.uint8 and out.shape[-1] == 3:
                        chosen = [n for n, k in enumerate(idxs) if _is_jpeg(jobs[k].output) and (n * share) % 100 + share > 99]
                        with _stage("download", stream):
                            # into pinned buffers of this worker, one synchronisation for all of them; the buffers
                            # are free again when the host encoders' futures have been collected below
                            staged = []
                            if os.environ.get("R360_STILL_PINNED", "1") == "0":          # experiment: pageable copies
                                staged = [out[n].contiguous().cpu() for n in chosen]
                            else:
                                for slot, n in enumerate(chosen):
                                    buf = worker.pinned_view(slot, out[n].shape)
                                    buf.copy_(out[n], non_blocking=True)
                                    staged.append(buf)
                                stream.synchronize()
                        for n, buf in zip(chosen, staged):
                            host_side[n] = _host_writers().submit(_write_image, jobs[idxs[n]].output, buf.numpy(), jobs[idxs[n]].jpeg_quality)
                    with _stage("encode", stream):
                        for n, k in enumerate(idxs):
                            if n in host_side:
                                continue
                            if _is_jpeg(jobs[k].output):
                                # 16-bit views are scaled to 8 bits on the device first (narrow_to_8bit)
                                encoded[n] = jc.encode(narrow_to_8bit(out[n]).contiguous() if out.dtype == torch.uint16 else out[n],
                                                       jobs[k].jpeg_quality, stream=stream)
                out_host = None
                if len(encoded) + len(host_side) < len(idxs):
                    with _stage("download", stream):
                        if out.dtype == torch.uint16:
                            out_host = out.contiguous().view(torch.int16).cpu().numpy().view(np.uint16)
                        else:
                            out_host = out.contiguous().cpu().numpy()
                stream.synchronize()
            with _stage("write"):
                for n, k in enumerate(idxs):
                    if n in host_side:
                        continue
                    if n in encoded:
                        jobs[k].output.parent.mkdir(parents=True, exist_ok=True)
                        jobs[k].output.write_bytes(encoded[n])
                    else:
                        img = out_host[n]
                        _write_image(jobs[k].output, img[..., 0] if img.shape[2] == 1 else img, jobs[k].jpeg_quality)
                    results[k] = (0, "")
            with _stage("host_encode_wait"):
                for n, fut in host_side.items():
                    fut.result()
                    results[idxs[n]] = (0, "")
        except Exception as exc:  # report per job, like a failing ffmpeg process would
            text = "%s: %s" % (type(exc).__name__, exc)
            for k in idxs:
                if results[k] is None:
                    results[k] = (1, text)
    return [r if r is not None else (1, "job skipped") for r in results]


def run_job_argv(cmd: Sequence[str], stop_event=None) -> Tuple[int, str]:
    """One job (the ``run_one`` contract): returns (rc, stderr_text)."""
    try:
        job = parse_job_argv(cmd)
    except Exception as exc:
        return 1, str(exc)
    if job.video:
        from . import video
        return video.run_video_jobs(job.source, [job], stop_event)[0]
    return _run_still_group(job.source, [job], stop_event)[0]


def run_jobs(jobs, stop_event=None, workers: int = 1) -> Iterator[Tuple[tuple, Tuple[int, str]]]:
    """Run planner jobs grouped by source; yields (job, (rc, err)) for every job as groups finish."""
    groups: "OrderedDict[str, List[int]]" = OrderedDict()
    parsed: List[Optional[ParsedJob]] = []
    early: Dict[int, Tuple[int, str]] = {}
    for n, (cmd, _src, _dst) in enumerate(jobs):
        try:
            pj = parse_job_argv(cmd)
            parsed.append(pj)
            groups.setdefault(str(pj.source) + ("|video" if pj.video else ""), []).append(n)
        except Exception as exc:
            parsed.append(None)
            early[n] = (1, str(exc))
    for n, res in early.items():
        yield jobs[n], res

    def run_group(idxs: List[int]):
        if stop_event is not None and stop_event.is_set():
            return idxs, [(130, "")] * len(idxs)
        pjs = [parsed[n] for n in idxs]
        try:
            if pjs[0].video:
                from . import video
                return idxs, video.run_video_jobs(pjs[0].source, pjs, stop_event)
            return idxs, _run_still_group(pjs[0].source, pjs, stop_event)
        except Exception:
            return idxs, [(1, traceback.format_exc())] * len(idxs)

    with ThreadPoolExecutor(max_workers=max(1, workers)) as pool:
        for idxs, results in pool.map(run_group, list(groups.values())):
            for n, res in zip(idxs, results):
                yield jobs[n], res
