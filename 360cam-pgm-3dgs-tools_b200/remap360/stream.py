"""Frame streaming: decoded frames in pinned host memory -> H2D -> remap kernel -> D2H -> pinned
host views, on three CUDA streams with a ring of device buffers, so that the copies of the
batches before and after overlap the kernel of the current one.

The reference pays one full decode per (source, view) and moves nothing to a device
(cli_tools/gs360_360PerspCut.py:569-590, one ffmpeg process per job); here a frame is uploaded
once and all of its views are cut from the copy in HBM.  Frames travel in small batches (default 2):
consecutive frames of a video share their map, and the tiled kernel samples the frames of a batch with
one set of coordinates and weights (r360_tiled.cuh)."""

from __future__ import annotations

from collections import deque
from typing import Callable, Deque, Iterable, Iterator, List, Optional, Sequence, Tuple

import torch

from .api import PerspectiveView, alloc_views, remap_erp


class _Slot:
    def __init__(self, batch, frame_shape, dtype, n_views, size, out_dtype, device):
        h, w, c = frame_shape
        self.host_in = torch.empty((batch, h, w, c), dtype=dtype).pin_memory()
        self.dev_in = torch.empty((batch, h, w, c), dtype=dtype, device=device)
        self.dev_out = alloc_views(batch, n_views, size[1], size[0], c, out_dtype, device)
        self.host_out = torch.empty((batch, n_views, size[1], size[0], c), dtype=out_dtype).pin_memory()
        self.ev_h2d = torch.cuda.Event()
        self.ev_kernel = torch.cuda.Event()
        self.ev_d2h = torch.cuda.Event()
        self.n = 0                                   # frames in the slot


class StreamingRemapper:
    """Cut every incoming ERP frame into ``views``; results come back in submission order.

        remapper = StreamingRemapper(views, (1600, 1600), (3840, 7680, 3), torch.uint8)
        for views_host in remapper.run(frames):      # frames: iterable of CPU tensors [H, W, C]
            ...                                       # [V, h, w, C] pinned; valid until the next item
    Frames may also be CUDA tensors that are complete on their producer's stream (frames decoded on the device).

    ``depth`` batches of ``batch`` frames are in flight at once.  Default 6 x 1: the host link, not the kernel, bounds
    this path (the kernel is ten times faster), and single-frame batches interleave the two copy directions most
    finely -- measured on B200, 16 x 8K frames per call -> 12 views: 15.87 Gpix/s = 0.96 of the link with 6 x 1, 15.18
    with 4 x 2, 13.2 with 4 x 4, 11.5 with 4 x 8 (profiles/README.md).  Callers whose frames share a long stream and
    who want the kernel's multi-frame items pass batch=2."""

    def __init__(self, views: Sequence[PerspectiveView], size: Tuple[int, int], frame_shape: Tuple[int, int, int],
                 dtype: torch.dtype = torch.uint8, *, interp: str = "cubic", convention: str = "halfpixel",
                 out_dtype: Optional[torch.dtype] = None, device=None, depth: int = 6, batch: int = 1, path: str = "auto",
                 frame_filter: Optional[Callable[[torch.Tensor, torch.cuda.Stream], None]] = None, hold: int = 0):
        import os
        depth = int(os.environ.get("R360_STREAM_DEPTH", depth))      # experiments: ring shape without code changes
        batch = int(os.environ.get("R360_STREAM_BATCH", batch))
        if depth < 1 or batch < 1:
            raise ValueError("depth and batch must be >= 1")
        if hold < 0 or hold > depth - 2 and hold > 0:
            raise ValueError("hold must leave at least two slots of the ring in circulation")
        self.views = list(views)
        self.size = (int(size[0]), int(size[1]))
        self.interp, self.convention, self.path = interp, convention, path
        self.device = torch.device(device if device is not None else "cuda")
        self.out_dtype = out_dtype or dtype
        self.depth, self.batch = int(depth), int(batch)
        # results stay valid until `hold` further batches have been handed out (0: until the generator is advanced):
        # lets a caller copy / encode a result on other threads while it pulls the next ones
        self.hold = int(hold)
        # in-place per-frame step on the uploaded frames [n, H, W, C], run on the kernel stream before the
        # remap (the cutter's video colour step, PC:299-309)
        self.frame_filter = frame_filter
        with torch.cuda.device(self.device):
            self._free: Deque[_Slot] = deque(_Slot(self.batch, frame_shape, dtype, len(self.views), self.size, self.out_dtype, self.device)
                                       for _ in range(self.depth))
            self.s_h2d = torch.cuda.Stream(self.device)
            self.s_kernel = torch.cuda.Stream(self.device)
            self.s_d2h = torch.cuda.Stream(self.device)
        self._inflight: Deque[_Slot] = deque()
        self.frames_done = 0

    # -- pipeline stages ---------------------------------------------------------------------------
    def _upload(self, slot: _Slot, frame: torch.Tensor) -> None:
        """Frame -> slot.dev_in[slot.n], asynchronously on the upload stream."""
        b = slot.n
        src = frame
        if isinstance(frame, torch.Tensor) and frame.is_cuda:
            pass                                             # decoded on the device (nvJPEG): a device-to-device copy
        elif not (isinstance(frame, torch.Tensor) and frame.is_pinned()):
            slot.host_in[b].copy_(torch.as_tensor(frame))    # pageable input: stage through the pinned buffer
            src = slot.host_in[b]
        with torch.cuda.stream(self.s_h2d):
            slot.dev_in[b].copy_(src, non_blocking=True)
            if src.is_cuda:
                src.record_stream(self.s_h2d)
        slot.n = b + 1

    def _launch(self, slot: _Slot) -> None:
        n = slot.n
        slot.ev_h2d.record(self.s_h2d)
        self.s_kernel.wait_event(slot.ev_h2d)
        if self.frame_filter is not None:
            self.frame_filter(slot.dev_in[:n], self.s_kernel)
        remap_erp(slot.dev_in[:n], self.views, self.size, interp=self.interp, convention=self.convention,
                  out=slot.dev_out[:n], out_dtype=self.out_dtype, path=self.path, stream=self.s_kernel)
        slot.ev_kernel.record(self.s_kernel)
        self.s_d2h.wait_event(slot.ev_kernel)
        with torch.cuda.stream(self.s_d2h):
            slot.host_out[:n].copy_(slot.dev_out[:n], non_blocking=True)
            slot.ev_d2h.record(self.s_d2h)
        self._inflight.append(slot)

    def _drain_one(self) -> Iterator[torch.Tensor]:
        slot = self._inflight.popleft()
        slot.ev_d2h.synchronize()
        for b in range(slot.n):
            self.frames_done += 1
            yield slot.host_out[b]
        slot.n = 0
        self._free.append(slot)          # only after the caller has moved past the slot's last result

    def run(self, frames: Iterable[torch.Tensor]) -> Iterator[torch.Tensor]:
        """Generator over results, in order.  A yielded tensor is a view of a pinned ring buffer: it
        stays valid until the generator is advanced again (with ``hold`` = h: until h more batches have been
        handed out after the batch it belongs to)."""
        cur: Optional[_Slot] = None
        for frame in frames:
            if cur is None:
                # the oldest drained slot is reused, and only once `hold` younger ones lie between it and the caller
                while len(self._free) <= self.hold and self._inflight:
                    yield from self._drain_one()
                cur = self._free.popleft()
                cur.n = 0
            self._upload(cur, frame)
            if cur.n == self.batch:
                self._launch(cur)
                cur = None
        if cur is not None and cur.n:
            self._launch(cur)
        elif cur is not None:
            self._free.append(cur)
        while self._inflight:
            yield from self._drain_one()
