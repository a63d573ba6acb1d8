"""Frame streaming: decoded frames in pinned host memory -> H2D -> remap kernel -> D2H -> pinned
host views, on three CUDA streams with a small ring of device buffers, so that the copies of
frame n+1 / n-1 overlap the kernel of frame n.

The reference pays one full decode per (source, view) and moves nothing to a device
(cli_tools/gs360_360PerspCut.py:569-590, one ffmpeg process per job); here a frame is uploaded
once and all of its views are cut from the copy in HBM."""

from __future__ import annotations

from collections import deque
from typing import Callable, Deque, Iterable, Iterator, List, Optional, Sequence, Tuple

import torch

from .api import PerspectiveView, alloc_views, remap_erp


class _Slot:
    def __init__(self, frame_shape, dtype, n_views, size, out_dtype, device):
        h, w, c = frame_shape
        self.host_in = torch.empty((h, w, c), dtype=dtype).pin_memory()
        self.dev_in = torch.empty((1, h, w, c), dtype=dtype, device=device)
        self.dev_out = alloc_views(1, n_views, size[1], size[0], c, out_dtype, device)
        self.host_out = torch.empty((n_views, size[1], size[0], c), dtype=out_dtype).pin_memory()
        self.ev_h2d = torch.cuda.Event()
        self.ev_kernel = torch.cuda.Event()
        self.ev_d2h = torch.cuda.Event()


class StreamingRemapper:
    """Cut every incoming ERP frame into ``views``; results come back in submission order.

        remapper = StreamingRemapper(views, (1600, 1600), (3840, 7680, 3), torch.uint8)
        for views_host in remapper.run(frames):      # frames: iterable of CPU tensors [H, W, C]
            ...                                       # [V, h, w, C] pinned; valid until the next item

    ``depth`` frames are in flight at once (default 3: one uploading, one in the kernel, one
    downloading)."""

    def __init__(self, views: Sequence[PerspectiveView], size: Tuple[int, int], frame_shape: Tuple[int, int, int],
                 dtype: torch.dtype = torch.uint8, *, interp: str = "cubic", convention: str = "halfpixel",
                 out_dtype: Optional[torch.dtype] = None, device=None, depth: int = 3, path: str = "auto",
                 frame_filter: Optional[Callable[[torch.Tensor, torch.cuda.Stream], None]] = None):
        if depth < 1:
            raise ValueError("depth must be >= 1")
        self.views = list(views)
        self.size = (int(size[0]), int(size[1]))
        self.interp, self.convention, self.path = interp, convention, path
        self.device = torch.device(device if device is not None else "cuda")
        self.out_dtype = out_dtype or dtype
        # in-place per-frame step on the uploaded frame [1, H, W, C], run on the kernel stream before the
        # remap (the cutter's video colour step, PC:299-309)
        self.frame_filter = frame_filter
        with torch.cuda.device(self.device):
            self._free: List[_Slot] = [_Slot(frame_shape, dtype, len(self.views), self.size, self.out_dtype, self.device)
                                       for _ in range(depth)]
            self.s_h2d = torch.cuda.Stream(self.device)
            self.s_kernel = torch.cuda.Stream(self.device)
            self.s_d2h = torch.cuda.Stream(self.device)
        self._inflight: Deque[_Slot] = deque()
        self.frames_done = 0

    # -- pipeline stages ---------------------------------------------------------------------------
    def _push(self, frame: torch.Tensor) -> None:
        slot = self._free.pop()
        src = frame
        if not (isinstance(frame, torch.Tensor) and frame.is_pinned()):
            slot.host_in.copy_(torch.as_tensor(frame))       # pageable input: stage through the pinned buffer
            src = slot.host_in
        with torch.cuda.stream(self.s_h2d):
            slot.dev_in[0].copy_(src, non_blocking=True)
            slot.ev_h2d.record(self.s_h2d)
        self.s_kernel.wait_event(slot.ev_h2d)
        if self.frame_filter is not None:
            self.frame_filter(slot.dev_in, self.s_kernel)
        remap_erp(slot.dev_in, self.views, self.size, interp=self.interp, convention=self.convention,
                  out=slot.dev_out, out_dtype=self.out_dtype, path=self.path, stream=self.s_kernel)
        slot.ev_kernel.record(self.s_kernel)
        self.s_d2h.wait_event(slot.ev_kernel)
        with torch.cuda.stream(self.s_d2h):
            slot.host_out.copy_(slot.dev_out[0], non_blocking=True)
            slot.ev_d2h.record(self.s_d2h)
        self._inflight.append(slot)

    def _pop(self) -> torch.Tensor:
        slot = self._inflight.popleft()
        slot.ev_d2h.synchronize()
        self._free.append(slot)
        self.frames_done += 1
        return slot.host_out

    def run(self, frames: Iterable[torch.Tensor]) -> Iterator[torch.Tensor]:
        """Generator over results, in order.  A yielded tensor is a view of a pinned ring buffer: it
        stays valid until the generator is advanced again."""
        for frame in frames:
            if not self._free:
                yield self._pop()
            self._push(frame)
        while self._inflight:
            yield self._pop()
