"""JPEG decode / encode on the GPU (libr360codec.so, a thin C ABI over NVIDIA nvJPEG -- library code).

The reference decodes the source again in every ffmpeg process and encodes every view with ffmpeg's
mjpeg encoder (gs360_360PerspCut.py:317-339); the dual-fisheye tool uses cv2.imread / cv2.imwrite
(DF:735, :1826-1840).  With the remap at several thousand frames per second those two CPU steps are what
a real run waits for, so the job runners use this codec for ``.jpg`` files when it is available and fall
back to OpenCV's (identical file format, different encoder) when it is not.

    codec = JpegCodec()
    frame = codec.decode(path.read_bytes())                # [H, W, 3] uint8 CUDA tensor, BGR like cv2
    data = codec.encode(view, quality=95)                  # bytes, 4:4:4
"""

from __future__ import annotations

import ctypes
import pathlib
import threading
from ctypes import POINTER, c_char_p, c_int32, c_size_t, c_void_p
from typing import Optional

import torch

from . import _lib
from .api import _describe, _stream_handle

CODEC_PATH = pathlib.Path(__file__).resolve().parent / "libr360codec.so"
EXPORTS = ("r360_jpeg_create", "r360_jpeg_destroy", "r360_codec_last_error", "r360_jpeg_info",
           "r360_jpeg_decode", "r360_jpeg_encode", "r360_jpeg_retrieve")
_codec_lib = None
_tls = threading.local()


class CodecError(RuntimeError):
    pass


def load() -> ctypes.CDLL:
    global _codec_lib
    if _codec_lib is not None:
        return _codec_lib
    if not CODEC_PATH.exists():
        raise ImportError("%s is missing: build it with `python 360cam-pgm-3dgs-tools_b200/build.py`" % CODEC_PATH)
    lib = ctypes.CDLL(str(CODEC_PATH))
    lib.r360_jpeg_create.argtypes = [POINTER(c_void_p)]
    lib.r360_jpeg_destroy.argtypes = [c_void_p]
    lib.r360_jpeg_destroy.restype = None
    lib.r360_codec_last_error.restype = c_char_p
    lib.r360_jpeg_info.argtypes = [c_void_p, c_char_p, c_size_t, POINTER(c_int32), POINTER(c_int32), POINTER(c_int32)]
    lib.r360_jpeg_decode.argtypes = [c_void_p, c_char_p, c_size_t, POINTER(_lib.Images), c_int32, c_int32, c_void_p]
    lib.r360_jpeg_encode.argtypes = [c_void_p, POINTER(_lib.Images), c_int32, c_int32, c_int32, POINTER(c_size_t), c_void_p]
    lib.r360_jpeg_retrieve.argtypes = [c_void_p, c_void_p, c_size_t, POINTER(c_size_t), c_void_p]
    _codec_lib = lib
    return lib


def _check(rc: int) -> None:
    if rc != 0:
        detail = load().r360_codec_last_error().decode() if rc == -6 else _lib.load().r360_error_string(rc).decode()
        raise CodecError("remap360 codec error %d: %s" % (rc, detail))


class JpegCodec:
    """One nvJPEG decoder + encoder.  Not thread-safe: use ``JpegCodec.for_thread()``."""

    def __init__(self, device="cuda"):
        self.device = torch.device(device)
        if self.device.index is None:
            self.device = torch.device("cuda", torch.cuda.current_device())
        self._handle = c_void_p()
        with torch.cuda.device(self.device):
            _check(load().r360_jpeg_create(ctypes.byref(self._handle)))

    @classmethod
    def for_thread(cls, device="cuda") -> "JpegCodec":
        key = "codec_%s" % (torch.device(device),)
        if getattr(_tls, key, None) is None:
            setattr(_tls, key, cls(device))
        return getattr(_tls, key)

    def info(self, data: bytes):
        w, h, c = c_int32(), c_int32(), c_int32()
        _check(load().r360_jpeg_info(self._handle, data, len(data), ctypes.byref(w), ctypes.byref(h), ctypes.byref(c)))
        return w.value, h.value, c.value

    def decode(self, data: bytes, *, channel_order: str = "bgr", out: Optional[torch.Tensor] = None,
               stream: Optional[torch.cuda.Stream] = None) -> torch.Tensor:
        """JPEG bytes -> [H, W, C] uint8 CUDA tensor (C = 3, or 1 for greyscale files)."""
        w, h, c = self.info(data)
        if out is None:
            out = torch.empty((h, w, c), dtype=torch.uint8, device=self.device)
        elif tuple(out.shape) != (h, w, c) or out.dtype != torch.uint8:
            raise ValueError("out must be %s uint8" % ((h, w, c),))
        desc = _describe(out[None], "out")
        with torch.cuda.device(self.device):
            _check(load().r360_jpeg_decode(self._handle, data, len(data), ctypes.byref(desc), 0,
                                           1 if channel_order == "rgb" else 0, _stream_handle(stream, self.device)))
        return out

    def encode(self, image: torch.Tensor, quality: int = 95, *, channel_order: str = "bgr",
               stream: Optional[torch.cuda.Stream] = None) -> bytes:
        """[H, W, C] uint8 CUDA tensor (rows may be padded) -> JPEG bytes (4:4:4, optimised Huffman)."""
        if image.dim() != 3 or image.dtype != torch.uint8 or image.shape[2] not in (1, 3):
            raise ValueError("image must be [H, W, 1 or 3] uint8")
        desc = _describe(image[None], "image")
        size = c_size_t()
        handle = _stream_handle(stream, self.device)
        with torch.cuda.device(self.device):
            _check(load().r360_jpeg_encode(self._handle, ctypes.byref(desc), 0, int(quality),
                                           1 if channel_order == "rgb" else 0, ctypes.byref(size), handle))
            buf = ctypes.create_string_buffer(size.value)
            got = c_size_t()
            _check(load().r360_jpeg_retrieve(self._handle, buf, size.value, ctypes.byref(got), handle))
        return buf.raw[:got.value]

    def __del__(self):
        try:
            if self._handle:
                load().r360_jpeg_destroy(self._handle)
                self._handle = c_void_p()
        except Exception:
            pass


def available() -> bool:
    """True when the codec library is present and a codec object can be created on the current device."""
    try:
        JpegCodec.for_thread()
        return True
    except Exception:
        return False
