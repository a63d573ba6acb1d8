"""remap360: B200-native panorama / dual-fisheye -> perspective view generation.

Host layer (Python + PyTorch for device memory and streams) over the hand-written sm_100a
kernels in ../csrc, reached through the C ABI declared in include/remap360.h."""

from .api import (FisheyeCalibration, PerspectiveView, UndistortItem, alloc_views, remap_erp,  # noqa: F401
                  remap_fisheye, sample_coordinates, undistort_fisheye)
from ._lib import Remap360Error, launch_count  # noqa: F401

__all__ = ["FisheyeCalibration", "PerspectiveView", "UndistortItem", "remap_erp", "remap_fisheye",
           "undistort_fisheye", "sample_coordinates",
           "Remap360Error", "launch_count"]
