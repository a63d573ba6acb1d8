#!/usr/bin/env python3
"""Drop-in for cli_tools/gs360_360PerspCut.py of the reference: same flags, presets, names and log
lines; the per-view ffmpeg processes are replaced by the CUDA remap (see remap360/perspcut.py).

    python gs360_360PerspCut.py -i <dir|video> [--preset default|fisheyelike|full360coverage|...] [-o OUT]

Importable under the reference's module name, e.g. ``import gs360_360PerspCut as cutter``
(gs360_GUI.py:54) or ``from gs360_360PerspCut import fov_from_focal_mm`` (gs360_Video2Frames.py:26)."""

import pathlib
import sys

sys.path.insert(0, str(pathlib.Path(__file__).resolve().parent))

from remap360.perspcut import *  # noqa: F401,F403
from remap360.perspcut import (BuildResult, PROGRESS_INTERVAL, StoreWithFlag, ViewSpec, build_ffmpeg_cmd,  # noqa: F401
                               build_ffmpeg_equisolid_cmd, build_view_jobs, create_arg_parser, main, on_signal,
                               parse_jobs, run_one, stop_event)

if __name__ == "__main__":
    main()
