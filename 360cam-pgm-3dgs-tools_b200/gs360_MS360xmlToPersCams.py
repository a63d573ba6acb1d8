#!/usr/bin/env python3
"""Drop-in for cli_tools/gs360_MS360xmlToPersCams.py of the reference: Metashape spherical-camera XML -> virtual
perspective cameras of the cutter's presets (transforms.json / COLMAP text / RealityScan XMP / Metashape XML), same
flags, files and log lines; ``--persp-cut`` runs the CUDA cutter next to this file (see remap360/ms_export.py).

    python gs360_MS360xmlToPersCams.py cameras.xml [--preset full360coverage] [--format all --points-ply cloud.ply]"""

import pathlib
import sys

sys.path.insert(0, str(pathlib.Path(__file__).resolve().parent))

from remap360.ms_export import (PRESET_CHOICES, build_arg_parser, build_views, compute_intrinsics, main,  # noqa: F401,E402
                                preset_size_and_focal)

if __name__ == "__main__":
    main()
