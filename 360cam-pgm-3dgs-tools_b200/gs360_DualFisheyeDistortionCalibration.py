#!/usr/bin/env python3
"""Drop-in for cli_tools/gs360_DualFisheyeDistortionCalibration.py of the reference: same flags,
messages, exit codes and output layout; colour pipeline, undistortion and the ten perspective views
run on the GPU (see remap360/dualfisheye_cli.py).

    python gs360_DualFisheyeDistortionCalibration.py -i <frames_dir> [--save-fisheye-output] [--input-lut X.cube]
"""

import pathlib
import sys

sys.path.insert(0, str(pathlib.Path(__file__).resolve().parent))

from remap360.dualfisheye import (SensorCalibration, build_sfm10_specs, compute_view_fov_deg,  # noqa: F401,E402
                                  estimate_auto_undistort_zoom, load_metashape_calibration,
                                  parse_sensor_dimensions, wrap_angle_deg)
from remap360.dualfisheye_cli import create_arg_parser, main, parse_undistort_zoom_arg  # noqa: F401,E402

if __name__ == "__main__":
    sys.exit(main())
