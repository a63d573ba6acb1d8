"""Algorithmic source footprint of a view set (TEST/MEASUREMENT INFRASTRUCTURE).

U = number of distinct source pixels touched by any tap of any view of one frame; together with
the output pixel count O it gives the algorithmic bytes per frame used for the roofline
(SURVEY.md section 8d):  bytes = (U + O) * channels * sizeof(sample).

    python -m oracle.footprint            # prints the table recorded in DESIGN.md / bench.py
"""

import json
import sys

import numpy as np

from . import geometry as geo
from . import sampler

PRESETS = {
    # name: (hfov, [(yaw, pitch), ...]) -- view sets of gs360_360PerspCut.py presets (tests/golden)
    "full360coverage": (104.2500326978036, [(0, 0), (45, 30), (45, -30), (90, 0), (135, 30), (135, -30), (180, 0),
                                            (-135, 30), (-135, -30), (-90, 0), (-45, 30), (-45, -30)]),
    "fisheyelike": (93.27315408323344, [(0, 0), (0, 30), (0, -30), (36, 0), (144, 0), (180, 0), (180, 30),
                                        (180, -30), (-144, 0), (-36, 0)]),
    "default": (112.61986494804043, [(y, 0) for y in (0, 45, 90, 135, 180, -135, -90, -45)]),
}


def erp_footprint(W, H, size, hfov, views, interp):
    touched = np.zeros((H, W), dtype=bool)
    per_view = 0
    k = 2 if interp == "linear" else 4
    off = k // 2 - 1
    for yaw, pitch in views:
        mx, my = geo.erp_map64(W, H, size, size, yaw, pitch, hfov, hfov)
        ix, _ = sampler.quantise(mx)
        iy, _ = sampler.quantise(my)
        one = np.zeros((H, W), dtype=bool)
        for ky in range(k):
            yy = np.clip(iy + ky - off, 0, H - 1)
            for kx in range(k):
                one[yy, np.mod(ix + kx - off, W)] = True
        per_view += int(one.sum())
        touched |= one
    return int(touched.sum()), per_view, len(views) * size * size


def main():
    rows = {}
    for name, W, H, size in (("full360coverage", 7680, 3840, 1600), ("fisheyelike", 7680, 3840, 1600),
                             ("default", 7680, 3840, 1600), ("default", 3840, 1920, 1600)):
        hfov, views = PRESETS[name]
        for interp in ("linear", "cubic"):
            u, sum_u, o = erp_footprint(W, H, size, hfov, views, interp)
            rows["%s_%dx%d_%s" % (name, W, H, interp)] = {"U_px": u, "sumU_px": sum_u, "O_px": o}
            print(name, W, H, interp, u, sum_u, o, file=sys.stderr)
    print(json.dumps(rows, indent=1))


if __name__ == "__main__":
    main()
