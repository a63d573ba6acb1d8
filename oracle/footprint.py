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


def taps_footprint(shape, maps, interp, wrap):
    """Distinct source pixels touched by the taps of the given (map_x, map_y[, valid]) maps over one source image of
    `shape` = (H, W); columns wrap (panorama) or taps outside the image are dropped (constant border)."""
    H, W = shape
    touched = np.zeros((H, W), dtype=bool)
    k = 2 if interp == "linear" else 4
    off = k // 2 - 1
    for m in maps:
        mx, my = m[0], m[1]
        ix, _ = sampler.quantise(mx)
        iy, _ = sampler.quantise(my)
        if len(m) > 2:                      # invalid pixels are filled, but cv2.remap has sampled them all the same
            pass
        for ky in range(k):
            for kx in range(k):
                yy, xx = iy + ky - off, ix + kx - off
                if wrap:
                    touched[np.clip(yy, 0, H - 1), np.mod(xx, W)] = True
                else:
                    ok = (yy >= 0) & (yy < H) & (xx >= 0) & (xx < W)
                    touched[yy[ok], xx[ok]] = True
    return int(touched.sum())


def dualfisheye_footprint(interp, size=1750):
    """Config 5: the template calibration, SFM10 layout, both lenses (tests/golden/dualfisheye.json)."""
    import pathlib
    meta = json.loads((pathlib.Path(__file__).resolve().parent.parent / "tests" / "golden" / "dualfisheye.json").read_text())
    cal = meta["sensors"]["0"]
    specs = [dict(s, width=size, height=size) for s in meta["sfm10_default"]]
    views = geo.dualfisheye_view_maps(cal, cal, specs)
    per_lens = {"X": [], "Y": []}
    for v in views.values():
        per_lens[v["lens_key"]].append((v["map_x"], v["map_y"]))
    shape = (int(cal["height"]), int(cal["width"]))
    return sum(taps_footprint(shape, maps, interp, wrap=False) for maps in per_lens.values()), len(specs) * size * size


HARD_VIEWS = [(180.0, 0.0), (179.9, 0.0), (-179.9, 0.0), (0.0, 90.0), (0.0, -90.0), (40.0, 60.0), (-40.0, -60.0)]


def main():
    rows = {}
    hfov, views = PRESETS["full360coverage"]
    for interp in ("linear", "cubic"):
        u, sum_u, o = erp_footprint(7680, 3840, 1600, hfov, list(views) + HARD_VIEWS, interp)
        rows["full360coverage+seam/pole_7680x3840_%s" % interp] = {"U_px": u, "sumU_px": sum_u, "O_px": o}
        u5, o5 = dualfisheye_footprint(interp)
        rows["dualfisheye_sfm10_1750_%s" % interp] = {"U_px": u5, "O_px": o5}
        print("extra", interp, u, o, u5, o5, file=sys.stderr)
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
