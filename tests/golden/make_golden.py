#!/usr/bin/env python3
"""Regenerate the fixtures in tests/golden/ by RUNNING THE REFERENCE.

Only works where the reference checkout exists (this build container):

    PYTHONDONTWRITEBYTECODE=1 python tests/golden/make_golden.py [/root/reference]

It imports the reference's own modules (nothing is copied) and records their
outputs; the tests then check the oracle and the product against those files,
so the reference does not need to exist where the tests run (the GPU box).

Files written
-------------
perspcut_views.json     ``gs360_360PerspCut.build_view_jobs`` for every preset and
                        a set of --addcam/--setcam/--delcam/--size/... variants
                        (jobs argv, names, ViewSpecs, FOVs, the four log lines)
perspcut_helpers.json   small pure helpers (fov_from_focal_mm, v_fov_from_hfov,
                        letter_tag, normalize_angle_deg, extra_suffix, parse_*)
dualfisheye.json        template calibration as parsed by the reference,
                        ``build_sfm10_specs`` output, lens choice + valid ratio
                        of ``build_perspective_spec_maps`` at 1750 px
dualfisheye_maps.npz    reference float32 maps: all views x both lenses at 96 px,
                        a strided sample (every 25th pixel) of the chosen-lens
                        maps at 1750 px, and 64 px maps for a synthetic
                        calibration with every Brown/affinity term non-zero
undistort.json          fisheye -> undistorted fisheye (DF:1008-1170): the auto zoom of
undistort_maps.npz      ``estimate_auto_undistort_zoom`` and strided ``build_remap_cache`` maps
                        + validity for the template calibration (auto and fixed zoom) and
                        the synthetic one; a small full-resolution case for cv2 parity
color_pipeline.npz      ``apply_input_color_pipeline`` (DF:684-725) on random uint8 / uint16 / 4-channel
                        images with two synthetic .cube LUTs (size 5 with a shifted domain, size 17),
                        passthrough and srgb; also the LUT text files as parsed by ``load_cube_lut``
df_cli.json             the DualFisheye CLI (DF:124-449, :2067-2853): every argparse action of
df_cli_outputs.npz      ``parse_arguments``; stdout / stderr / exit code of ``main()`` for dry runs and
                        usage errors (temp paths replaced by <TMP>); and one real run on two tiny
                        smooth X/Y pairs (LUT, undistorted fisheye, 10 views + masks as PNG) whose
                        input and output images are stored for the end-to-end parity test
ms_export.json          the MS360xmlToPersCams pose exporters on a synthetic spherical project: view sets and intrinsics
                        per preset, and whole ``main()`` runs (transforms.json, COLMAP text, RealityScan XMP, Metashape XML,
                        rotated point cloud; stdout / stderr / exit codes)
gui_geometry.npz        lon / lat and equirectangular pixel coordinates of 33 x 33 pixel centres of ten views,
                        computed by the reference's own ``direction_from_uv`` / ``lonlat_to_xy`` (gs360_GUI.py:342-424,
                        taken out of the file's syntax tree because the module needs tkinter to import)
df_metadata.json        pose / COLMAP / Metashape-XML export of the dual-fisheye tool on synthetic aligned projects:
                        cameras as loaded (chunk / component similarity), label pairs, pose frames, the COLMAP
                        model, the written files, and stdout / files of ``main()`` runs with --metadata-only
v2f_fisheye.json        the ``v360=<fisheye|equisolid>:rectilinear:...`` filter string and view FOVs that
                        gs360_Video2Frames.py builds for --fisheye-perspective (V2F:467-487)
cv2_remap.npz           ``cv2.remap`` outputs (the routine the reference calls,
                        DF:2001-2008) for small random sources x dtypes x
                        interpolations x borders
"""

import argparse
import json
import math
import pathlib
import sys

import numpy as np

HERE = pathlib.Path(__file__).resolve().parent


def _args_for(pc, argv, video=False, bit_depth=8):
    ap = pc.create_arg_parser()
    args = ap.parse_args(argv)
    for name in ("size", "hfov", "focal_mm"):
        setattr(args, name + "_explicit", getattr(args, name + "_explicit", False))
    args.input_is_video = video
    args.video_bit_depth = bit_depth
    return args


PERSPCUT_CASES = [
    # (name, argv after "-i IN", files, video, bit_depth)
    ("default", [], ["pano0001.jpg"], False, 8),
    ("fisheyelike", ["--preset", "fisheyelike"], ["pano0001.jpg"], False, 8),
    ("full360coverage", ["--preset", "full360coverage"], ["pano0001.jpg"], False, 8),
    ("2views", ["--preset", "2views"], ["pano0001.jpg"], False, 8),
    ("evenMinus30", ["--preset", "evenMinus30"], ["pano0001.jpg"], False, 8),
    ("evenPlus30", ["--preset", "evenPlus30"], ["pano0001.jpg"], False, 8),
    ("fisheyeXY", ["--preset", "fisheyeXY"], ["pano0001.jpg"], False, 8),
    ("fisheyeXY_sized", ["--preset", "fisheyeXY", "--size", "2000", "--hfov", "170"],
     ["pano0001.jpg"], False, 8),
    ("two_files_png", ["--ext", "png"], ["a.tif", "b.png"], False, 8),
    ("count4_topbottom", ["--count", "4", "--hfov", "105", "--add-top", "--add-bottom"],
     ["p.jpg"], False, 8),
    ("topdown_hidden", ["--add-topdown", "--count", "6"], ["p.jpg"], False, 8),
    ("addcam_mix", ["--addcam", "B,D:U,F:D20,H:U12.5", "--addcam-deg", "25"], ["p.jpg"], False, 8),
    ("delcam_setcam", ["--delcam", "B,d,8", "--setcam", "A=30,C:-10,E=U,G=D15"], ["p.jpg"], False, 8),
    ("setcam_extra", ["--addcam", "A,C", "--setcam", "A_U=50,A_D:+5,C:+10,C_U:-5"], ["p.jpg"], False, 8),
    ("setcam_clamp", ["--setcam", "A=120,B=-95"], ["p.jpg"], False, 8),
    ("count30", ["--count", "30", "--size", "800"], ["p.jpg"], False, 8),
    ("focal_sensor", ["--focal-mm", "18", "--sensor-mm", "36x24", "--size", "1200"], ["p.jpg"], False, 8),
    ("sensor_apsc", ["--focal-mm", "10", "--sensor-mm", "23.5 15.6"], ["p.jpg"], False, 8),
    ("hfov_explicit", ["--hfov", "90", "--size", "1024"], ["p.jpg"], False, 8),
    ("preset_override", ["--preset", "full360coverage", "--focal-mm", "16", "--size", "1800"],
     ["p.jpg"], False, 8),
    ("fisheyelike_user_del", ["--preset", "fisheyelike", "--delcam", "B", "--addcam", "C:U10"],
     ["p.jpg"], False, 8),
    ("jpeg95", ["--jpeg-quality-95"], ["p.jpg"], False, 8),
    ("video_jpg", ["--preset", "fisheyelike", "-f", "2"], ["vid.mp4"], True, 8),
    ("video_png10", ["--preset", "fisheyelike", "-f", "2", "--ext", "png"], ["vid.mp4"], True, 10),
    ("video_tif_keep709", ["-f", "0.5", "--ext", "tif", "--keep-rec709", "--start", "3", "--end", "12.5"],
     ["clip.mov"], True, 8),
    ("video_fisheyeXY", ["--preset", "fisheyeXY", "-f", "1"], ["vid.mp4"], True, 8),
]


def dump_perspcut(pc):
    out = {}
    for name, extra, files, video, depth in PERSPCUT_CASES:
        argv = ["-i", "/tmp/in"] + list(extra)
        args = _args_for(pc, argv, video, depth)
        paths = [pathlib.Path("/tmp/in") / f for f in files]
        res = pc.build_view_jobs(args, paths, pathlib.Path("/tmp/out"))
        out[name] = {
            "argv": argv, "files": files, "video": video, "bit_depth": depth,
            "jobs": [[list(cmd), src, dst] for cmd, src, dst in res.jobs],
            "view_specs": [
                {"source_path": str(v.source_path), "output_name": v.output_name,
                 "view_id": v.view_id, "yaw_deg": v.yaw_deg, "pitch_deg": v.pitch_deg,
                 "hfov_deg": v.hfov_deg, "vfov_deg": v.vfov_deg, "width": v.width,
                 "height": v.height, "projection": v.projection}
                for v in res.view_specs],
            "focal_used_mm": res.focal_used_mm, "focal_35mm_equiv": res.focal_35mm_equiv,
            "hfov_deg": res.hfov_deg, "vfov_deg": res.vfov_deg,
            "preview_views_line": res.preview_views_line, "sensor_line": res.sensor_line,
            "realityscan_line": res.realityscan_line, "metashape_line": res.metashape_line,
            "args_after": {k: getattr(args, k) for k in ("count", "size", "focal_mm", "add_top", "add_bottom")},
        }
    (HERE / "perspcut_views.json").write_text(json.dumps(out, indent=1, sort_keys=True) + "\n")

    helpers = {
        "fov_from_focal_mm": [[f, s, pc.fov_from_focal_mm(f, s)]
                              for f, s in ((12, 36), (14, 36), (17, 36), (6, 36), (10, 23.5), (50, 36))],
        "focal_from_hfov_deg": [[a, s, pc.focal_from_hfov_deg(a, s)] for a, s in ((90, 36), (105, 36), (60, 24))],
        "v_fov_from_hfov": [[a, w, h, pc.v_fov_from_hfov(a, w, h)] for a, w, h in ((90, 1600, 1600), (100, 1920, 1080), (70, 800, 1200))],
        "letter_tag": [[i, pc.letter_tag(i)] for i in (0, 1, 7, 25, 26, 27, 40)],
        "letter_to_index1": [[s, pc.letter_to_index1(s)] for s in ("A", "b", " h ", "12", "Z")],
        "normalize_angle_deg": [[a, pc.normalize_angle_deg(a)] for a in (0, 45, 180, -180, 181, 359.5, 540, -181, 179.9999999)],
        "extra_suffix": [[d, dd, pc.extra_suffix(d, dd)] for d, dd in ((30, 30), (-30, 30), (20, 30), (-12.5, 30), (25, 25), (-40, 25))],
        "parse_jobs": [[s, pc.parse_jobs(s)] for s in ("1", "7", "0", "-3")],
        "parse_sensor": [[s, pc.parse_sensor(s)] for s in ("36 36", "36x24", "23.5,15.6", "24")],
        "parse_addcam_spec": [[s, d, {str(k): v for k, v in pc.parse_addcam_spec(s, d).items()}]
                              for s, d in (("B", 30.0), ("B:U,D:D20", 30.0), ("A=u15, C", 20.0), ("", 30.0))],
        "parse_delcam_spec": [[s, sorted(pc.parse_delcam_spec(s))] for s in ("B,D", "a, 3 ,H", "")],
    }
    setcam = []
    for s, d in (("A=30,B:-5", 30.0), ("A=U,B=D20,C:U15", 30.0), ("A_U=5,A_D:+5", 30.0), ("", 30.0)):
        a, b, c, e = pc.parse_setcam_spec(s, d)
        setcam.append([s, d, {str(k): v for k, v in a.items()}, {str(k): v for k, v in b.items()},
                       {"%d%s" % k: v for k, v in c.items()}, {"%d%s" % k: v for k, v in e.items()}])
    helpers["parse_setcam_spec"] = setcam
    (HERE / "perspcut_helpers.json").write_text(json.dumps(helpers, indent=1, sort_keys=True) + "\n")


def _calib_dict(c):
    return {k: getattr(c, k) for k in ("sensor_id", "model_type", "width", "height", "f", "cx", "cy",
                                       "k1", "k2", "k3", "k4", "p1", "p2", "b1", "b2")}


def dump_dualfisheye(df):
    sensor_map, cam_to_sensor = df.load_metashape_calibration(df.DEFAULT_CAMERA_XML)
    calib = sensor_map[sorted(sensor_map)[0]]
    meta = {
        "template_xml": "cli_tools/templates/Osmo360-Fisheye-Distortion.xml",
        "sensors": {k: _calib_dict(v) for k, v in sensor_map.items()},
        "n_camera_labels": len(cam_to_sensor),
        "camera_to_sensor_sample": dict(sorted(cam_to_sensor.items())[:6]),
        "compute_view_fov_deg": [[f, s, list(df.compute_view_fov_deg(f, s))]
                                 for f, s in ((14.0, "36 36"), (12.0, "36x24"), (0.05, "36 36"), (4000.0, "36 36"))],
        "wrap_angle_deg": [[a, df.wrap_angle_deg(a)] for a in (0, 180, -180, 220, 320, 539, -181)],
    }
    specs = df.build_sfm10_specs(1750, 14.0, "36 36", 40.0, 40.0)
    meta["sfm10_default"] = specs
    meta["sfm10_alt"] = df.build_sfm10_specs(1200, 12.0, "36x24", 35.0, 25.0)

    arrays = {}
    # (1) default layout at 1750 px: lens choice, valid ratio, strided sample
    maps = df.build_perspective_spec_maps(sensor_map, calib.sensor_id, calib.sensor_id, specs, 0.0, 180.0, 190.0)
    stride = 25
    meta["maps_1750"] = {"stride": stride, "views": {}}
    for vid, m in maps.items():
        meta["maps_1750"]["views"][vid] = {"lens_key": m["lens_key"], "valid_ratio": float(np.mean(m["valid"]))}
        arrays["s1750_%s_x" % vid] = m["map_x"][::stride, ::stride].copy()
        arrays["s1750_%s_y" % vid] = m["map_y"][::stride, ::stride].copy()
        arrays["s1750_%s_v" % vid] = m["valid"][::stride, ::stride].copy()
    # (2) every view x both lenses at 96 px, full maps (includes invalid regions)
    small = df.build_sfm10_specs(96, 14.0, "36 36", 40.0, 40.0)
    for spec in small:
        for lens_key, lens_yaw in (("X", 0.0), ("Y", 180.0)):
            yaw_rel = df.wrap_angle_deg(spec["yaw_deg"] - lens_yaw)
            mx, my, valid = df.build_direct_perspective_map_for_lens(
                calib, yaw_rel, spec["pitch_deg"], spec["hfov_deg"], spec["vfov_deg"], 96, 96, 190.0)
            key = "m96_%s_%s" % (spec["view_id"], lens_key)
            arrays[key + "_x"], arrays[key + "_y"], arrays[key + "_v"] = mx, my, valid
    # (3) synthetic calibration exercising k4, p1, p2, b1, b2 and a non-square sensor
    syn = df.SensorCalibration(sensor_id="9", model_type="equisolid_fisheye", width=3000, height=2800,
                               f=820.5, cx=12.25, cy=-7.5, k1=0.08, k2=-0.011, k3=0.0021, k4=-0.0003,
                               p1=0.0007, p2=-0.0004, b1=1.75, b2=-0.6)
    meta["synthetic_calibration"] = _calib_dict(syn)
    syn_views = [(0.0, 0.0), (35.0, 20.0), (-60.0, -35.0), (95.0, 5.0)]
    meta["synthetic_views"] = syn_views
    for n, (yaw, pitch) in enumerate(syn_views):
        mx, my, valid = df.build_direct_perspective_map_for_lens(syn, yaw, pitch, 100.0, 80.0, 80, 64, 185.0)
        arrays["syn%d_x" % n], arrays["syn%d_y" % n], arrays["syn%d_v" % n] = mx, my, valid
    np.savez_compressed(HERE / "dualfisheye_maps.npz", **arrays)
    (HERE / "dualfisheye.json").write_text(json.dumps(meta, indent=1, sort_keys=True) + "\n")


def dump_undistort(df):
    sensor_map, _ = df.load_metashape_calibration(df.DEFAULT_CAMERA_XML)
    tmpl = sensor_map[sorted(sensor_map)[0]]
    syn = df.SensorCalibration(sensor_id="9", model_type="equisolid_fisheye", width=3000, height=2800,
                               f=820.5, cx=12.25, cy=-7.5, k1=0.08, k2=-0.011, k3=0.0021, k4=-0.0003,
                               p1=0.0007, p2=-0.0004, b1=1.75, b2=-0.6)
    tiny = df.SensorCalibration(sensor_id="7", model_type="equisolid_fisheye", width=320, height=288,
                                f=88.0, cx=1.5, cy=-2.25, k1=0.05, k2=-0.008, k3=0.001, k4=0.0,
                                p1=0.0005, p2=-0.0003, b1=0.4, b2=-0.2)
    wide = df.SensorCalibration(sensor_id="8", model_type="equisolid_fisheye", width=320, height=288,
                                f=118.0, cx=-3.0, cy=2.0, k1=0.03, k2=0.004, k3=0.0, k4=0.0,
                                p1=0.0, p2=0.0, b1=0.0, b2=0.0)      # image circle larger than the sensor
    meta, arrays = {"cases": {}}, {}
    cases = [("wide_auto", wide, None, 190.0, 2), ("wide_auto_fov120", wide, None, 120.0, 2),("tmpl_auto", tmpl, None, 190.0, 32), ("tmpl_z1", tmpl, 1.0, 190.0, 32),
             ("tmpl_z125_fov170", tmpl, 1.25, 170.0, 32), ("syn_auto", syn, None, 185.0, 25),
             ("tiny_auto", tiny, None, 190.0, 1), ("tiny_z09_fov150", tiny, 0.9, 150.0, 1)]
    for name, cal, zoom, fov, stride in cases:
        rc = df.build_remap_cache(cal, zoom, fov)
        meta["cases"][name] = {"calibration": _calib_dict(cal), "zoom_arg": zoom, "lens_fov_deg": fov,
                               "undistort_zoom": float(rc.undistort_zoom), "stride": stride,
                               "valid_ratio": float(np.mean(rc.valid_mask))}
        arrays[name + "_x"] = rc.map_x[::stride, ::stride].copy()
        arrays[name + "_y"] = rc.map_y[::stride, ::stride].copy()
        arrays[name + "_v"] = rc.valid_mask[::stride, ::stride].copy()
    meta["auto_zoom"] = {"tmpl_190": float(df.estimate_auto_undistort_zoom(tmpl, lens_fov_deg=190.0)),
                         "tmpl_150": float(df.estimate_auto_undistort_zoom(tmpl, lens_fov_deg=150.0)),
                         "syn_185": float(df.estimate_auto_undistort_zoom(syn, lens_fov_deg=185.0)),
                         "wide_190": float(df.estimate_auto_undistort_zoom(wide, lens_fov_deg=190.0)),
                         "wide_120": float(df.estimate_auto_undistort_zoom(wide, lens_fov_deg=120.0)),
                         "wide_190_n48": float(df.estimate_auto_undistort_zoom(wide, sample_count=48, lens_fov_deg=190.0)),
                         "tiny_190": float(df.estimate_auto_undistort_zoom(tiny, lens_fov_deg=190.0)),
                         "tiny_190_n64": float(df.estimate_auto_undistort_zoom(tiny, sample_count=64, lens_fov_deg=190.0))}
    np.savez_compressed(HERE / "undistort_maps.npz", **arrays)
    (HERE / "undistort.json").write_text(json.dumps(meta, indent=1, sort_keys=True) + "\n")


def _write_cube(path, size, rng, domain=None, title=True):
    lines = []
    if title:
        lines += ['TITLE "synthetic %d"' % size, "# generated by tests/golden/make_golden.py"]
    lines.append("LUT_3D_SIZE %d" % size)
    if domain is not None:
        lines.append("DOMAIN_MIN %g %g %g" % tuple(domain[0]))
        lines.append("DOMAIN_MAX %g %g %g" % tuple(domain[1]))
    grid = np.linspace(0.0, 1.0, size)
    for b in grid:
        for g in grid:
            for r in grid:                                    # red fastest, like every .cube file
                v = np.array([r ** 0.8 + 0.05 * g, 0.9 * g + 0.1 * b * b, b ** 1.3 - 0.04 * r]) + rng.normal(0, 0.01, 3)
                lines.append("%.6f %.6f %.6f" % tuple(v))
    path.write_text("\n".join(lines) + "\n")


def dump_color(df):
    import tempfile
    rng = np.random.default_rng(4242)
    arrays = {}
    with tempfile.TemporaryDirectory() as tmp:
        tmp = pathlib.Path(tmp)
        luts = {"s5": (5, ((-0.05, 0.0, 0.02), (1.1, 0.95, 1.0))), "s17": (17, None)}
        for name, (size, domain) in luts.items():
            _write_cube(tmp / (name + ".cube"), size, rng, domain)
            arrays["cube_%s_text" % name] = np.array((tmp / (name + ".cube")).read_text())
            lut = df.load_cube_lut(tmp / (name + ".cube"))
            arrays["cube_%s_table" % name] = lut.table
            arrays["cube_%s_min" % name], arrays["cube_%s_max" % name] = lut.domain_min, lut.domain_max
            for dt, ch in (("uint8", 3), ("uint16", 3), ("uint8", 4)):
                img = rng.integers(0, np.iinfo(dt).max + 1, (48, 64, ch)).astype(dt)
                if dt == "uint8" and ch == 3:                 # every code value of one channel, and the knees
                    img[0, :, :] = np.arange(64)[:, None] * 4
                    img[1, :, 0] = np.arange(64) * 4 + 3
                arrays["img_%s_c%d_%s" % (dt, ch, name)] = img
                for space in ("passthrough", "srgb"):
                    arrays["out_%s_c%d_%s_%s" % (dt, ch, name, space)] = df.apply_input_color_pipeline(img, lut, space)
    v = np.linspace(0, 1, 4097).astype(np.float32)
    arrays["rec709_to_srgb_in"], arrays["rec709_to_srgb_out"] = v, df.rec709_to_srgb(v)
    np.savez_compressed(HERE / "color_pipeline.npz", **arrays)


TINY_XML = """<?xml version="1.0" encoding="UTF-8"?>
<document version="2.3.0"><chunk label="Chunk 1" enabled="true"><sensors next_id="1">
<sensor id="0" label="unknown" type="equisolid_fisheye"><resolution width="{w}" height="{h}"/>
<calibration type="equisolid_fisheye" class="adjusted"><resolution width="{w}" height="{h}"/>
<f>{f}</f><cx>0.4</cx><cy>-0.3</cy><k1>0.06</k1><k2>-0.004</k2><k3>0.0007</k3></calibration>
</sensor></sensors></chunk></document>
"""


def _smooth_image(rng, h, w, dtype=np.uint8):
    yy, xx = np.mgrid[0:h, 0:w].astype(np.float64)
    img = np.empty((h, w, 3))
    for c in range(3):
        a = rng.uniform(0.5, 2.5, 4)
        ph = rng.uniform(0, 6.28, 4)
        img[..., c] = (0.5 + 0.2 * np.sin(a[0] * xx / w * 6.28 + ph[0]) * np.cos(a[1] * yy / h * 6.28 + ph[1])
                       + 0.15 * np.sin(a[2] * (xx + yy) / (w + h) * 6.28 + ph[2]) + 0.1 * np.cos(a[3] * yy / h * 3.14 + ph[3]))
    top = np.iinfo(dtype).max
    return np.clip(np.rint(img * top), 0, top).astype(dtype)


def _run_df_main(df, argv, tmp):
    import contextlib
    import io
    out, err = io.StringIO(), io.StringIO()
    old_argv, code = sys.argv, 0
    sys.argv = ["gs360_DualFisheyeDistortionCalibration.py"] + argv
    try:
        with contextlib.redirect_stdout(out), contextlib.redirect_stderr(err):
            try:
                df.main()
            except SystemExit as exc:
                code = exc.code if isinstance(exc.code, int) else 1
    finally:
        sys.argv = old_argv
    norm = lambda t: t.replace(str(tmp), "<TMP>")
    return {"argv": [norm(a) for a in argv], "stdout": norm(out.getvalue()), "stderr": norm(err.getvalue()), "exit": code}


def dump_df_cli(df):
    import argparse
    import tempfile
    import cv2
    # (1) the parser
    captured = {}
    real_parse = argparse.ArgumentParser.parse_args

    def spy(self, *a, **k):
        captured["parser"] = self
        return real_parse(self, ["--input-dir", "x"])
    argparse.ArgumentParser.parse_args = spy
    try:
        df.parse_arguments()
    finally:
        argparse.ArgumentParser.parse_args = real_parse
    actions = []
    for act in captured["parser"]._actions:
        if not act.option_strings or act.dest == "help":
            continue
        default = act.default
        if act.dest == "camera_xml":
            default = "<TEMPLATES>/" + pathlib.Path(default).name
        elif act.dest == "dlogm_lut":
            default = "<TEMPLATES>/" + pathlib.Path(default).name
        elif act.dest == "workers":
            default = "<CPU_COUNT>"
        actions.append({"flags": list(act.option_strings), "dest": act.dest, "default": default,
                        "type": getattr(act.type, "__name__", None), "choices": list(act.choices) if act.choices else None,
                        "action": type(act).__name__, "required": bool(act.required)})
    meta = {"actions": actions, "runs": {}}
    arrays = {}
    rng = np.random.default_rng(99)
    with tempfile.TemporaryDirectory() as tmp:
        tmp = pathlib.Path(tmp).resolve()
        frames = tmp / "frames"
        frames.mkdir()
        W = H = 160
        (tmp / "cal.xml").write_text(TINY_XML.format(w=W, h=H, f=44.0))
        arrays["cal_xml"] = np.array((tmp / "cal.xml").read_text())
        for name in ("a_X", "a_Y", "b_X", "b_Y", "c_X"):
            img = _smooth_image(rng, H, W)
            cv2.imwrite(str(frames / (name + ".png")), img)
            arrays["in_" + name] = img
        (frames / "notes.txt").write_text("not an image")
        masks = tmp / "masks"
        masks.mkdir()
        for name in ("a_X", "a_Y", "b_X", "b_Y"):
            m = (rng.random((H, W)) > 0.5).astype(np.uint8) * 255
            cv2.imwrite(str(masks / (name + ".png")), m)
            arrays["mask_" + name] = m
        (tmp / "emptymasks").mkdir()
        _write_cube(tmp / "look.cube", 5, rng, None)
        arrays["cube_text"] = np.array((tmp / "look.cube").read_text())
        base = ["--input-dir", str(frames), "--camera-xml", str(tmp / "cal.xml"), "--perspective-size", "48", "--workers", "2"]
        runs = {
            "dry_default": base + ["--dry-run"],
            "dry_all_outputs": base + ["--dry-run", "--save-fisheye-output", "--save-color-corrected-output",
                                       "--mask-input-dir", str(masks), "--input-lut", str(tmp / "look.cube"),
                                       "--perspective-ext", "PNG", "--undistort-zoom", "1.2", "--limit", "3",
                                       "--report-json", "r.json"],
            "err_no_input": ["--camera-xml", str(tmp / "cal.xml")],
            "err_missing_dir": ["--input-dir", str(tmp / "nope"), "--camera-xml", str(tmp / "cal.xml")],
            "err_all_disabled": base + ["--no-perspective"],
            "err_suffixes": base + ["--suffixes", "_X"],
            "err_zoom": base + ["--undistort-zoom", "-1"],
            "err_workers": base + ["--workers", "0"],
            "err_masks_missing": base + ["--mask-input-dir", str(tmp / "emptymasks")],
            "err_xml_missing": ["--input-dir", str(frames), "--camera-xml", str(tmp / "absent.xml")],
            "err_color_profile_lut_missing": base + ["--input-color-profile", "osmo360-dlogm", "--dlogm-lut", str(tmp / "none.cube")],
            "real": base + ["--save-fisheye-output", "--save-color-corrected-output", "--mask-input-dir", str(masks),
                            "--input-lut", str(tmp / "look.cube"), "--perspective-ext", "png", "--interpolation", "linear",
                            "--mask-value", "7"],
        }
        for name, argv in runs.items():
            meta["runs"][name] = _run_df_main(df, argv, tmp)
        for sub in ("frames_perspective_colmap/Images", "frames_perspective_colmap/Masks", "frames_undistorted",
                    "frames_colorcorrected"):
            d = tmp / sub
            files = sorted(p.name for p in d.iterdir()) if d.is_dir() else []
            meta.setdefault("real_files", {})[sub] = files
            for fn in files:
                arrays["out_%s_%s" % (sub.replace("/", "_"), fn)] = cv2.imread(str(d / fn), cv2.IMREAD_UNCHANGED)
    (HERE / "df_cli.json").write_text(json.dumps(meta, indent=1, sort_keys=True) + "\n")
    np.savez_compressed(HERE / "df_cli_outputs.npz", **arrays)


def dump_cv2():
    import cv2
    rng = np.random.default_rng(20261017)
    arrays = {}
    h, w, n = 40, 56, 48
    flags = {"nearest": cv2.INTER_NEAREST, "linear": cv2.INTER_LINEAR, "cubic": cv2.INTER_CUBIC,
             "lanczos4": cv2.INTER_LANCZOS4}
    mx = (rng.random((n, n)) * (w + 8) - 4).astype(np.float32)
    my = (rng.random((n, n)) * (h + 8) - 4).astype(np.float32)
    # exact bin centres / boundaries and integer positions too
    mx[0, :] = np.arange(n, dtype=np.float32) + np.float32(1.0 / 64)
    my[0, :] = 7.0
    mx[1, :] = np.arange(n, dtype=np.float32)
    my[1, :] = np.arange(n, dtype=np.float32) * np.float32(0.5) + np.float32(0.5)
    arrays["map_x"], arrays["map_y"] = mx, my
    for dt in ("uint8", "uint16", "float32"):
        for ch in (1, 3):
            if dt == "float32":
                src = rng.random((h, w, ch), dtype=np.float32)
            else:
                src = rng.integers(0, np.iinfo(dt).max + 1, (h, w, ch)).astype(dt)
            if ch == 1:
                src = src[..., 0]
            arrays["src_%s_c%d" % (dt, ch)] = src
            for interp, flag in flags.items():
                for bv in (0, 37):
                    arrays["out_%s_c%d_%s_b%d" % (dt, ch, interp, bv)] = cv2.remap(
                        src, mx, my, flag, borderMode=cv2.BORDER_CONSTANT, borderValue=(float(bv),) * 4)
    arrays["cv2_version"] = np.array(cv2.__version__)
    np.savez_compressed(HERE / "cv2_remap.npz", **arrays)


def _extrinsics_xml_text(n_pairs=3, with_chunk_transform=True, two_sensors=False):
    """A synthetic aligned dual-fisheye project: the template's sensor block(s), X/Y cameras with 4x4 transforms,
    a chunk similarity (rotation / translation / scale children), a component transform given as 16 numbers, one
    disabled camera and one camera without a transform."""
    rng = np.random.default_rng(2024)

    def rot(ax, ay, az):
        cx, sx, cy, sy, cz, sz = math.cos(ax), math.sin(ax), math.cos(ay), math.sin(ay), math.cos(az), math.sin(az)
        rx = np.array([[1, 0, 0], [0, cx, -sx], [0, sx, cx]])
        ry = np.array([[cy, 0, sy], [0, 1, 0], [-sy, 0, cy]])
        rz = np.array([[cz, -sz, 0], [sz, cz, 0], [0, 0, 1]])
        return rz @ ry @ rx

    calib = ("<calibration type=\"equisolid_fisheye\" class=\"adjusted\"><resolution width=\"3840\" height=\"3840\"/>"
             "<f>1049.9268186384606</f><cx>-0.053481903280599763</cx><cy>-0.040449115818567277</cy>"
             "<k1>0.10190869149858893</k1><k2>0.00079808296648272998</k2><k3>-0.00031893309097734927</k3></calibration>")
    sensors = ["<sensor id=\"%d\" label=\"unknown\" type=\"equisolid_fisheye\"><resolution width=\"3840\" height=\"3840\"/>%s</sensor>"
               % (k, calib) for k in range(2 if two_sensors else 1)]
    cams = []
    cid = 0
    for n in range(n_pairs):
        base = rot(*(rng.random(3) * 2.0 - 1.0))
        centre = rng.random(3) * 10.0 - 5.0
        for lens, flip in (("X", np.eye(3)), ("Y", rot(0.0, math.pi, 0.0))):
            m = np.eye(4)
            m[:3, :3] = base @ flip
            m[:3, 3] = centre + (0.01 if lens == "Y" else 0.0)
            sid = 1 if (two_sensors and lens == "Y") else 0
            cams.append("<camera id=\"%d\" sensor_id=\"%d\" component_id=\"0\" label=\"frame%04d_%s\"><transform>%s</transform></camera>"
                        % (cid, sid, n + 1, lens, " ".join(repr(float(v)) for v in m.reshape(-1))))
            cid += 1
    cams.append("<camera id=\"%d\" sensor_id=\"0\" component_id=\"0\" label=\"frame9000_X\" enabled=\"false\"><transform>%s</transform></camera>"
                % (cid, " ".join(repr(float(v)) for v in np.eye(4).reshape(-1))))
    cams.append("<camera id=\"%d\" sensor_id=\"0\" component_id=\"0\" label=\"frame9001_X\"/>" % (cid + 1))
    comp_m = np.eye(4)
    comp_m[:3, :3] = 1.7 * rot(0.3, -0.2, 0.9)
    comp_m[:3, 3] = [1.0, -2.0, 0.5]
    comp = "<components next_id=\"1\" active_id=\"0\"><component id=\"0\" label=\"Component 1\"><transform>%s</transform></component></components>" \
           % " ".join(repr(float(v)) for v in comp_m.reshape(-1))
    chunk_tf = ""
    if with_chunk_transform:
        r = rot(-0.4, 0.25, 0.1)
        chunk_tf = ("<transform><rotation locked=\"false\">%s</rotation><translation locked=\"false\">3.5 -1.25 0.75</translation>"
                    "<scale locked=\"true\">2.5</scale></transform>") % " ".join(repr(float(v)) for v in r.reshape(-1))
    return ("<?xml version=\"1.0\" encoding=\"UTF-8\"?>\n<document version=\"2.3.0\"><chunk label=\"Chunk 1\" enabled=\"true\">"
            "<sensors next_id=\"2\">%s</sensors>%s<cameras next_id=\"%d\" next_group_id=\"0\">%s</cameras>%s</chunk></document>\n"
            % ("".join(sensors), comp, cid + 2, "".join(cams), chunk_tf))


def _ply_bytes(binary, with_color=True, n=7):
    import struct
    rng = np.random.default_rng(77)
    pts = rng.random((n, 3)) * 20.0 - 10.0
    cols = rng.integers(0, 256, (n, 3))
    head = ["ply", "format %s 1.0" % ("binary_little_endian" if binary else "ascii"), "comment synthetic",
            "element vertex %d" % n, "property float x", "property float y", "property double z"]
    if with_color:
        head += ["property uchar red", "property uchar green", "property uchar blue"]
    head += ["element face 0", "property list uchar int vertex_indices", "end_header"]
    blob = ("\n".join(head) + "\n").encode("ascii")
    for p3, c3 in zip(pts, cols):
        if binary:
            blob += struct.pack("<ffd", *p3) + (struct.pack("<BBB", *[int(v) for v in c3]) if with_color else b"")
        else:
            blob += (" ".join([repr(float(np.float32(p3[0]))), repr(float(np.float32(p3[1]))), repr(float(p3[2]))] +
                              ([str(int(v)) for v in c3] if with_color else [])) + "\n").encode("ascii")
    return blob


def dump_df_metadata(df):
    """Pose / COLMAP / Metashape-XML export of the dual-fisheye tool (DF:917-963, :1348-1686, :2812-2833) on
    synthetic aligned projects: library-level results and whole ``main()`` runs with --metadata-only."""
    import base64
    import tempfile
    out = {"inputs": {}, "cases": [], "cli": []}
    with tempfile.TemporaryDirectory() as tmp_s:
        tmp = pathlib.Path(tmp_s)
        variants = {"chunk": dict(with_chunk_transform=True), "component": dict(with_chunk_transform=False),
                    "two_sensors": dict(with_chunk_transform=True, two_sensors=True)}
        plys = {"binary_color": _ply_bytes(True, True), "ascii_color": _ply_bytes(False, True), "binary_plain": _ply_bytes(True, False)}
        for name, blob in plys.items():
            (tmp / (name + ".ply")).write_bytes(blob)
            out["inputs"][name + ".ply"] = base64.b64encode(blob).decode("ascii")
        specs = df.build_sfm10_specs(output_size=320, focal_mm=14.0, sensor_mm="36 36", yaw_delta_deg=40.0, pitch_delta_deg=40.0)
        lens_of = {"A": "X", "A_U": "X", "A_D": "X", "B": "X", "J": "X", "E": "Y", "F": "Y", "F_U": "Y", "F_D": "Y", "G": "Y"}
        for vname, kw in variants.items():
            text = _extrinsics_xml_text(**kw)
            xml_path = tmp / (vname + ".xml")
            xml_path.write_text(text)
            out["inputs"][vname + ".xml"] = text
            cams = df.msxml_converter.load_metashape_cameras(xml_path)
            sensor_map, camera_to_sensor = df.load_metashape_calibration(xml_path)
            labels = set(df.build_camera_transform_map(xml_path).keys())
            pairs = df.build_metadata_only_resolved_pairs(camera_to_sensor, sensor_map, "_X", "_Y", labels)
            cache = {(p[4], p[5]): {vid: {"lens_key": k} for vid, k in lens_of.items()} for p in pairs}
            frames = df.build_perspective_pose_frames(xml_path, pairs, {p[1] for p in pairs[:2]}, specs, cache, ".jpg", 0.0, 180.0)
            cameras, images = df.build_colmap_model_from_pose_frames(frames, 320, 14.0, "36 24")
            case = {"xml": vname + ".xml", "cameras_loaded": [[cid, label, mat] for cid, label, mat in cams],
                    "pairs": [[p[0], p[1], str(p[2]), str(p[3]), p[4], p[5]] for p in pairs],
                    "frames": [{k: v for k, v in f.items()} for f in frames], "colmap_cameras": cameras,
                    "colmap_images": images, "files": {}}
            for pname in plys:
                points = df.build_colmap_points_from_metashape_ply(tmp / (pname + ".ply"))
                odir = tmp / ("out_" + vname + "_" + pname)
                df.camera_converter.write_colmap_text_model(odir, cameras, images, points)
                df.camera_converter.export_metashape_perspective_xml(odir / "persp.xml", cameras, images)
                case["files"][pname] = {f.name: f.read_text() for f in sorted(odir.iterdir())}
            out["cases"].append(case)
        # whole-program runs (metadata only: no images are read)
        root = tmp / "persp_out"
        base = ["--metadata-only", "--camera-extrinsics-xml", str(tmp / "chunk.xml"), "--pointcloud-ply",
                str(tmp / "binary_color.ply"), "--perspective-output-dir", str(root), "--perspective-size", "256"]
        for extra in (["--dry-run"], [], ["--camera-extrinsics-xml", str(tmp / "two_sensors.xml"), "--perspective-ext", "png"]):
            run = _run_df_main(df, base + extra, tmp)
            run["files"] = {}
            if root.exists():
                for f in sorted(root.rglob("*")):
                    if f.is_file():
                        run["files"][str(f.relative_to(root))] = f.read_text()
                        f.unlink()
            out["cli"].append(run)
        for bad in (["--metadata-only", "--camera-extrinsics-xml", str(tmp / "chunk.xml")],
                    ["--metadata-only", "--pointcloud-ply", str(tmp / "binary_color.ply")],
                    ["--metadata-only", "--camera-extrinsics-xml", str(tmp / "chunk.xml"), "--pointcloud-ply", str(tmp / "nope.ply")]):
            out["cli"].append(_run_df_main(df, bad, tmp))
    (HERE / "df_metadata.json").write_text(json.dumps(out, indent=1) + "\n")


def _ms_xml_text():
    """A synthetic spherical Metashape project for the MS360xmlToPersCams tool: a sensor with data_type / black_level /
    sensitivity, four cameras (a plain label, one that already carries a view suffix, one with a path separator, a
    disabled one), a chunk similarity given as rotation / translation / scale children."""
    rng = np.random.default_rng(4711)

    def rot(ax, ay, az):
        cx, sx, cy, sy, cz, sz = math.cos(ax), math.sin(ax), math.cos(ay), math.sin(ay), math.cos(az), math.sin(az)
        rx = np.array([[1, 0, 0], [0, cx, -sx], [0, sx, cx]])
        ry = np.array([[cy, 0, sy], [0, 1, 0], [-sy, 0, cy]])
        rz = np.array([[cz, -sz, 0], [sz, cz, 0], [0, 0, 1]])
        return rz @ ry @ rx

    cams = []
    for cid, (label, extra) in enumerate((("IMG_0001", ""), ("IMG_0002_A", ""), ("sub/IMG_0003", ""), ("IMG_0004", " enabled=\"false\""))):
        m = np.eye(4)
        m[:3, :3] = rot(*(rng.random(3) * 2.0 - 1.0))
        m[:3, 3] = rng.random(3) * 6.0 - 3.0
        cams.append("<camera id=\"%d\" sensor_id=\"0\" component_id=\"0\" label=\"%s\"%s><transform>%s</transform></camera>"
                    % (cid, label, extra, " ".join(repr(float(v)) for v in m.reshape(-1))))
    r = rot(0.2, -0.35, 0.6)
    chunk_tf = ("<transform><rotation locked=\"false\">%s</rotation><translation locked=\"false\">-1.5 0.25 2.0</translation>"
                "<scale locked=\"true\">1.75</scale></transform>") % " ".join(repr(float(v)) for v in r.reshape(-1))
    sensor = ("<sensor id=\"0\" label=\"spherical\" type=\"spherical\"><resolution width=\"7680\" height=\"3840\"/>"
              "<data_type>uint16</data_type><black_level>1 2 3</black_level><sensitivity>1 0.5 0.25</sensitivity></sensor>")
    return ("<?xml version=\"1.0\" encoding=\"UTF-8\"?>\n<document version=\"2.3.0\"><chunk label=\"Chunk 1\" enabled=\"true\">"
            "<sensors next_id=\"1\">%s</sensors><cameras next_id=\"4\" next_group_id=\"0\">%s</cameras>%s</chunk></document>\n"
            % (sensor, "".join(cams), chunk_tf))


def _run_ms_main(ms, argv, tmp):
    import contextlib
    import io
    out, err = io.StringIO(), io.StringIO()
    old_argv, code = sys.argv, 0
    sys.argv = ["gs360_MS360xmlToPersCams.py"] + argv
    try:
        with contextlib.redirect_stdout(out), contextlib.redirect_stderr(err):
            try:
                ms.main()
            except SystemExit as exc:
                code = exc.code if isinstance(exc.code, int) else 1
    finally:
        sys.argv = old_argv
    norm = lambda t: t.replace(str(tmp), "<TMP>")
    return {"argv": [norm(a) for a in argv], "stdout": norm(out.getvalue()), "stderr": norm(err.getvalue()), "exit": code}


def dump_ms_export(ms):
    """The MS360xmlToPersCams pose exporters (MS:592-720 view sets and intrinsics, :987-1250 writers, :2054-2177
    main) on a synthetic spherical project: view sets / intrinsics per preset and whole ``main()`` runs -- stdout,
    stderr, exit code and every written file (text as text, PLY as base64)."""
    import base64
    import tempfile
    out = {"inputs": {}, "presets": {}, "runs": []}
    for preset in ms.PRESET_CHOICES:
        cfg = ms.preset_config(preset)
        focal = cfg["focal_mm"] if cfg.get("hfov_deg") is None else ms.focal_from_hfov_deg(float(cfg["hfov_deg"]), ms.SENSOR_W_MM)
        out["presets"][preset] = {"views": [list(v) for v in ms.build_views(preset)], "size": int(cfg["size"]), "focal_mm": focal,
                                  "intrinsics": list(ms.compute_intrinsics(focal, int(cfg["size"]), int(cfg["size"])))}
    with tempfile.TemporaryDirectory() as tmp_s:
        tmp = pathlib.Path(tmp_s)
        text = _ms_xml_text()
        (tmp / "cameras.xml").write_text(text)
        out["inputs"]["cameras.xml"] = text
        for name, blob in (("binary_color.ply", _ply_bytes(True, True)), ("binary_plain.ply", _ply_bytes(True, False))):
            (tmp / name).write_bytes(blob)
            out["inputs"][name] = base64.b64encode(blob).decode("ascii")
        xml = str(tmp / "cameras.xml")
        runs = [
            [xml, "--format", "all", "--points-ply", str(tmp / "binary_color.ply")],
            [xml, "--format", "all", "--preset", "cube105", "--scale", "100", "--world-rot-axis", "1 0 0", "--world-rot-deg", "90",
             "--points-ply", str(tmp / "binary_plain.ply"), "--pc-rotate-x-plus180", "--ext", ".png", "-o", str(tmp / "out_cube")],
            [xml, "--preset", "fisheyelike", "-o", str(tmp / "out_fl")],
            [xml, "--format", "transforms", "--preset", "evenMinus30", "-o", str(tmp / "out_even")],
            [xml, "--format", "realityscan", "--preset", "2views", "--world-rot-axis", "0,0,1", "--world-rot-deg", "-30", "-o", str(tmp / "out_rs")],
            [xml, "--format", "colmap", "--preset", "default", "--points-ply", str(tmp / "binary_color.ply"), "--pc-rotate-x-minus90",
             "-o", str(tmp / "out_colmap")],
            [xml, "--format", "colmap", "-o", str(tmp / "out_err1")],
            [str(tmp / "missing.xml")],
            [xml, "--format", "metashape-multi-camera-system", "--preset", "default"],
            [xml, "--format", "all", "--points-ply", str(tmp / "nope.ply"), "-o", str(tmp / "out_err2")],
        ]
        for argv in runs:
            before = {f for f in tmp.rglob("*") if f.is_file()}
            run = _run_ms_main(ms, argv, tmp)
            run["files"] = {}
            for f in sorted(set(f for f in tmp.rglob("*") if f.is_file()) - before):
                rel = str(f.relative_to(tmp))
                run["files"][rel] = ("base64:" + base64.b64encode(f.read_bytes()).decode("ascii")) if f.suffix == ".ply" else f.read_text()
                f.unlink()
            out["runs"].append(run)
    (HERE / "ms_export.json").write_text(json.dumps(out, indent=1) + "\n")


def dump_gui_geometry(reference_dir):
    """The in-repo float64 statement of the view geometry (gs360_GUI.py:342-395, :419-424), evaluated by the
    reference's OWN functions.  gs360_GUI.py cannot be imported here (it needs tkinter), so the five pure functions
    are taken out of its syntax tree and compiled as they stand."""
    import ast
    import typing
    src = (pathlib.Path(reference_dir) / "gs360_GUI.py").read_text()
    wanted = ("normalize_vector", "rotate_pitch", "rotate_yaw", "direction_from_uv", "lonlat_to_xy")
    tree = ast.parse(src)
    picked = [node for node in tree.body if isinstance(node, ast.FunctionDef) and node.name in wanted]
    assert sorted(n.name for n in picked) == sorted(wanted)
    env = {"math": math, "Tuple": typing.Tuple, "Sequence": typing.Sequence, "Iterable": typing.Iterable, "List": typing.List}
    exec(compile(ast.Module(body=picked, type_ignores=[]), "gs360_GUI.py", "exec"), env)
    views = [(0.0, 0.0, 104.2500326978036, 104.2500326978036), (45.0, 30.0, 104.2500326978036, 104.2500326978036),
             (-135.0, -30.0, 104.2500326978036, 104.2500326978036), (180.0, 0.0, 93.2731514, 93.2731514),
             (179.9, 0.0, 112.6198649, 112.6198649), (0.0, 90.0, 112.6198649, 112.6198649), (0.0, -90.0, 100.0, 80.0),
             (123.4, 56.7, 60.0, 45.0), (-70.0, -60.0, 120.0, 90.0), (320.0, 0.0, 104.25, 104.25)]
    n = 33                                              # 33 x 33 pixel centres of an n x n view: u = (2i + 1) / n - 1
    uv = [(2 * i + 1) / n - 1.0 for i in range(n)]
    out = {"views": np.array(views), "uv": np.array(uv)}
    for (W, H) in ((7680, 3840), (3840, 1920)):
        xs = np.empty((len(views), n, n)); ys = np.empty_like(xs); lons = np.empty_like(xs); lats = np.empty_like(xs)
        for k, (yaw, pitch, hf, vf) in enumerate(views):
            # the clamps of sample_view_segments (GUI:437-440)
            hfr, vfr = math.radians(min(max(hf, 1e-3), 179.9)), math.radians(min(max(vf, 1e-3), 179.9))
            for j, v in enumerate(uv):
                for i, u in enumerate(uv):
                    lon, lat = env["direction_from_uv"](u, v, hfr, vfr, math.radians(yaw), math.radians(pitch))
                    x, y = env["lonlat_to_xy"](lon, lat, W, H)
                    xs[k, j, i], ys[k, j, i], lons[k, j, i], lats[k, j, i] = x, y, lon, lat
        out["x_%d" % W], out["y_%d" % W] = xs, ys
        if W == 7680:
            out["lon"], out["lat"] = lons, lats
    np.savez_compressed(HERE / "gui_geometry.npz", **out)


def dump_v2f(reference_dir):
    """The v360 filter string gs360_Video2Frames.py builds for --fisheye-perspective (V2F:467-487).  The code
    sits inside main(); the block is cut out of the reference's source text and executed as it stands."""
    import textwrap
    import types
    src = (pathlib.Path(reference_dir) / "cli_tools" / "gs360_Video2Frames.py").read_text()
    start = src.index("        focal_mm = max(args.fisheye_focal_mm, 1e-6)")
    end = src.index("        vf_chain = [v360_filter", start)
    block = textwrap.dedent(src[start:end])
    import gs360_360PerspCut as pc
    cases = []
    for projection in ("equidistant", "equisolid", "bogus"):
        for input_fov, focal, size in ((190.0, 8.0, 1600), (220.0, 12.0, 1024), (400.0, 0.5, 3), (0.2, 300.0, 2000),
                                       (180.0, 17.5, 1750)):
            env = {"args": types.SimpleNamespace(fisheye_focal_mm=focal, fisheye_size=size,
                                                 fisheye_projection=projection, fisheye_input_fov=input_fov),
                   "fov_from_focal_mm": pc.fov_from_focal_mm, "v_fov_from_hfov": pc.v_fov_from_hfov,
                   "FISHEYE_SENSOR_WIDTH_MM": 36.0}
            exec(block, env)
            cases.append({"projection": projection, "input_fov": input_fov, "focal_mm": focal, "size": size,
                          "filter": env["v360_filter"], "hfov_deg": env["hfov_deg"], "vfov_deg": env["vfov_deg"]})
    (HERE / "v2f_fisheye.json").write_text(json.dumps({"cases": cases}, indent=1) + "\n")


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("reference", nargs="?", default="/root/reference")
    ap.add_argument("--only", default="", help="regenerate one family only (v2f)")
    ns = ap.parse_args()
    if ns.only == "gui_geometry":
        dump_gui_geometry(ns.reference)
        return
    if ns.only in ("v2f", "df_metadata", "ms_export"):
        sys.dont_write_bytecode = True
        sys.path.insert(0, str(pathlib.Path(ns.reference) / "cli_tools"))
        if ns.only == "v2f":
            dump_v2f(ns.reference)
        elif ns.only == "ms_export":
            import gs360_MS360xmlToPersCams as ms
            dump_ms_export(ms)
        else:
            import gs360_DualFisheyeDistortionCalibration as df
            dump_df_metadata(df)
        return
    sys.dont_write_bytecode = True
    sys.path.insert(0, str(pathlib.Path(ns.reference) / "cli_tools"))
    import gs360_360PerspCut as pc
    import gs360_DualFisheyeDistortionCalibration as df
    dump_perspcut(pc)
    dump_dualfisheye(df)
    dump_undistort(df)
    dump_color(df)
    dump_df_cli(df)
    dump_cv2()
    dump_v2f(ns.reference)
    dump_gui_geometry(ns.reference)
    dump_df_metadata(df)
    import gs360_MS360xmlToPersCams as ms
    dump_ms_export(ms)
    for p in sorted(HERE.glob("*.json")) + sorted(HERE.glob("*.npz")):
        print("%9d  %s" % (p.stat().st_size, p.name))


if __name__ == "__main__":
    main()
