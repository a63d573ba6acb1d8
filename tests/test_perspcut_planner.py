"""The drop-in planner (remap360.perspcut) against outputs recorded from the reference's
gs360_360PerspCut.build_view_jobs / helpers (tests/golden/make_golden.py)."""

import pathlib

import pytest

from remap360 import perspcut as pc


def _args(argv, video, depth):
    args = pc.create_arg_parser().parse_args(argv)
    for name in ("size", "hfov", "focal_mm"):
        setattr(args, name + "_explicit", getattr(args, name + "_explicit", False))
    args.input_is_video = video
    args.video_bit_depth = depth
    return args


def test_every_recorded_case(golden_views):
    assert len(golden_views) >= 25
    for name, case in golden_views.items():
        args = _args(case["argv"], case["video"], case["bit_depth"])
        files = [pathlib.Path("/tmp/in") / f for f in case["files"]]
        res = pc.build_view_jobs(args, files, pathlib.Path("/tmp/out"))
        assert [[list(c), s, d] for c, s, d in res.jobs] == case["jobs"], name
        got_specs = [{"source_path": str(v.source_path), "output_name": v.output_name, "view_id": v.view_id,
                      "yaw_deg": v.yaw_deg, "pitch_deg": v.pitch_deg, "hfov_deg": v.hfov_deg,
                      "vfov_deg": v.vfov_deg, "width": v.width, "height": v.height, "projection": v.projection}
                     for v in res.view_specs]
        assert got_specs == case["view_specs"], name
        for key in ("focal_used_mm", "focal_35mm_equiv", "hfov_deg", "vfov_deg", "preview_views_line",
                    "sensor_line", "realityscan_line", "metashape_line"):
            assert getattr(res, key) == case[key], (name, key)
        assert res.total == len(case["jobs"])
        assert {k: getattr(args, k) for k in case["args_after"]} == case["args_after"], name
        # the GUI edits job argv lists in place (gs360_GUI.py:19092-19147)
        assert all(isinstance(c, list) and all(isinstance(t, str) for t in c) for c, _, _ in res.jobs)


def test_readme_focal_numbers(golden_views):
    # README.md:68-74 of the reference: 12 / 17 / 14 mm -> 533.33333 / 755.55556 / 622.22222 px
    for preset, mm, px in (("default", 12.0, "533.33333"), ("fisheyelike", 17.0, "755.55556"),
                           ("full360coverage", 14.0, "622.22222")):
        case = golden_views[preset]
        assert case["focal_used_mm"] == mm and ("f=  " + px) in case["metashape_line"]
        args = _args(case["argv"], False, 8)
        res = pc.build_view_jobs(args, [pathlib.Path("/tmp/in/pano0001.jpg")], pathlib.Path("/tmp/out"))
        assert ("f=  " + px) in res.metashape_line and "focal length=  %.3f mm" % mm in res.realityscan_line


def test_helpers(golden_helpers):
    h = golden_helpers
    for f, s, want in h["fov_from_focal_mm"]:
        assert pc.fov_from_focal_mm(f, s) == want
    for a, s, want in h["focal_from_hfov_deg"]:
        assert pc.focal_from_hfov_deg(a, s) == want
    for a, w, hh, want in h["v_fov_from_hfov"]:
        assert pc.v_fov_from_hfov(a, w, hh) == want
    for i, want in h["letter_tag"]:
        assert pc.letter_tag(i) == want
    for s, want in h["letter_to_index1"]:
        assert pc.letter_to_index1(s) == want
    for a, want in h["normalize_angle_deg"]:
        assert pc.normalize_angle_deg(a) == want
    for d, dd, want in h["extra_suffix"]:
        assert pc.extra_suffix(d, dd) == want
    for s, want in h["parse_jobs"]:
        assert pc.parse_jobs(s) == want
    assert pc.parse_jobs("auto") >= 1
    for s, want in h["parse_sensor"]:
        assert pc.parse_sensor(s) == want
    for s, d, want in h["parse_addcam_spec"]:
        assert {str(k): v for k, v in pc.parse_addcam_spec(s, d).items()} == want
    for s, want in h["parse_delcam_spec"]:
        assert sorted(pc.parse_delcam_spec(s)) == want
    for s, d, a, b, c, e in h["parse_setcam_spec"]:
        ga, gb, gc, ge = pc.parse_setcam_spec(s, d)
        assert {str(k): v for k, v in ga.items()} == a and {str(k): v for k, v in gb.items()} == b
        assert {"%d%s" % k: v for k, v in gc.items()} == c and {"%d%s" % k: v for k, v in ge.items()} == e


def test_error_behaviour_matches_the_reference():
    with pytest.raises(ValueError, match="invalid --addcam token"):
        pc.parse_addcam_spec("B:+10", 30.0)
    with pytest.raises(ValueError, match="invalid --setcam token"):
        pc.parse_setcam_spec("A", 30.0)
    with pytest.raises(ValueError, match="invalid --setcam token"):
        pc.parse_setcam_spec("A=up", 30.0)
    with pytest.raises(ValueError):
        pc.letter_to_index1("?")
    args = _args(["-i", "/tmp/in", "--count", "0"], False, 8)
    with pytest.raises(SystemExit) as ei:
        pc.build_view_jobs(args, [pathlib.Path("/tmp/in/a.jpg")], pathlib.Path("/tmp/out"))
    assert ei.value.code == 1
    with pytest.raises(ValueError, match="fps must be specified"):
        pc.build_ffmpeg_cmd("ffmpeg", pathlib.Path("v.mp4"), pathlib.Path("o.jpg"), 8, 8, 0.0, 0.0, 90.0, 90.0,
                            "cubic", ".jpg", video_mode=True)


def test_parser_surface():
    ap = pc.create_arg_parser()
    dests = {a.dest for a in ap._actions}
    for d in ("input_dir", "out_dir", "preset", "count", "addcam", "addcam_deg", "add_top", "add_bottom",
              "add_topdown", "delcam", "setcam", "size", "ext", "jpeg_quality_95", "fps", "start", "end",
              "keep_rec709", "hfov", "focal_mm", "sensor_mm", "jobs", "print_cmd", "ffmpeg", "dry_run"):
        assert d in dests, d
    ns = ap.parse_args(["-i", "x"])
    assert (ns.preset, ns.count, ns.size, ns.focal_mm, ns.sensor_mm, ns.ext, ns.jobs, ns.addcam_deg) == \
        ("default", 8, 1600, 12.0, "36 36", "jpg", "auto", 30.0)
    assert not hasattr(ns, "size_explicit")
    ns = ap.parse_args(["-i", "x", "--size", "900", "--focal-mm", "20"])
    assert ns.size_explicit and ns.focal_mm_explicit and not hasattr(ns, "hfov_explicit")
    assert pc.PROGRESS_INTERVAL == 5 and not pc.stop_event.is_set()
    assert pc.EXTS == {".tif", ".tiff", ".jpg", ".jpeg", ".png"}


def test_run_one_honours_stop_event():
    pc.stop_event.set()
    try:
        assert pc.run_one(["ffmpeg", "-i", "a", "b"]) == (130, "")
    finally:
        pc.stop_event.clear()
