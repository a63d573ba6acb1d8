"""Host-side job handling: the ffmpeg-style argv kept for GUI compatibility is parsed back into a
view job; CLI process contract in dry-run; dual-fisheye host helpers against recorded values."""

import json
import pathlib
import subprocess
import sys

import pytest

from conftest import PKG_DIR
from remap360 import dualfisheye as dfh
from remap360 import executor, perspcut as pc, video


def _jobs(argv, video_mode=False, depth=8):
    args = pc.create_arg_parser().parse_args(argv)
    for name in ("size", "hfov", "focal_mm"):
        setattr(args, name + "_explicit", getattr(args, name + "_explicit", False))
    args.input_is_video, args.video_bit_depth = video_mode, depth
    files = [pathlib.Path("/tmp/in/vid.mp4")] if video_mode else [pathlib.Path("/tmp/in/a.jpg")]
    return pc.build_view_jobs(args, files, pathlib.Path("/tmp/out"))


def test_job_argv_round_trip():
    res = _jobs(["-i", "/tmp/in", "--preset", "fisheyelike", "--jpeg-quality-95"])
    assert len(res.jobs) == len(res.view_specs) == 10
    for (cmd, _src, dst), spec in zip(res.jobs, res.view_specs):
        job = executor.parse_job_argv(cmd)
        assert job.source == pathlib.Path("/tmp/in/a.jpg") and job.output.name == dst == spec.output_name
        assert (job.width, job.height) == (spec.width, spec.height) == (1600, 1600)
        assert (job.yaw, job.pitch, job.hfov, job.vfov) == (spec.yaw_deg, spec.pitch_deg, spec.hfov_deg, spec.vfov_deg)
        assert job.projection == "rectilinear" and job.interp == "cubic" and not job.video and job.jpeg_quality == 95


def test_video_and_fisheye_jobs_are_recognised():
    res = _jobs(["-i", "/tmp/in/vid.mp4", "-f", "2", "--ext", "png", "--start", "3", "--end", "9.5"], True, 10)
    job = executor.parse_job_argv(res.jobs[0][0])
    assert job.video and job.fps == 2.0 and job.start == 3.0 and job.end == 9.5 and job.pix_fmt == "rgb48le"
    assert "%07d" in str(job.output)
    res = _jobs(["-i", "/tmp/in", "--preset", "fisheyeXY"])
    job = executor.parse_job_argv(res.jobs[0][0])
    assert job.projection == "fisheye" and job.hfov == 180.0 and job.width == 3600
    # the GUI rewrites -vf in place (select filter, gs360_GUI.py:19092-19147): still parseable
    cmd = list(_jobs(["-i", "/tmp/in"]).jobs[0][0])
    cmd[cmd.index("-vf") + 1] = "select='eq(n\\,3)'," + cmd[cmd.index("-vf") + 1]
    assert executor.parse_job_argv(cmd).yaw == 0.0
    with pytest.raises(executor.JobError):
        executor.parse_job_argv(["ffmpeg", "-i", "x.jpg", "y.jpg"])
    assert executor.run_job_argv(["ffmpeg", "-i", "x.jpg", "y.jpg"])[0] == 1


def test_fps_filter_frame_selection():
    # 30 fps source, 2 fps output: each output slot takes the last input frame that rounds into it
    sel = video._select_frames(61, 30.0, 2.0, None, None)
    assert len(sel) == 5 and sel == sorted(sel) and sel[0] <= 7 and sel[-1] == 60
    # upsampling repeats frames, downsampling with a window starts inside it
    up = video._select_frames(10, 10.0, 20.0, None, None)
    assert len(up) == 19 and up[:4] == [0, 0, 1, 1]
    assert video._select_frames(100, 25.0, 5.0, 2.0, 3.0)[0] >= 50
    assert video._select_frames(0, 25.0, 5.0, None, None) == []


def test_cli_dry_run_process_contract(tmp_path):
    (tmp_path / "in").mkdir()
    (tmp_path / "in" / "pano0001.jpg").write_bytes(b"not decoded in a dry run")
    out = subprocess.run([sys.executable, str(PKG_DIR / "gs360_360PerspCut.py"), "-i", str(tmp_path / "in"),
                          "--preset", "full360coverage", "--dry-run"], capture_output=True, text=True, timeout=120)
    assert out.returncode == 0
    lines = out.stdout.strip().splitlines()
    assert len([ln for ln in lines if ln.startswith("$ ")]) == 12
    assert lines[-1] == "[DRY] Exiting without execution (total 12 commands)"
    bad = subprocess.run([sys.executable, str(PKG_DIR / "gs360_360PerspCut.py"), "-i", str(tmp_path / "nope")],
                         capture_output=True, text=True, timeout=120)
    assert bad.returncode == 1 and "[ERR] Input path not found" in bad.stderr


def test_dualfisheye_host_helpers_match_reference(golden_df):
    for f, s, want in golden_df["compute_view_fov_deg"]:
        assert list(dfh.compute_view_fov_deg(f, s)) == want
    for a, want in golden_df["wrap_angle_deg"]:
        assert dfh.wrap_angle_deg(a) == want
    assert dfh.build_sfm10_specs(1750, 14.0, "36 36", 40.0, 40.0) == golden_df["sfm10_default"]
    assert dfh.build_sfm10_specs(1200, 12.0, "36x24", 35.0, 25.0) == golden_df["sfm10_alt"]
    with pytest.raises(ValueError, match="yaw-delta"):
        dfh.build_sfm10_specs(100, 14.0, "36 36", 180.0, 40.0)
    with pytest.raises(ValueError, match="must be > 0"):
        dfh.compute_view_fov_deg(0.0, "36 36")


def test_auto_undistort_zoom_and_items_match_reference(golden_undistort):
    meta, _ = golden_undistort
    cals = {"tmpl": "tmpl_auto", "syn": "syn_auto", "tiny": "tiny_auto", "wide": "wide_auto"}
    for key, want in meta["auto_zoom"].items():
        parts = key.split("_")
        cal = dfh.SensorCalibration(**meta["cases"][cals[parts[0]]]["calibration"])
        n = int(parts[2][1:]) if len(parts) > 2 else 192
        got = dfh.estimate_auto_undistort_zoom(cal, sample_count=n, lens_fov_deg=float(parts[1]))
        assert abs(got - want) <= 2e-6 * want, (key, got, want)
    wide = dfh.SensorCalibration(**meta["cases"]["wide_auto"]["calibration"])
    items = dfh.build_undistort_items([wide, wide])
    assert [it.src_slot for it in items] == [0, 1]
    assert abs(items[0].zoom - meta["cases"]["wide_auto"]["undistort_zoom"]) < 1e-5
    assert dfh.build_undistort_items([wide], undistort_zoom=0.0)[0].zoom == 1e-6          # DF:1145
    bad = dfh.SensorCalibration(sensor_id="1", model_type="frame", width=10, height=10, f=5.0)
    with pytest.raises(ValueError, match="Unsupported sensor model"):
        dfh.build_undistort_items([bad])


def test_calibration_xml_loader(tmp_path, golden_df):
    want = golden_df["sensors"]["0"]
    xml = tmp_path / "cal.xml"
    xml.write_text("""<?xml version="1.0"?><document><chunk><sensors>
      <sensor id="0" type="equisolid_fisheye"><resolution width="%d" height="%d"/>
        <calibration type="equisolid_fisheye" class="initial"><resolution width="%d" height="%d"/><f>1050</f></calibration>
        <calibration type="equisolid_fisheye" class="adjusted"><resolution width="%d" height="%d"/>
          <f>%r</f><cx>%r</cx><cy>%r</cy><k1>%r</k1><k2>%r</k2><k3>%r</k3></calibration></sensor>
      <sensor id="7"><resolution width="10" height="10"/></sensor>
      </sensors><cameras><camera id="0" sensor_id="0" label="frame_X"/></cameras></chunk></document>""" % (
        want["width"], want["height"], want["width"], want["height"], want["width"], want["height"],
        want["f"], want["cx"], want["cy"], want["k1"], want["k2"], want["k3"]))
    sensors, cams = dfh.load_metashape_calibration(xml)
    assert list(sensors) == ["0"] and cams == {"frame_X": "0"}
    got = sensors["0"]
    for key in ("width", "height", "f", "cx", "cy", "k1", "k2", "k3", "k4", "p1", "p2", "b1", "b2", "model_type"):
        assert getattr(got, key) == want[key], key


def test_rgb48le_outputs_replicate_the_byte(tmp_path):
    """PNG / TIFF views of a >8-bit video are 16-bit files holding the 8-bit result widened by swscale's rule."""
    import numpy as np
    cv2 = pytest.importorskip("cv2")
    img = np.arange(3 * 4 * 3, dtype=np.uint8).reshape(3, 4, 3) * 7
    wide = executor.widen_for_pix_fmt(img, "rgb48le")
    assert wide.dtype == np.uint16 and np.array_equal(wide, img.astype(np.uint16) * 257)
    assert wide.max() <= 65535 and executor.widen_for_pix_fmt(np.full((1, 1, 3), 255, np.uint8), "rgb48le").max() == 65535
    assert executor.widen_for_pix_fmt(img, "rgb24") is img and executor.widen_for_pix_fmt(img, None) is img
    executor._write_image(tmp_path / "v.png", img, 100, "rgb48le")
    back = cv2.imread(str(tmp_path / "v.png"), cv2.IMREAD_UNCHANGED)
    assert back.dtype == np.uint16 and np.array_equal(back, wide)
    executor._write_image(tmp_path / "v.jpg", img, 100, "rgb48le")              # JPEG stays 8-bit
    assert cv2.imread(str(tmp_path / "v.jpg"), cv2.IMREAD_UNCHANGED).dtype == np.uint8
