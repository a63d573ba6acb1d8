"""GPU end-to-end checks of the host layer: the drop-in job runner writes the right files, the
dual-fisheye stage picks the reference's lenses and renders what the oracle renders, the streaming
remapper returns frames in order.  Run with ``pytest -m gpu``."""

import os
import pathlib

import numpy as np
import pytest

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu

from oracle import geometry as geo  # noqa: E402
from oracle import sampler  # noqa: E402


@pytest.fixture(scope="module")
def r360():
    import remap360
    if not torch.cuda.is_available():
        pytest.fail("-m gpu tests need a CUDA device")
    return remap360


def test_run_jobs_writes_every_view_of_a_still(r360, tmp_path):
    cv2 = pytest.importorskip("cv2")
    from remap360 import executor, perspcut as pc
    rng = np.random.default_rng(5)
    src = rng.integers(0, 256, (256, 512, 3), dtype=np.uint8)
    (tmp_path / "in").mkdir()
    cv2.imwrite(str(tmp_path / "in" / "pano0001.png"), src)
    args = pc.create_arg_parser().parse_args(["-i", str(tmp_path / "in"), "--preset", "full360coverage",
                                              "--size", "96", "--ext", "png"])
    args.size_explicit, args.hfov_explicit, args.focal_mm_explicit = True, False, False
    args.input_is_video, args.video_bit_depth = False, 8
    res = pc.build_view_jobs(args, [tmp_path / "in" / "pano0001.png"], tmp_path / "out")
    done = list(executor.run_jobs(res.jobs, pc.stop_event, workers=2))
    assert len(done) == 12 and all(rc == 0 for _job, (rc, _err) in done)
    for spec in res.view_specs:
        got = cv2.imread(str(tmp_path / "out" / spec.output_name), cv2.IMREAD_UNCHANGED)
        mx, my = geo.erp_map64(512, 256, 96, 96, spec.yaw_deg, spec.pitch_deg, spec.hfov_deg, spec.vfov_deg)
        want = sampler.sample(src, mx, my, "cubic", "erp")            # PNG is lossless; BGR in, BGR out
        assert got.shape == want.shape
        assert (np.abs(got.astype(int) - want.astype(int)) <= 1).mean() >= 0.999
    # the single-job entry point (what the GUI submits to its thread pool)
    rc, err = pc.run_one(res.jobs[3][0])
    assert (rc, err) == (0, "")
    rc, err = pc.run_one(["ffmpeg", "-i", str(tmp_path / "missing.png"), "-vf",
                          "v360=input=equirect:output=rectilinear:w=8:h=8:yaw=0:pitch=0:roll=0:h_fov=90:v_fov=90:interp=cubic",
                          str(tmp_path / "x.png")])
    assert rc == 1 and "failed to read" in err


@pytest.mark.parametrize("cpu_codec", [False, True])
def test_jpeg_in_jpeg_out_through_the_gpu_codec(r360, tmp_path, monkeypatch, cpu_codec):
    """JPEG panorama -> JPEG views: decoded / encoded by nvJPEG (default) or OpenCV (R360_CPU_CODEC); both
    must give the oracle's views to JPEG accuracy, and the dual-fisheye CLI writes decodable JPEG views."""
    cv2 = pytest.importorskip("cv2")
    from remap360 import executor, perspcut as pc
    if cpu_codec:
        monkeypatch.setenv("R360_CPU_CODEC", "1")
    else:
        assert executor._gpu_codec() is not None
    yy, xx = np.mgrid[0:512, 0:1024].astype(np.float64)
    src = np.stack([127 + 90 * np.sin(xx / 1024 * 6.2832 * (k + 1)) * np.cos(yy / 512 * 3.1416 * (k + 1)) for k in range(3)],
                   axis=-1).astype(np.uint8)
    (tmp_path / "in").mkdir()
    cv2.imwrite(str(tmp_path / "in" / "pano.jpg"), src, [cv2.IMWRITE_JPEG_QUALITY, 98])
    decoded = cv2.imread(str(tmp_path / "in" / "pano.jpg"))
    args = pc.create_arg_parser().parse_args(["-i", str(tmp_path / "in"), "--size", "200"])
    args.size_explicit, args.hfov_explicit, args.focal_mm_explicit = True, False, False
    args.input_is_video, args.video_bit_depth = False, 8
    res = pc.build_view_jobs(args, [tmp_path / "in" / "pano.jpg"], tmp_path / "out")
    done = list(executor.run_jobs(res.jobs, pc.stop_event, workers=1))
    assert len(done) == 8 and all(rc == 0 for _job, (rc, _err) in done), done
    for spec in res.view_specs:
        got = cv2.imread(str(tmp_path / "out" / spec.output_name))
        mx, my = geo.erp_map64(1024, 512, 200, 200, spec.yaw_deg, spec.pitch_deg, spec.hfov_deg, spec.vfov_deg)
        want = sampler.sample(decoded, mx, my, "cubic", "erp")
        mse = np.mean((got.astype(float) - want.astype(float)) ** 2)
        assert got.shape == want.shape and 10 * np.log10(255.0 ** 2 / max(mse, 1e-9)) > 38.0


def test_fisheye_xy_preset_jobs_run(r360, tmp_path):
    """Preset fisheyeXY: two `output=fisheye:d_fov=180` jobs per panorama (PC:871-887)."""
    cv2 = pytest.importorskip("cv2")
    from remap360 import executor, perspcut as pc
    rng = np.random.default_rng(6)
    src = rng.integers(0, 256, (256, 512, 3), dtype=np.uint8)
    (tmp_path / "in").mkdir()
    cv2.imwrite(str(tmp_path / "in" / "pano0001.png"), src)
    args = pc.create_arg_parser().parse_args(["-i", str(tmp_path / "in"), "--preset", "fisheyeXY", "--ext", "png"])
    args.size_explicit = args.hfov_explicit = args.focal_mm_explicit = False
    args.input_is_video, args.video_bit_depth = False, 8
    res = pc.build_view_jobs(args, [tmp_path / "in" / "pano0001.png"], tmp_path / "out")
    jobs = []
    for argv, src_name, dst_name in res.jobs:                    # shrink the 3600 px outputs for the test
        argv = [a.replace("w=3600:h=3600", "w=120:h=120") for a in argv]
        jobs.append((argv, src_name, dst_name))
    done = list(executor.run_jobs(jobs, pc.stop_event, workers=1))
    assert len(done) == 2 and all(rc == 0 for _job, (rc, _err) in done), done
    hf, vf = geo.fisheye_fov_from_dfov(180.0, 120, 120)
    for spec in res.view_specs:
        got = cv2.imread(str(tmp_path / "out" / spec.output_name), cv2.IMREAD_UNCHANGED)
        mx, my = geo.erp_map64(512, 256, 120, 120, spec.yaw_deg, spec.pitch_deg, hf, vf, projection="fisheye")
        want = sampler.sample(src, mx, my, "cubic", "erp")
        assert got.shape == want.shape
        assert (np.abs(got.astype(int) - want.astype(int)) <= 1).mean() >= 0.999


def test_dualfisheye_stage_matches_reference_lens_choice_and_oracle(r360, golden_df):
    from remap360 import dualfisheye as dfh
    cal = golden_df["sensors"]["0"]
    sc = dfh.SensorCalibration(**{k: cal[k] for k in ("sensor_id", "model_type", "width", "height", "f", "cx", "cy",
                                                      "k1", "k2", "k3", "k4", "p1", "p2", "b1", "b2")})
    specs = dfh.build_sfm10_specs(160, 14.0, "36 36", 40.0, 40.0)
    views, cals, info = dfh.choose_lenses(sc, sc, specs)
    want_lens = {vid: v["lens_key"] for vid, v in golden_df["maps_1750"]["views"].items()}
    assert {vid: i["lens_key"] for vid, i in info.items()} == want_lens
    assert all(i["valid_ratio"] == 1.0 for i in info.values())
    # render a small pair and compare with the oracle on the chosen lens
    rng = np.random.default_rng(9)
    pair = rng.integers(0, 256, (1, 2, 3840, 3840, 3), dtype=np.uint8)
    out = r360.remap_fisheye(torch.from_numpy(pair).cuda(), cals, views, (160, 160), interp="cubic")[0].cpu().numpy()
    for k in (0, 4, 9):
        v = views[k]
        mx, my, ok = geo.fisheye_map64(cal, v.yaw_deg, v.pitch_deg, v.hfov_deg, v.vfov_deg, 160, 160, 190.0)
        want = sampler.apply_invalid_fill(sampler.sample(pair[0, v.src_slot], mx, my, "cubic", "constant", 0), ok, 0)
        assert (np.abs(out[k].astype(int) - want.astype(int)) <= 1).mean() >= 0.999


def test_streaming_remapper_keeps_order_and_matches_batched_call(r360):
    from remap360.stream import StreamingRemapper
    rng = np.random.default_rng(2)
    frames = [torch.from_numpy(rng.integers(0, 256, (128, 256, 3), dtype=np.uint8)) for _ in range(7)]
    views = [r360.PerspectiveView(y, p, 90.0, 90.0) for y, p in ((0, 0), (120, 20), (-100, -35))]
    sr = StreamingRemapper(views, (64, 64), (128, 256, 3), torch.uint8, interp="linear", depth=3)
    got = [res.clone() for res in sr.run(frames)]           # pageable inputs are staged through pinned memory
    assert len(got) == 7 and sr.frames_done == 7
    ref = r360.remap_erp(torch.stack(frames).cuda(), views, (64, 64), interp="linear").cpu()
    for k in range(7):
        assert torch.equal(got[k], ref[k])
    pinned = [f.pin_memory() for f in frames[:3]]
    again = [res.clone() for res in sr.run(pinned)]
    assert all(torch.equal(again[k], ref[k]) for k in range(3))


@pytest.mark.parametrize("ext,depth,keep,decoder", [("png", 8, False, "opencv"), ("png", 10, True, "opencv"), ("jpg", 8, False, "opencv"),
                                                    ("png", 8, False, "nvjpeg")])
def test_video_source_is_decoded_once_colour_converted_and_cut(r360, tmp_path, monkeypatch, ext, depth, keep, decoder):
    """A video source (PC's video branch): frames picked with ffmpeg's fps= rule, the job's colorspace filter applied
    on the device before the remap, views written as <stem>_%07d_<id>.<ext> from 0; a >8-bit source asks for
    rgb48le (16-bit PNG holding the widened 8-bit result)."""
    cv2 = pytest.importorskip("cv2")
    from remap360 import color, executor, perspcut as pc, video
    # "opencv": frames decoded on the host (bit-identical to the frames this test decodes itself); "nvjpeg": the
    # Motion-JPEG packets decoded on the device (another IDCT: the source frames differ by a level or two)
    monkeypatch.setenv("R360_VIDEO_DECODER", "opencv" if decoder == "opencv" else "auto")
    path = tmp_path / "clip.avi"
    wr = cv2.VideoWriter(str(path), cv2.VideoWriter_fourcc(*"MJPG"), 12.0, (512, 256))
    if not wr.isOpened():
        pytest.skip("OpenCV cannot write MJPG AVI here")
    yy, xx = np.mgrid[0:256, 0:512].astype(np.float64)
    for n in range(12):
        frame = np.stack([127 + 100 * np.sin((xx + 17 * n) / 512 * 6.2832 * (k + 1)) * np.cos(yy / 256 * 3.1416 * (k + 1))
                          for k in range(3)], axis=-1).astype(np.uint8)
        wr.write(frame)
    wr.release()
    cap = cv2.VideoCapture(str(path))
    decoded = []
    while True:
        ok, fr = cap.read()
        if not ok:
            break
        decoded.append(fr)
    cap.release()
    assert len(decoded) == 12
    argv = ["-i", str(path), "--preset", "2views", "--size", "80", "--ext", ext, "-f", "4"] + (["--keep-rec709"] if keep else [])
    args = pc.create_arg_parser().parse_args(argv)
    args.size_explicit, args.hfov_explicit, args.focal_mm_explicit = True, False, False
    args.input_is_video, args.video_bit_depth = True, depth
    res = pc.build_view_jobs(args, [path], tmp_path / "out")
    assert all("colorspace=iall=bt709:all=smpte170m" in " ".join(cmd) for cmd, _s, _d in res.jobs)
    done = list(executor.run_jobs(res.jobs, pc.stop_event, workers=1))
    assert len(done) == len(res.jobs) and all(rc == 0 for _job, (rc, err) in done), done
    picked = video._select_frames(12, 12.0, 4.0, None, None)
    assert picked == [0, 3, 6, 9] or len(picked) in (4, 5)
    for spec in res.view_specs:
        for n, src_idx in enumerate(picked):
            name = spec.output_name % n if "%" in spec.output_name else spec.output_name
            got = cv2.imread(str(tmp_path / "out" / name), cv2.IMREAD_UNCHANGED)
            assert got is not None, name
            # the decoder's BT.601 R'G'B' re-matrixed to the BT.709 the filter declares, then the filter itself
            fixed = color.correct_decoder_matrix(torch.from_numpy(decoded[src_idx]).cuda()[None], channel_order="bgr")
            conv = color.convert_video_color(fixed, keep_rec709=keep, channel_order="bgr")[0].cpu().numpy()
            mx, my = geo.erp_map64(512, 256, 80, 80, spec.yaw_deg, spec.pitch_deg, spec.hfov_deg, spec.vfov_deg)
            want = sampler.sample(conv, mx, my, "cubic", "erp")
            if ext == "png" and depth > 8:
                assert got.dtype == np.uint16 and np.array_equal(got % 257, np.zeros_like(got))
                got = (got // 257).astype(np.uint8)
            assert got.dtype == np.uint8 and got.shape == want.shape
            tol = (1 if decoder == "opencv" else 6) if ext == "png" else 12      # JPEG views / another decoder: codec error on top
            assert (np.abs(got.astype(int) - want.astype(int)) <= tol).mean() >= 0.99, (name, ext)
    assert not (tmp_path / "out" / (res.view_specs[0].output_name % len(picked))).exists()
    # the colour step can be switched off: frames then stay in the decoder's BGR values
    monkeypatch.setenv("R360_VIDEO_COLOR", "0")
    done = list(executor.run_jobs(res.jobs[:1], pc.stop_event, workers=1))
    assert done[0][1][0] == 0


def test_sixteen_bit_still_to_jpeg_view_is_scaled_not_saturated(r360, tmp_path):
    """A 16-bit TIFF source cut into .jpg views (the default --ext): the views hold the scaled 8-bit picture, on the
    nvJPEG path and on the OpenCV one (cv2.imwrite alone would saturate every sample >= 255 to 255)."""
    cv2 = pytest.importorskip("cv2")
    from remap360 import executor, perspcut as pc
    src_dir = tmp_path / "in"
    src_dir.mkdir()
    yy, xx = np.mgrid[0:256, 0:512].astype(np.float64)
    img = np.stack([20000 + 15000 * np.sin(xx / 512 * 6.2832 * (k + 1)) * np.cos(yy / 256 * 3.1416) for k in range(3)], axis=-1).astype(np.uint16)
    assert cv2.imwrite(str(src_dir / "pano.tif"), img)
    args = pc.create_arg_parser().parse_args(["-i", str(src_dir), "--preset", "2views", "--size", "96"])
    args.size_explicit, args.hfov_explicit, args.focal_mm_explicit = True, False, False
    args.input_is_video, args.video_bit_depth = False, 8
    res = pc.build_view_jobs(args, [src_dir / "pano.tif"], tmp_path / "out")
    for cpu_codec in ("", "1"):
        if cpu_codec:
            os.environ["R360_CPU_CODEC"] = "1"
        try:
            done = list(executor.run_jobs(res.jobs, pc.stop_event, workers=1))
        finally:
            os.environ.pop("R360_CPU_CODEC", None)
        assert all(rc == 0 for _j, (rc, _e) in done), done
        for spec in res.view_specs:
            got = cv2.imread(str(tmp_path / "out" / spec.output_name), cv2.IMREAD_UNCHANGED)
            mx, my = geo.erp_map64(512, 256, 96, 96, spec.yaw_deg, spec.pitch_deg, spec.hfov_deg, spec.vfov_deg)
            want16 = sampler.sample(img, mx, my, "cubic", "erp")
            want8 = executor.narrow_to_8bit(want16)
            assert got.dtype == np.uint8 and (np.abs(got.astype(int) - want8.astype(int)) <= 6).mean() >= 0.99, spec.output_name
            assert 40 < float(got.mean()) < 120                # the saturated picture would sit at ~255


def test_two_worker_processes_write_what_one_process_writes(r360, tmp_path):
    """remap360.multigpu: still sources dealt out to one worker process per device, a video split by frame ranges.
    On a one-GPU box both workers share the device; the files must be the same as a single-process run's."""
    cv2 = pytest.importorskip("cv2")
    from remap360 import multigpu, perspcut as pc
    src_dir = tmp_path / "in"
    src_dir.mkdir()
    rng = np.random.default_rng(3)
    for k in range(5):
        assert cv2.imwrite(str(src_dir / ("p%d.png" % k)), rng.integers(0, 256, (256, 512, 3), dtype=np.uint8))
    args = pc.create_arg_parser().parse_args(["-i", str(src_dir), "--preset", "2views", "--size", "64", "--ext", "png"])
    args.size_explicit, args.hfov_explicit, args.focal_mm_explicit = True, False, False
    args.input_is_video, args.video_bit_depth = False, 8
    files = sorted(src_dir.iterdir())
    outs = {}
    for world in (1, 2):
        out_dir = tmp_path / ("out%d" % world)
        res = pc.build_view_jobs(args, files, out_dir)
        done = list(multigpu.run_jobs(res.jobs, None, 1, devices=world))
        assert len(done) == len(res.jobs) and all(rc == 0 for _j, (rc, _e) in done), done
        outs[world] = {p.name: p.read_bytes() for p in sorted(out_dir.iterdir())}
    assert len(outs[1]) == 10 and outs[1] == outs[2]


def test_ms_tool_persp_cut_runs_the_cuda_cutter(tmp_path):
    """gs360_MS360xmlToPersCams --persp-cut (MS:2019-2051): poses are exported, then the cutter next to the tool is
    spawned on <xml_dir>/360imgs -- here that cutter is the CUDA one; the cut files carry the names the poses refer to."""
    import json
    import subprocess
    import sys
    cv2 = pytest.importorskip("cv2")
    golden = json.loads((pathlib.Path(__file__).parent / "golden" / "ms_export.json").read_text())
    (tmp_path / "cameras.xml").write_text(golden["inputs"]["cameras.xml"])
    (tmp_path / "360imgs").mkdir()
    yy, xx = np.mgrid[0:256, 0:512].astype(np.float32)
    pano = np.stack([127 + 100 * np.sin(xx / 512 * 6.2832 * (k + 1)) * np.cos(yy / 256 * 3.1416) for k in range(3)], axis=-1).astype(np.uint8)
    cv2.imwrite(str(tmp_path / "360imgs" / "IMG_0001.jpg"), pano)
    tool = pathlib.Path(__file__).resolve().parent.parent / "360cam-pgm-3dgs-tools_b200" / "gs360_MS360xmlToPersCams.py"
    proc = subprocess.run([sys.executable, str(tool), str(tmp_path / "cameras.xml"), "--preset", "default", "--format", "transforms",
                           "--persp-cut", "--cut-out", str(tmp_path / "cut")], capture_output=True, text=True, timeout=600)
    assert proc.returncode == 0, proc.stderr[-2000:]
    assert "[INFO] Running cut:" in proc.stdout
    frames = json.loads((tmp_path / "perspective_cams" / "transforms.json").read_text())["frames"]
    names = {f["file_path"] for f in frames if f["file_path"].startswith("IMG_0001_")}
    cut = {p.name for p in (tmp_path / "cut").glob("*.jpg")}
    assert len(cut) == 8 and cut == names
    img = cv2.imread(str(tmp_path / "cut" / "IMG_0001_A.jpg"))
    assert img is not None and img.shape == (1600, 1600, 3)
