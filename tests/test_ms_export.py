"""The MS360xmlToPersCams pose exporters (remap360/ms_export.py) against whole runs of the reference recorded by
tests/golden/make_golden.py (tests/golden/ms_export.json): view sets and intrinsics of every preset, then ``main()``
with the recorded arguments -- stdout, stderr, exit code and every written file byte for byte.  Host code only."""

import base64
import contextlib
import io
import json
import pathlib
import sys

import pytest

from remap360 import ms_export as ms


@pytest.fixture(scope="module")
def golden(golden_dir):
    return json.loads((golden_dir / "ms_export.json").read_text())


def test_view_sets_and_intrinsics_of_every_preset(golden):
    assert sorted(golden["presets"]) == sorted(ms.PRESET_CHOICES)
    for preset, want in golden["presets"].items():
        assert [list(v) for v in ms.build_views(preset)] == want["views"], preset
        size, focal = ms.preset_size_and_focal(preset)
        assert size == want["size"] and focal == want["focal_mm"], preset
        assert list(ms.compute_intrinsics(focal, size, size)) == want["intrinsics"], preset           # float for float


def test_view_sets_agree_with_the_cutter(golden):
    """The exporter's view ids / angles are the cutter's own (remap360.perspcut) for the presets both know."""
    from remap360 import perspcut as pc
    for preset in ("default", "fisheyelike", "full360coverage", "2views", "evenMinus30", "evenPlus30"):
        args = pc.create_arg_parser().parse_args(["-i", "in", "--preset", preset])
        args.size_explicit = args.hfov_explicit = args.focal_mm_explicit = False
        args.input_is_video, args.video_bit_depth = False, 8
        res = pc.build_view_jobs(args, [pathlib.Path("in/pano.jpg")], pathlib.Path("out"))
        got = sorted((s.view_id, round(s.yaw_deg, 9), round(s.pitch_deg, 9)) for s in res.view_specs)
        want = sorted((vid, round(float(y), 9), round(float(p), 9)) for vid, y, p in ms.build_views(preset))
        assert got == want, preset


def _run(argv):
    out, err, code = io.StringIO(), io.StringIO(), 0
    with contextlib.redirect_stdout(out), contextlib.redirect_stderr(err):
        try:
            ms.main(argv)
        except SystemExit as exc:
            code = exc.code if isinstance(exc.code, int) else 1
    return out.getvalue(), err.getvalue(), code


def test_whole_runs_equal_the_reference(golden, tmp_path):
    for name, payload in golden["inputs"].items():
        if name.endswith(".ply"):
            (tmp_path / name).write_bytes(base64.b64decode(payload))
        else:
            (tmp_path / name).write_text(payload)
    norm = lambda t: t.replace(str(tmp_path), "<TMP>")
    for run in golden["runs"]:
        argv = [a.replace("<TMP>", str(tmp_path)) for a in run["argv"]]
        before = {f for f in tmp_path.rglob("*") if f.is_file()}
        out, err, code = _run(argv)
        label = " ".join(run["argv"])
        assert code == run["exit"], label
        assert norm(err) == run["stderr"], label
        assert norm(out) == run["stdout"], label
        written = sorted(set(f for f in tmp_path.rglob("*") if f.is_file()) - before)
        assert [str(f.relative_to(tmp_path)) for f in written] == sorted(run["files"]), label
        for f in written:
            want = run["files"][str(f.relative_to(tmp_path))]
            if want.startswith("base64:"):
                assert f.read_bytes() == base64.b64decode(want[7:]), (label, f.name)
            else:
                assert f.read_text() == want, (label, f.name)
            f.unlink()


def test_multi_camera_system_format_is_refused_not_faked(tmp_path, golden):
    (tmp_path / "cameras.xml").write_text(golden["inputs"]["cameras.xml"])
    _out, err, code = _run([str(tmp_path / "cameras.xml"), "--format", "metashape-multi-camera-system", "--preset", "fisheyelike"])
    assert code == 1 and "not available" in err


def test_cut_command_is_the_references(tmp_path):
    tool = pathlib.Path("gs360_360PerspCut.py")
    assert ms.cut_command("fisheyelike", tmp_path, None, tool) == [sys.executable, str(tool), "-i", str(tmp_path), "--preset", "fisheyelike"]
    assert ms.cut_command("cube105", tmp_path, tmp_path / "o", tool) == [sys.executable, str(tool), "-i", str(tmp_path), "--count", "4",
                                                                        "--hfov", "105.0", "--add-top", "--add-bottom", "-o", str(tmp_path / "o")]
