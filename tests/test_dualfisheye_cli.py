"""The DualFisheye command line (remap360/dualfisheye_cli.py) against the reference's own behaviour recorded by
tests/golden/make_golden.py: every argparse action, stdout / stderr / exit code of dry runs and usage errors,
and (``-m gpu``) one real run whose output files are compared with the files the reference wrote."""

import contextlib
import io
import json
import os
import pathlib

import numpy as np
import pytest

pytest.importorskip("torch")
from remap360 import dualfisheye_cli as cli  # noqa: E402

GOLDEN = pathlib.Path(__file__).resolve().parent / "golden"


@pytest.fixture(scope="module")
def gold():
    return json.loads((GOLDEN / "df_cli.json").read_text()), np.load(GOLDEN / "df_cli_outputs.npz")


def _materialise(tmp: pathlib.Path, arrays) -> None:
    """Recreate the input tree the recorded runs used."""
    cv2 = pytest.importorskip("cv2")
    (tmp / "frames").mkdir()
    (tmp / "masks").mkdir()
    (tmp / "emptymasks").mkdir()
    for key in arrays.files:
        if key.startswith("in_"):
            cv2.imwrite(str(tmp / "frames" / (key[3:] + ".png")), arrays[key])
        elif key.startswith("mask_"):
            cv2.imwrite(str(tmp / "masks" / (key[5:] + ".png")), arrays[key])
    (tmp / "frames" / "notes.txt").write_text("not an image")
    (tmp / "look.cube").write_text(str(arrays["cube_text"]))
    (tmp / "cal.xml").write_text(str(arrays["cal_xml"]))


def _run(argv, tmp):
    out, err = io.StringIO(), io.StringIO()
    with contextlib.redirect_stdout(out), contextlib.redirect_stderr(err):
        code = cli.main([a.replace("<TMP>", str(tmp)) for a in argv])
    norm = lambda t: t.replace(str(tmp), "<TMP>")   # noqa: E731
    return code, norm(out.getvalue()), norm(err.getvalue())


def test_parser_has_every_reference_flag_with_the_same_defaults(gold):
    meta, _ = gold
    mine = {a.dest + "|" + ",".join(a.option_strings): a for a in cli.create_arg_parser()._actions if a.option_strings}
    for want in meta["actions"]:
        act = mine.get(want["dest"] + "|" + ",".join(want["flags"]))
        assert act is not None, want["flags"]
        assert type(act).__name__ == want["action"] and bool(act.required) == want["required"]
        assert (list(act.choices) if act.choices else None) == want["choices"]
        assert getattr(act.type, "__name__", None) == want["type"]
        if want["default"] == "<CPU_COUNT>":
            assert act.default == max(1, os.cpu_count() or 1)
        elif isinstance(want["default"], str) and want["default"].startswith("<TEMPLATES>/"):
            assert pathlib.Path(act.default).name == want["default"].split("/", 1)[1]
        else:
            assert act.default == want["default"], want["flags"]
    assert len(mine) - 1 == len(meta["actions"])            # nothing extra but -h/--help
    assert pathlib.Path(cli.DEFAULT_CAMERA_XML).is_file()


@pytest.mark.parametrize("name", ["dry_default", "dry_all_outputs", "err_no_input", "err_missing_dir", "err_all_disabled",
                                  "err_suffixes", "err_zoom", "err_workers", "err_masks_missing", "err_xml_missing",
                                  "err_color_profile_lut_missing"])
def test_dry_runs_and_usage_errors_match_the_reference_transcripts(gold, tmp_path, name):
    meta, arrays = gold
    tmp = tmp_path.resolve()
    _materialise(tmp, arrays)
    want = meta["runs"][name]
    code, out, err = _run(want["argv"], tmp)
    assert code == want["exit"]
    assert out == want["stdout"]
    assert err == want["stderr"]


def test_default_template_calibration_is_the_reference_template(golden_df):
    from remap360 import dualfisheye as dfh
    sensors, cams = dfh.load_metashape_calibration(cli.DEFAULT_CAMERA_XML)
    assert list(sensors) == ["0"] and cams == {}
    want = golden_df["sensors"]["0"]
    for key, value in want.items():
        assert getattr(sensors["0"], key) == value, key


def test_metadata_export_failure_is_reported_like_the_reference(gold, tmp_path):
    """An extrinsics XML without a <cameras> block: the export step fails with load_metashape_cameras' message, the
    run ends with errors=1 and exit code 2 (DF:2829-2849)."""
    meta, arrays = gold
    tmp = tmp_path.resolve()
    _materialise(tmp, arrays)
    code, out, err = _run(["--input-dir", "<TMP>/frames", "--camera-xml", "<TMP>/cal.xml", "--dry-run",
                           "--camera-extrinsics-xml", "<TMP>/cal.xml"], tmp)
    assert code == 2 and "errors=1" in out
    assert err == "[ERR] perspective camera metadata export failed (missing <cameras> in XML)\n"      # MS:549-551
    code, out, err = _run(["--metadata-only", "--camera-extrinsics-xml", "<TMP>/cal.xml"], tmp)
    assert code == 1 and err == "[ERR] --metadata-only requires --pointcloud-ply.\n"


@pytest.mark.gpu
def test_real_run_writes_what_the_reference_wrote(gold, tmp_path):
    """Two smooth 160 px X/Y pairs, LUT + sRGB, undistorted fisheye, ten 48 px views + masks, PNG, bilinear,
    mask value 7.  Transcript equal up to line order; images within 2 LSB on >= 99.5 % of the pixels (the
    reference evaluates its maps in float32, up to 0.05 px from the float64 maps used here; masks are
    nearest-neighbour, so they may differ where a coordinate sits on a pixel boundary)."""
    cv2 = pytest.importorskip("cv2")
    meta, arrays = gold
    tmp = tmp_path.resolve()
    _materialise(tmp, arrays)
    want = meta["runs"]["real"]
    code, out, err = _run(want["argv"], tmp)
    assert (code, err) == (0, "")
    assert sorted(out.splitlines()) == sorted(want["stdout"].splitlines())
    for sub, names in meta["real_files"].items():
        for fn in names:
            ref = arrays["out_%s_%s" % (sub.replace("/", "_"), fn)]
            got = cv2.imread(str(tmp / sub / fn), cv2.IMREAD_UNCHANGED)
            assert got is not None and got.shape == ref.shape and got.dtype == ref.dtype, (sub, fn)
            d = np.abs(got.astype(int) - ref.astype(int))
            if sub.endswith("Masks"):
                assert (d == 0).mean() >= 0.97, (sub, fn, (d == 0).mean())
            elif sub == "frames_colorcorrected":
                assert d.max() <= 1 and (d == 0).mean() >= 0.999, (sub, fn, d.max())
            else:
                assert (d <= 2).mean() >= 0.995, (sub, fn, (d <= 2).mean(), d.max())


@pytest.mark.gpu
def test_jpeg_frames_are_decoded_and_views_encoded_on_the_gpu(gold, tmp_path, monkeypatch):
    """JPEG X/Y frames in, JPEG views out: the nvJPEG run and the OpenCV-codec run give the same views to JPEG
    accuracy, and both report the reference's lines."""
    cv2 = pytest.importorskip("cv2")
    meta, arrays = gold
    tmp = tmp_path.resolve()
    _materialise(tmp, arrays)
    for p in sorted((tmp / "frames").glob("*.png")):
        cv2.imwrite(str(p.with_suffix(".jpg")), cv2.imread(str(p)), [cv2.IMWRITE_JPEG_QUALITY, 97])
        p.unlink()
    base = ["--input-dir", "<TMP>/frames", "--camera-xml", "<TMP>/cal.xml", "--perspective-size", "64", "--workers", "2"]
    outs = {}
    for mode in ("gpu", "cpu"):
        if mode == "cpu":
            monkeypatch.setenv("R360_CPU_CODEC", "1")
        code, out, err = _run(base + ["--perspective-output-dir", "<TMP>/out_" + mode], tmp)
        assert (code, err) == (0, "") and "[DONE] processed=4 skipped=0 total=4 persp_outputs=20" in out
        outs[mode] = {p.name: cv2.imread(str(p)) for p in sorted((tmp / ("out_" + mode) / "Images").glob("*.jpg"))}
        assert len(outs[mode]) == 20
    for name, img in outs["gpu"].items():
        other = outs["cpu"][name]
        mse = np.mean((img.astype(float) - other.astype(float)) ** 2)
        assert img.shape == other.shape == (64, 64, 3) and 10 * np.log10(255.0 ** 2 / max(mse, 1e-9)) > 36.0, name
