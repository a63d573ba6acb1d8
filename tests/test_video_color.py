"""The cutter's video colour step, ``colorspace=iall=bt709:all=smpte170m[:trc=iec61966-2-1]`` (PC:299-309,
V2F:462-464): host-side parsing and matrix on the CPU, the ``r360_convert_color`` kernel against the
float64 statement of the formula on the GPU.  ffmpeg itself is not in this image, so the
bar is the formula: integer outputs within 1 LSB, >= 99.9 % equal."""

import pathlib

import numpy as np
import pytest

from oracle import color as ocolor

torch = pytest.importorskip("torch")


def test_primaries_matrix_properties():
    from remap360 import color
    m = color.primaries_matrix("bt709", "smpte170m")
    assert np.allclose(m.sum(axis=1), 1.0, atol=1e-12)          # D65 white stays white
    assert np.allclose(m, np.linalg.solve(ocolor.xyz_from_rgb64(ocolor.SMPTE170M_PRIMARIES),
                                          ocolor.xyz_from_rgb64(ocolor.BT709_PRIMARIES)), atol=1e-14)
    assert np.allclose(color.primaries_matrix("bt709", "bt709"), np.eye(3))
    back = color.primaries_matrix("smpte170m", "bt709")
    assert np.allclose(back @ m, np.eye(3), atol=1e-12)
    # published BT.709 RGB -> XYZ matrix (first row), to 4 decimals
    assert np.allclose(color.rgb_to_xyz_matrix("bt709")[0], [0.4124, 0.3576, 0.1805], atol=5e-4)


def test_filter_parsing_follows_the_cutters_strings():
    from remap360 import color, executor, perspcut as pc
    assert color.parse_colorspace_filter("colorspace=iall=bt709:all=smpte170m:trc=iec61966-2-1:range=jpeg:format=yuv444p") \
        == ("bt709", "bt709", "smpte170m", "iec61966-2-1")
    assert color.parse_colorspace_filter("colorspace=iall=bt709:all=smpte170m:format=yuv444p") \
        == ("bt709", "bt709", "smpte170m", "smpte170m")
    with pytest.raises(ValueError):
        color.parse_colorspace_filter("colorspace=iall=bt2020:all=smpte170m")
    with pytest.raises(ValueError):
        color.parse_colorspace_filter("v360=input=equirect")
    # the job runner finds the filter in the argv the planner emits for a video source
    for keep, want in ((False, "iec61966-2-1"), (True, "smpte170m")):
        argv = pc.build_ffmpeg_cmd("ffmpeg", pathlib.Path("/in/v.mp4"), pathlib.Path("/out/v_%07d_A.jpg"), 1600, 1600,
                                   0.0, 0.0, 100.0, 100.0, "cubic", ".jpg", video_mode=True, fps=2.0, keep_rec709=keep)
        job = executor.parse_job_argv(argv)
        assert job.colorspace is not None and color.parse_colorspace_filter(job.colorspace)[3] == want
    still = pc.build_ffmpeg_cmd("ffmpeg", pathlib.Path("/in/a.jpg"), pathlib.Path("/out/a_A.jpg"), 1600, 1600,
                                0.0, 0.0, 100.0, 100.0, "cubic", ".jpg")
    assert executor.parse_job_argv(still).colorspace is None


def test_oracle_curves_are_inverse_pairs_and_continuous():
    v = np.linspace(-0.2, 1.2, 2801)
    for name in ("bt709", "srgb", "linear"):
        assert np.allclose(ocolor.trc_encode64(name, ocolor.trc_decode64(name, v)), v, atol=2e-7)
    # continuity at the knee of each curve
    for name, knee in (("bt709", 4.5 * 0.018053968510807), ("srgb", 0.04045)):
        lo, hi = ocolor.trc_decode64(name, np.array([knee - 1e-9, knee + 1e-9]))
        assert abs(hi - lo) < 1e-6
    # keep-rec709: curves cancel, only the primaries move; grey stays grey
    grey = np.full((4, 4, 3), 128, dtype=np.uint8)
    assert np.array_equal(ocolor.video_color_step64(grey, "smpte170m"), grey)


# ---------------------------------------------------------------------------------------------------------

@pytest.mark.gpu
@pytest.mark.parametrize("dtype", [np.uint8, np.uint16, np.float32])
@pytest.mark.parametrize("keep_rec709", [False, True])
@pytest.mark.parametrize("order", ["bgr", "rgb"])
def test_kernel_against_formula(dtype, keep_rec709, order):
    from remap360 import color
    rng = np.random.default_rng(5)
    shape = (2, 37, 53, 4 if dtype == np.uint8 else 3)
    if dtype == np.float32:
        img = rng.random(shape, dtype=np.float32) * 1.2 - 0.1
    else:
        img = rng.integers(0, np.iinfo(dtype).max + 1, shape).astype(dtype)
    if dtype == np.uint16:
        dev = torch.from_numpy(img.view(np.int16)).cuda().view(torch.uint16)
    else:
        dev = torch.from_numpy(img).cuda()
    got_t = color.convert_video_color(dev, keep_rec709=keep_rec709, channel_order=order)
    got = got_t.view(torch.int16).cpu().numpy().view(np.uint16) if dtype == np.uint16 else got_t.cpu().numpy()
    rgb = img.copy()
    if order == "bgr":
        rgb[..., :3] = img[..., 2::-1]
    want = ocolor.video_color_step64(rgb, "smpte170m" if keep_rec709 else "iec61966-2-1")
    if order == "bgr":
        want[..., :3] = want[..., 2::-1].copy()
    if shape[-1] == 4:
        assert np.array_equal(got[..., 3], img[..., 3])          # extra channel copied through
    if dtype == np.float32:
        assert np.abs(got[..., :3] - want[..., :3]).max() < 2e-5
    else:
        d = np.abs(got[..., :3].astype(np.int64) - want[..., :3].astype(np.int64))
        assert d.max() <= (1 if dtype == np.uint8 else 8)
        assert (d == 0).mean() >= (0.999 if dtype == np.uint8 else 0.5)
    # in place
    again = color.convert_video_color(dev, keep_rec709=keep_rec709, channel_order=order, out=dev)
    assert again.data_ptr() == dev.data_ptr() and torch.equal(again, got_t)


@pytest.mark.gpu
def test_streaming_remapper_applies_the_frame_filter():
    """Colour step then remap on the kernel stream == the two calls made one after the other."""
    import remap360
    from remap360 import color
    from remap360.stream import StreamingRemapper
    rng = np.random.default_rng(9)
    frames = [torch.from_numpy(rng.integers(0, 256, (256, 512, 3), dtype=np.uint8)) for _ in range(4)]
    views = [remap360.PerspectiveView(0.0, 0.0, 90.0, 90.0), remap360.PerspectiveView(120.0, -20.0, 90.0, 90.0)]

    def flt(dev_frame, stream):
        color.convert_video_color(dev_frame, channel_order="bgr", out=dev_frame, stream=stream)

    rem = StreamingRemapper(views, (96, 96), (256, 512, 3), torch.uint8, frame_filter=flt)
    got = [o.clone() for o in rem.run(iter(frames))]
    for f, g in zip(frames, got):
        want = remap360.remap_erp(color.convert_video_color(f.cuda()[None], channel_order="bgr"), views, (96, 96))[0].cpu()
        assert torch.equal(g, want)
