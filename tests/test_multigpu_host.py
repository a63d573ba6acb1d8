"""Host logic of the multi-GPU job runner and of the video branch (no GPU): the sharding plan, the 2-rank union
property under a real gloo process group, result merging, ffmpeg's -ss / -to / fps frame selection, the stream
bit-depth tags, the 16 -> 8 bit narrowing of JPEG views and the decoder matrix correction."""

import os
import pathlib
import socket
import sys

import numpy as np
import pytest

from conftest import PKG_DIR, ROOT
from remap360 import color, executor, multigpu, perspcut as pc, video


def _jobs(argv, files, video_mode=False):
    args = pc.create_arg_parser().parse_args(argv)
    for name in ("size", "hfov", "focal_mm"):
        setattr(args, name + "_explicit", getattr(args, name + "_explicit", False))
    args.input_is_video, args.video_bit_depth = video_mode, 8
    return pc.build_view_jobs(args, [pathlib.Path(f) for f in files], pathlib.Path("/tmp/out")).jobs


def _mixed_jobs():
    stills = _jobs(["-i", "/tmp/in", "--preset", "default"], ["/tmp/in/%02d.jpg" % k for k in range(7)])
    vid = _jobs(["-i", "/tmp/in/clip.mp4", "-f", "5", "--preset", "2views", "--start", "1", "--end", "6"], ["/tmp/in/clip.mp4"], True)
    return stills + vid, {"/tmp/in/clip.mp4": (300, 30.0)}


def test_plan_covers_every_job_once_and_keeps_sources_together():
    jobs, counts = _mixed_jobs()
    for world in (1, 2, 3, 8, 16):
        plan = multigpu.plan_shards(jobs, world)
        stills = [n for p in plan for n in p["stills"]]
        assert sorted(stills) == [n for n, j in enumerate(jobs) if not executor.parse_job_argv(j[0]).video]
        for p in plan:                                     # all views of a source on one rank, sources contiguous
            srcs = [str(executor.parse_job_argv(jobs[n][0]).source) for n in p["stills"]]
            assert srcs == sorted(srcs) and all(srcs.count(s) == 8 for s in set(srcs))
        names = [o for p in plan for o in multigpu.expected_outputs(jobs, p, counts)]
        assert len(names) == len(set(names))               # no output written twice
        assert sorted(names) == sorted(multigpu.expected_outputs(jobs, multigpu.plan_shards(jobs, 1)[0], counts))


def test_video_shards_are_contiguous_ranges_of_global_frame_numbers():
    jobs, counts = _mixed_jobs()
    vid = [j for j in jobs if executor.parse_job_argv(j[0]).video]
    total = len(video._select_frames(300, 30.0, 5.0, 1.0, 6.0))
    assert total == 31                                     # source times [1, 7] at 5 fps
    seen = []
    for p in multigpu.plan_shards(vid, 4):
        outs = multigpu.expected_outputs(vid, p, counts)
        nums = sorted({int(pathlib.Path(o).name.split("_")[1]) for o in outs})
        assert nums == list(range(nums[0], nums[-1] + 1))
        seen += nums
    assert seen == list(range(total))


def test_merge_results_prefers_failures():
    merged = multigpu.merge_results([{0: (0, ""), 1: (0, "")}, {1: (1, "boom"), 2: (130, "")}, {2: (0, "")}], 4)
    assert merged[0] == (0, "") and merged[1] == (1, "boom") and merged[2] == (130, "") and merged[3][0] == 1


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _gloo_rank(rank, world, port, tmp):
    import torch.distributed as dist
    for p in (str(ROOT), str(PKG_DIR)):
        if p not in sys.path:
            sys.path.insert(0, p)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        jobs, counts = _mixed_jobs()
        mine = multigpu.expected_outputs(jobs, multigpu.plan_shards(jobs, world)[rank], counts)
        for name in mine:                                   # stand-in for the render: one file per output
            path = pathlib.Path(tmp) / pathlib.Path(name).name
            assert not path.exists(), name
            path.write_text(str(rank))
        gathered = [None] * world
        dist.all_gather_object(gathered, mine)
        if rank == 0:
            single = multigpu.expected_outputs(jobs, multigpu.plan_shards(jobs, 1)[0], counts)
            union = [o for part in gathered for o in part]
            assert sorted(union) == sorted(single) and len(union) == len(set(union))
            assert all(part for part in gathered)           # every rank got work
        dist.barrier()
    finally:
        dist.destroy_process_group()


def test_two_ranks_under_gloo_produce_the_single_rank_output_set(tmp_path):
    mp = pytest.importorskip("torch.multiprocessing")
    port = _free_port()
    mp.spawn(_gloo_rank, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    jobs, counts = _mixed_jobs()
    single = multigpu.expected_outputs(jobs, multigpu.plan_shards(jobs, 1)[0], counts)
    assert sorted(p.name for p in tmp_path.iterdir()) == sorted(pathlib.Path(o).name for o in single)


def test_frame_selection_follows_ffmpeg_ss_to_fps():
    # -ss 2 -i in -to 3: the output clock restarts at 0, so source times [2, 5] are kept (not [2, 3])
    sel = video._select_frames(250, 25.0, 25.0, 2.0, 3.0)
    assert sel[0] == 50 and sel[-1] == 125 and len(sel) == 76
    assert video._select_frames(250, 25.0, 5.0, None, 1.0) == [2, 7, 12, 17, 22, 25][:6] or len(video._select_frames(250, 25.0, 5.0, None, 1.0)) == 6
    assert video._select_frames(10, 25.0, 5.0, 5.0, None) == []       # window past the end of the file
    assert video._bucket_key(executor.parse_job_argv(_mixed_jobs()[0][-1][0]))[3] == 5.0


def test_stream_pixel_format_tags():
    tag = lambda b: int.from_bytes(b, "little")           # noqa: E731
    assert pc.bit_depth_from_pixel_format_tag(tag(b"I420")) == 8
    assert pc.bit_depth_from_pixel_format_tag(tag(b"J420")) == 8
    assert pc.bit_depth_from_pixel_format_tag(tag(b"Y3\x0b\x0a")) == 10      # yuv420p10le
    assert pc.bit_depth_from_pixel_format_tag(tag(b"Y3\x0a\x0c")) == 10      # yuv422p12le counts as "more than 8"
    assert pc.bit_depth_from_pixel_format_tag(tag(b"P010")) == 10
    assert pc.bit_depth_from_pixel_format_tag(tag(b"G3\x00\x0a")) == 10      # gbrp10le
    assert pc.bit_depth_from_pixel_format_tag(0) == 8 and pc.bit_depth_from_pixel_format_tag(-1) == 8


def test_sixteen_bit_views_are_scaled_not_saturated_for_jpeg(tmp_path):
    cv2 = pytest.importorskip("cv2")
    ramp = np.linspace(0, 65535, 256 * 64).astype(np.uint16).reshape(64, 256)
    img = np.stack([ramp, ramp[::-1], ramp], axis=-1)
    narrow = executor.narrow_to_8bit(img)
    assert narrow.dtype == np.uint8 and narrow.min() == 0 and narrow.max() == 255
    assert np.abs(narrow.astype(np.float64) - img / 257.0).max() <= 0.5 + 1e-9
    assert executor.narrow_to_8bit(narrow) is narrow
    executor._write_image(tmp_path / "v.jpg", img, 100)
    back = cv2.imread(str(tmp_path / "v.jpg"), cv2.IMREAD_UNCHANGED)
    assert back.dtype == np.uint8 and abs(float(back.mean()) - 127.5) < 2.0      # cv2.imwrite alone gives ~254.5
    executor._write_image(tmp_path / "v.png", img, 100)
    assert cv2.imread(str(tmp_path / "v.png"), cv2.IMREAD_UNCHANGED).dtype == np.uint16


def test_decoder_matrix_correction_is_the_601_to_709_rematrix():
    m = color.decoder_matrix_correction("bt601", "bt709")
    assert np.allclose(m.sum(axis=1), 1.0)                 # greys stay put
    # a Y'CbCr sample decoded with either matrix: the correction maps one R'G'B' onto the other
    ycc = np.array([0.4, 0.1, -0.2])
    rgb601, rgb709 = color.ycbcr_to_rgb_matrix("bt601") @ ycc, color.ycbcr_to_rgb_matrix("bt709") @ ycc
    assert np.allclose(m @ rgb601, rgb709)
    assert np.allclose(color.decoder_matrix_correction("bt709", "bt709"), np.eye(3))
    assert abs(color.ycbcr_to_rgb_matrix("bt709")[0, 2] - 1.5748) < 1e-12 and abs(color.ycbcr_to_rgb_matrix("bt601")[2, 1] - 1.772) < 1e-12
