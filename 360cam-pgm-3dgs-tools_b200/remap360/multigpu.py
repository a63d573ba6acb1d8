"""Planner jobs on every GPU of the box: one worker process per device, no data-path collective.

The reference parallelises inside the tool -- a thread pool over (file, view) jobs in the cutter
(gs360_360PerspCut.py:1049-1051) and over pairs in the dual-fisheye tool
(gs360_DualFisheyeDistortionCalibration.py:2761-2810).  Here the unit that moves between workers is what the GPU
path shares: a SOURCE.  All views of a still image go to one device (one decode, one upload), still sources are
dealt out as contiguous ranges (``sharding.shard_range``), and a video source is split by contiguous ranges of its
OUTPUT frame numbers -- every rank opens the file, seeks to its range and writes ``<stem>_%07d_<view>`` with the
global numbers (gs360_360PerspCut.py:746-749), so the union of the ranks' files is the single-GPU result.

``plan_shards`` is the pure part (no CUDA, no files): it is what the CPU tests (2-rank gloo) exercise.
``run_jobs`` is the drop-in for ``executor.run_jobs``: with one visible device it IS ``executor.run_jobs``;
with several it spawns one process per device (``torch.multiprocessing``, spawn) and merges the per-job results."""

from __future__ import annotations

import os
import threading
from collections import OrderedDict
from typing import Dict, Iterator, List, Optional, Sequence, Tuple

from .sharding import shard_range


def group_jobs(jobs) -> "OrderedDict[str, List[int]]":
    """Job indices grouped by source in first-appearance order (the executor's own grouping); unparsable jobs get a
    group of their own so that they are reported by exactly one rank."""
    from . import executor
    groups: "OrderedDict[str, List[int]]" = OrderedDict()
    for n, (cmd, _src, _dst) in enumerate(jobs):
        try:
            pj = executor.parse_job_argv(cmd)
            key = str(pj.source) + ("|video" if pj.video else "")
        except Exception:
            key = "|unparsable|%d" % n
        groups.setdefault(key, []).append(n)
    return groups


def plan_shards(jobs, world: int) -> List[Dict[str, object]]:
    """Per rank: ``{"stills": [job indices], "videos": [(job indices, (rank, world))]}``.

    Still sources are split into `world` contiguous ranges of sources; every video source is given to every rank
    with its frame shard.  The plan depends on the job list only, so every rank (or a parent process) computes the
    same one."""
    if world < 1:
        raise ValueError("world must be >= 1")
    groups = group_jobs(jobs)
    stills = [idxs for key, idxs in groups.items() if not key.endswith("|video")]
    videos = [idxs for key, idxs in groups.items() if key.endswith("|video")]
    plan = []
    for rank in range(world):
        lo, hi = shard_range(len(stills), world, rank)
        plan.append({"stills": [n for idxs in stills[lo:hi] for n in idxs],
                     "videos": [(list(idxs), (rank, world)) for idxs in videos]})
    return plan


def expected_outputs(jobs, shard_plan: Dict[str, object], frame_counts: Optional[Dict[str, Tuple[int, float]]] = None) -> List[str]:
    """Output paths one rank's share of `jobs` produces (test / dry-run helper).  Video sources need
    ``frame_counts[source] = (frames, fps)``."""
    from . import executor, video
    out: List[str] = []
    for n in shard_plan["stills"]:
        try:
            out.append(str(executor.parse_job_argv(jobs[n][0]).output))
        except Exception:
            pass
    for idxs, (rank, world) in shard_plan["videos"]:
        for n in idxs:
            pj = executor.parse_job_argv(jobs[n][0])
            n_in, in_fps = (frame_counts or {})[str(pj.source)]
            wanted = video._select_frames(n_in, in_fps, float(pj.fps), pj.start, pj.end)
            lo, hi = shard_range(len(wanted), world, rank)
            out.extend(str(pj.output) % k for k in range(lo, hi))
    return out


def merge_results(per_rank: Sequence[Dict[int, Tuple[int, str]]], n_jobs: int) -> List[Tuple[int, str]]:
    """Job results of all ranks -> one (rc, err) per job: a job fails if it failed on any rank that ran it (video
    jobs run on every rank); 130 (cancelled) outranks success."""
    merged: List[Optional[Tuple[int, str]]] = [None] * n_jobs
    for results in per_rank:
        for n, (rc, err) in results.items():
            cur = merged[n]
            if cur is None or (cur[0] == 0 and rc != 0) or (cur[0] == 130 and rc not in (0, 130)):
                merged[n] = (rc, err)
    return [m if m is not None else (1, "job was not assigned to any device") for m in merged]


def run_shard(jobs, shard_plan: Dict[str, object], stop_event=None, workers: int = 1, device=None) -> Dict[int, Tuple[int, str]]:
    """One rank's share on the current (or given) device."""
    import torch
    from . import executor, video
    results: Dict[int, Tuple[int, str]] = {}
    if device is not None:
        torch.cuda.set_device(device)
    still_jobs = [jobs[n] for n in shard_plan["stills"]]
    for (job, res), n in zip(_in_order(executor.run_jobs(still_jobs, stop_event, workers), still_jobs), shard_plan["stills"]):
        results[n] = res
    for idxs, shard in shard_plan["videos"]:
        parsed = [executor.parse_job_argv(jobs[n][0]) for n in idxs]
        for n, res in zip(idxs, video.run_video_jobs(parsed[0].source, parsed, stop_event, shard=shard)):
            results[n] = res
    return results


def _in_order(pairs, jobs_list):
    """executor.run_jobs yields (job, result) as groups finish; put the results back into list order."""
    got = {id(job): res for job, res in pairs}
    return [(job, got.get(id(job), (1, "job skipped"))) for job in jobs_list]


def _worker(rank: int, world: int, jobs, workers: int, out_queue, stop_flag) -> None:
    try:
        import torch
        device = rank % max(1, torch.cuda.device_count())      # more ranks than devices only in tests (oversubscribed)
        torch.cuda.set_device(device)
        from . import sharding
        sharding.bind_to_gpu_cpus(device)
        stop = threading.Event()

        def watch():
            stop_flag.wait()
            stop.set()
        threading.Thread(target=watch, daemon=True).start()
        res = run_shard(jobs, plan_shards(jobs, world)[rank], stop, workers)
        out_queue.put((rank, res, None))
    except BaseException as exc:                       # the parent must hear from every rank
        import traceback
        out_queue.put((rank, {}, "%s: %s\n%s" % (type(exc).__name__, exc, traceback.format_exc())))


def visible_devices() -> int:
    forced = os.environ.get("R360_DEVICES")
    try:
        import torch
        n = torch.cuda.device_count()
    except Exception:
        n = 0
    return max(0, min(n, int(forced))) if forced else n


def run_jobs(jobs, stop_event=None, workers: int = 1, devices: Optional[int] = None) -> Iterator[Tuple[tuple, Tuple[int, str]]]:
    """Drop-in for ``executor.run_jobs`` that uses every visible GPU.  Yields (job, (rc, err)) for every job."""
    from . import executor
    jobs = list(jobs)
    world = visible_devices() if devices is None else int(devices)
    n_sources = len(group_jobs(jobs))
    has_video = any(key.endswith("|video") for key in group_jobs(jobs))
    if world <= 1 or (n_sources < 2 and not has_video):
        yield from executor.run_jobs(jobs, stop_event, workers)
        return
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    out_queue, stop_flag = ctx.Queue(), ctx.Event()
    procs = [ctx.Process(target=_worker, args=(rank, world, jobs, workers, out_queue, stop_flag), daemon=True)
             for rank in range(world)]
    for p in procs:
        p.start()
    per_rank, failures = [], []
    pending = world
    while pending:
        try:
            rank, res, err = out_queue.get(timeout=0.2)
        except Exception:                               # queue.Empty: poll the caller's stop request and dead workers
            if stop_event is not None and stop_event.is_set():
                stop_flag.set()
            dead = [r for r, p in enumerate(procs) if not p.is_alive() and p.exitcode not in (0, None)]
            if dead and out_queue.empty():
                for r in dead:
                    failures.append("worker of device %d exited with code %s" % (r, procs[r].exitcode))
                pending -= len(dead)
                procs = [p if k not in dead else _Done() for k, p in enumerate(procs)]
            continue
        pending -= 1
        per_rank.append(res)
        if err:
            failures.append("device %d: %s" % (rank, err))
    for p in procs:
        p.join(timeout=5)
    merged = merge_results(per_rank, len(jobs))
    if failures:
        merged = [(rc, err) if rc != 1 or err != "job was not assigned to any device" else (1, "; ".join(failures)) for rc, err in merged]
    for job, res in zip(jobs, merged):
        yield job, res


class _Done:
    exitcode = 0

    def is_alive(self):
        return False

    def join(self, timeout=None):
        return None
