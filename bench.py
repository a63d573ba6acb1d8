#!/usr/bin/env python3
"""Throughput of the panorama -> perspective remap hot path on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
                    [--interp cubic|linear] [--frames B] [--preset full360coverage]

One step = one pass of the hot path over one batch of B synthetic 8K ERP frames: every frame is
cut into the preset's views (full360coverage: 12 x 1600^2, 14 mm).  BASELINE.json's metric is
output Mpix/s (and views/s) per GPU plus the fraction of the HBM roofline, next to the
reference's CPU remap (cv2.remap, cli_tools/gs360_DualFisheyeDistortionCalibration.py:2001-2014)
timed on this box's host cores.

  value      kernel throughput with the frames already resident in HBM (CUDA events)
  e2e        the same frames pushed through the public streaming API from pinned host memory,
             H2D of every frame and D2H of every view inside the timed region
  roofline   algorithmic bytes of the dominant kernel / its measured duration vs MEASURED_PEAKS
  cpu_baseline  cv2.remap + prebuilt float32 maps on the host cores, bounded sample (rank 0, N=1)

`--impl reference` times only that CPU arm and prints the same line with "impl": "reference".
Multi-GPU (torchrun): frames are sharded across ranks, no data-path collective; weak scaling.
"""

import argparse
import json
import os
import pathlib
import statistics
import subprocess
import sys
import threading
import time

ROOT = pathlib.Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT / "360cam-pgm-3dgs-tools_b200"))
sys.path.insert(0, str(ROOT))

ERP_W, ERP_H, CHANNELS = 7680, 3840, 3

# Distinct source pixels touched by one frame's view set (U) -- computed by `python -m
# oracle.footprint` with the float64 oracle maps; O = views * size^2.  Algorithmic bytes per
# frame = (U + O) * channels * sizeof(u8)  (DESIGN.md section 5).
FOOTPRINT_PX = {
    ("full360coverage", "linear"): 24_335_997,
    ("full360coverage", "cubic"): 26_463_848,
    ("fisheyelike", "linear"): 19_305_992,
    ("fisheyelike", "cubic"): 19_998_440,
    ("default", "linear"): 18_073_800,
    ("default", "cubic"): 18_250_224,
}


# DRAM traffic of ONE launch of the tiled kernel at the bench configuration (16 frames x 12 views),
# dram__bytes_read.sum + dram__bytes_write.sum from the `ncu --set full` captures summarised in
# profiles/ (r01_tiled_*_b16.json).  Only valid for the default workload.
NCU_TRAFFIC_BYTES = {
    ("full360coverage", "cubic", 16): 4291430000,     # profiles/r01_tiled_cubic_b16.json
    ("full360coverage", "linear", 16): 5259589000,    # profiles/r01_tiled_linear_b16.json
}


def preset_views(preset, size):
    """(views, hfov) of a gs360_360PerspCut preset through the drop-in planner."""
    from remap360 import perspcut
    ap = perspcut.create_arg_parser()
    args = ap.parse_args(["-i", "/nonexistent", "--preset", preset, "--size", str(size)])
    args.size_explicit = False if size == 1600 else True
    args.input_is_video = False
    res = perspcut.build_view_jobs(args, [pathlib.Path("/nonexistent/frame.png")], pathlib.Path("/nonexistent/out"))
    return [(v.view_id, v.yaw_deg, v.pitch_deg, v.hfov_deg, v.vfov_deg) for v in res.view_specs
            if v.projection == "perspective"]


class ClockSampler:
    """SM clock and throttle reasons of one GPU, sampled every 10 ms on a host thread while the timed region
    runs (B200_PROFILING.md's clocks line).  NVML directly (the library behind nvidia-smi; no process start-up
    inside a region that lasts ~100 ms), `nvidia-smi -lms` as the fallback.  Only samples whose host
    timestamp falls inside [mark_start, mark_end] are summarised."""
    QUERY = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
             "clocks_event_reasons.sw_power_cap")
    REASON_BITS = (("hw_slowdown", 0x8), ("hw_thermal_slowdown", 0x40), ("sw_thermal_slowdown", 0x20),
                   ("sw_power_cap", 0x4))

    def __init__(self, index):
        self.index, self.rows, self.proc, self.thread = index, [], None, None
        self.t0 = self.t1 = None
        self._stop = threading.Event()
        self.source = None

    def _physical_index(self):
        vis = os.environ.get("CUDA_VISIBLE_DEVICES")
        if vis:
            ids = [t.strip() for t in vis.split(",") if t.strip()]
            if self.index < len(ids) and ids[self.index].isdigit():
                return int(ids[self.index])
        return self.index

    def __enter__(self):
        try:
            import pynvml
            pynvml.nvmlInit()
            handle = pynvml.nvmlDeviceGetHandleByIndex(self._physical_index())
            max_mhz = float(pynvml.nvmlDeviceGetMaxClockInfo(handle, pynvml.NVML_CLOCK_SM))
            reasons_fn = getattr(pynvml, "nvmlDeviceGetCurrentClocksEventReasons", None) or \
                pynvml.nvmlDeviceGetCurrentClocksThrottleReasons

            def poll():
                while not self._stop.is_set():
                    try:
                        mhz = float(pynvml.nvmlDeviceGetClockInfo(handle, pynvml.NVML_CLOCK_SM))
                        bits = int(reasons_fn(handle))
                        self.rows.append((time.monotonic(), mhz, max_mhz,
                                          [name for name, bit in self.REASON_BITS if bits & bit]))
                    except Exception:
                        pass
                    time.sleep(0.01)
            self.source = "nvml"
            self.thread = threading.Thread(target=poll, daemon=True)
            self.thread.start()
            return self
        except Exception:
            pass
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(self._physical_index()), "--query-gpu=" + self.QUERY,
                 "--format=csv,noheader,nounits", "-lms", "20"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.source = "nvidia-smi"
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None
        return self

    def _read(self):
        for line in self.proc.stdout:
            r = [c.strip() for c in line.split(",")]
            if len(r) < 7:
                continue
            try:
                self.rows.append((time.monotonic(), float(r[0]), float(r[1]),
                                  [name for name, flag in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown",
                                                               "sw_power_cap"), r[3:7]) if flag.lower().startswith("active")]))
            except ValueError:
                continue

    def wait_ready(self, timeout=5.0):
        """Block until the first sample has arrived (nvidia-smi needs a moment to start)."""
        end = time.monotonic() + timeout
        while not self.rows and time.monotonic() < end:
            time.sleep(0.005)

    def mark_start(self):
        self.t0 = time.monotonic()

    def mark_end(self):
        self.t1 = time.monotonic()

    def __exit__(self, *exc):
        self._stop.set()
        if self.proc:
            self.proc.terminate()
            try:
                self.proc.wait(timeout=5)
            except subprocess.TimeoutExpired:
                self.proc.kill()
        if self.thread:
            self.thread.join(timeout=2)

    def summary(self):
        rows = [r for r in self.rows if (self.t0 is None or r[0] >= self.t0) and (self.t1 is None or r[0] <= self.t1)]
        if not rows:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0, "source": self.source}
        reasons = sorted({name for r in rows for name in r[3]})
        return {"sm_mhz": statistics.median(r[1] for r in rows), "sm_max_mhz": max(r[2] for r in rows),
                "reasons": reasons, "samples": len(rows), "source": self.source}


# --------------------------------------------------------------------------------------------
# CPU arm: the reference's remap (cv2.remap on prebuilt float32 maps, DF:2001-2014)
# --------------------------------------------------------------------------------------------

def cpu_reference_arm(views, size, interp, seconds_budget, steps=None, warmup=1):
    """Returns (Mpix/s, info).  One CPU 'step' = ONE frame cut into all views (a bounded sample of
    the GPU step, which is B such frames); maps are built once outside the timed region exactly
    as the reference does (DF:1857-1907 builds maps once, DF:1996-2014 applies them per frame)."""
    import cv2
    import numpy as np
    from oracle import geometry as geo
    cores = os.cpu_count() or 1
    cv2.setNumThreads(cores)
    rng = np.random.default_rng(1234)
    frame = rng.integers(0, 256, (ERP_H, ERP_W, CHANNELS), dtype=np.uint8)
    pad = 4
    padded = np.concatenate([frame[:, -pad:], frame, frame[:, :pad]], axis=1)
    padded = np.ascontiguousarray(np.concatenate([padded[:1].repeat(pad, 0), padded, padded[-1:].repeat(pad, 0)], 0))
    t0 = time.perf_counter()
    maps = []
    for _, yaw, pitch, hfov, vfov in views:
        mx, my = geo.erp_map64(ERP_W, ERP_H, size, size, yaw, pitch, hfov, vfov)
        maps.append(((mx + pad).astype(np.float32), (my + pad).astype(np.float32)))
    map_build_s = time.perf_counter() - t0
    flag = {"linear": cv2.INTER_LINEAR, "cubic": cv2.INTER_CUBIC, "nearest": cv2.INTER_NEAREST}[interp]

    def one_frame():
        for mx, my in maps:
            cv2.remap(padded, mx, my, flag, borderMode=cv2.BORDER_CONSTANT, borderValue=0)

    for _ in range(max(1, warmup)):
        one_frame()
    times = []
    t_start = time.perf_counter()
    while True:
        t = time.perf_counter()
        one_frame()
        times.append(time.perf_counter() - t)
        if steps is not None and len(times) >= steps:
            break
        if steps is None and (time.perf_counter() - t_start) > seconds_budget and len(times) >= 3:
            break
    pix = len(views) * size * size
    mean_t = sum(times) / len(times)
    info = {"cores": cv2.getNumThreads(), "host_cpus": cores, "frames_timed": len(times),
            "s_per_frame": mean_t, "map_build_s_once": map_build_s, "cv2": cv2.__version__,
            "cv2_threads": cv2.getNumThreads()}
    return pix / mean_t / 1e6, info


# --------------------------------------------------------------------------------------------

def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", choices=["ours", "reference"], default="ours")
    ap.add_argument("--interp", choices=["cubic", "linear"], default="cubic",
                    help="cubic is the reference's hard-wired default (PC:730, DF:232)")
    ap.add_argument("--preset", default="full360coverage")
    ap.add_argument("--frames", type=int, default=16, help="8K frames per GPU per step")
    ap.add_argument("--size", type=int, default=1600)
    ap.add_argument("--path", default="auto")
    ap.add_argument("--cpu-seconds", type=float, default=15.0)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-variants", action="store_true", help="skip timing the other interpolation")
    ns = ap.parse_args()

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if ns.warmup < 3 and ns.impl == "ours":
        ns.warmup = 3
    views = preset_views(ns.preset, ns.size)
    n_views = len(views)
    workload = "%dx%d u8x3 ERP x%d frames/GPU -> %s preset %d views %dx%d, %s" % (
        ERP_W, ERP_H, ns.frames, ns.preset, n_views, ns.size, ns.size, ns.interp)
    config = {"workload": workload, "preset": ns.preset, "views": n_views, "out_size": ns.size,
              "interp": ns.interp, "frames_per_gpu_per_step": ns.frames, "convention": "halfpixel",
              "content": "uniform noise (seeded)", "l2_policy": "inputs (1.4 GB/step) larger than L2"}

    # ------------------------------------------------------------------ reference arm
    if ns.impl == "reference":
        if rank != 0:
            return
        value, info = cpu_reference_arm(views, ns.size, ns.interp, ns.cpu_seconds, steps=ns.steps, warmup=ns.warmup)
        sample = "1 frame (all %d views) per step, cv2.remap %s on prebuilt float32 maps, %d cv2 threads" % (
            n_views, ns.interp, info["cv2_threads"])
        line = {"impl": "reference", "metric": "output Mpix/s", "value": value, "unit": "Mpix/s",
                "n_gpus": ns.gpus, "steps": ns.steps, "warmup": ns.warmup, "ms_per_step": info["s_per_frame"] * 1e3,
                "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u8",
                "data": "synthetic", "config": config,
                "views_per_s": value * 1e6 / (ns.size * ns.size),
                "cpu_baseline": {"value": value, "unit": "Mpix/s", "cores": info["cores"], "kind": "reference",
                                 "sample": sample, "detail": info},
                "e2e": {"value": value, "unit": "Mpix/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
                "gpu_launches": 0}
        print(json.dumps(line))
        return

    # ------------------------------------------------------------------ our arm
    import torch
    import remap360
    from remap360.stream import StreamingRemapper

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (no CPU fallback)")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    dist = None
    saved_stdout_fd = None
    if world > 1:
        import torch.distributed as dist
        # NCCL writes its version banner to file descriptor 1 when the communicator is created: keep stdout for
        # the one JSON line by pointing fd 1 at stderr until that line is printed
        sys.stdout.flush()
        saved_stdout_fd = os.dup(1)
        os.dup2(2, 1)
        dist.init_process_group("nccl", device_id=dev)

    pviews = [remap360.PerspectiveView(y, p, hf, vf, view_id=vid) for vid, y, p, hf, vf in views]
    # frames are sharded by index: rank r owns global frames [r*B, (r+1)*B); seeded per frame
    frames = torch.empty((ns.frames, ERP_H, ERP_W, CHANNELS), dtype=torch.uint8, device=dev)
    for f in range(ns.frames):
        g = torch.Generator(device=dev)
        g.manual_seed(1234 + rank * ns.frames + f)
        frames[f] = torch.randint(0, 256, (ERP_H, ERP_W, CHANNELS), dtype=torch.uint8, device=dev, generator=g)
    out = torch.empty((ns.frames, n_views, ns.size, ns.size, CHANNELS), dtype=torch.uint8, device=dev)

    def step(interp=ns.interp):
        remap360.remap_erp(frames, pviews, (ns.size, ns.size), interp=interp, out=out, path=ns.path)

    def barrier():
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(interp, steps, warmup, sample_clocks):
        """K steps bracketed by barrier + synchronize, CUDA events on the launch stream, max over ranks."""
        sampler = ClockSampler(local_rank) if sample_clocks else None
        if sampler:
            sampler.__enter__()
        for _ in range(warmup):
            step(interp)
        if sampler:
            sampler.wait_ready()
        barrier()
        l0 = remap360.launch_count()
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(steps + 1)]
        if sampler:
            sampler.mark_start()
        ev[0].record()
        for k in range(steps):
            step(interp)
            ev[k + 1].record()
        barrier()
        if sampler:
            sampler.mark_end()
            sampler.__exit__(None, None, None)
        per_step = [ev[k].elapsed_time(ev[k + 1]) for k in range(steps)]
        total = ev[0].elapsed_time(ev[steps])
        if dist is not None:
            t = torch.tensor([total], dtype=torch.float64, device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            total = float(t.item())
        return total / steps, per_step, remap360.launch_count() - l0, sampler

    ms_per_step, step_ms, launches, clocks = timed(ns.interp, ns.steps, ns.warmup, True)
    out_pix_step = ns.frames * n_views * ns.size * ns.size
    value = out_pix_step * world / (ms_per_step * 1e-3) / 1e6          # Mpix/s, whole job

    # roofline of the dominant kernel (remap_tiled_kernel: one launch per step covers the whole batch;
    # the small fallback-tile launch that follows it is part of the step time used here)
    peaks_file = ROOT / "MEASURED_PEAKS.json"
    if peaks_file.exists():
        peak, peak_src = float(json.loads(peaks_file.read_text())["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    else:
        peak, peak_src = 6650.0, "fallback (B200_PROFILING.md)"

    def roofline_for(interp, per_step_ms):
        u_px = FOOTPRINT_PX.get((ns.preset, interp))
        if u_px is None or ns.size != 1600:
            return None
        bytes_per_launch = (u_px + n_views * ns.size * ns.size) * CHANNELS * ns.frames
        kernel_ms = statistics.median(per_step_ms)
        achieved = bytes_per_launch / (kernel_ms * 1e-3) / 1e9
        return {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                "traffic": NCU_TRAFFIC_BYTES.get((ns.preset, interp, ns.frames)), "peak_source": peak_src,
                "algorithmic_bytes_per_launch": bytes_per_launch, "kernel_ms": kernel_ms,
                "kernel": "remap_tiled_kernel<%s, u8, u8> (+ remap_fallback_kernel for %s)" % (
                    interp, "tiles the plan routes to the direct path")}

    roofline = roofline_for(ns.interp, step_ms)
    # the other interpolation on the same workload (kernel-only), for context
    variants = {}
    if not ns.no_variants:
        other = "linear" if ns.interp == "cubic" else "cubic"
        o_ms, o_steps, _, _ = timed(other, max(3, ns.steps // 2), 3, False)
        variants[other] = {"value": out_pix_step * world / (o_ms * 1e-3) / 1e6, "unit": "Mpix/s", "ms_per_step": o_ms,
                           "roofline": roofline_for(other, o_steps)}

        # content class (S) of SURVEY.md section 8(d): smooth band-limited frames (eight low-frequency sinusoids
        # in lon / lat per channel, continuous across the seam) instead of noise; the kernels have no
        # data-dependent control flow, so the number should not move
        keep = frames[0].clone()
        lon = torch.linspace(0.0, 2.0 * torch.pi, ERP_W + 1, device=dev)[:-1]
        lat = torch.linspace(-0.5 * torch.pi, 0.5 * torch.pi, ERP_H, device=dev)
        gen = torch.Generator(device=dev)
        gen.manual_seed(4321 + rank)
        smooth = torch.zeros((ERP_H, ERP_W, CHANNELS), dtype=torch.float32, device=dev)
        for _ in range(8):
            kx = torch.randint(1, 9, (CHANNELS,), device=dev, generator=gen).float()
            ky = torch.randint(1, 9, (CHANNELS,), device=dev, generator=gen).float()
            ph = torch.rand((2, CHANNELS), device=dev, generator=gen) * 2.0 * torch.pi
            smooth += torch.sin(lon[None, :, None] * kx + ph[0]) * torch.cos(lat[:, None, None] * ky + ph[1])
        smooth = ((smooth / 16.0 + 0.5).clamp_(0.0, 1.0) * 255.0).round_().to(torch.uint8)
        for f in range(ns.frames):
            frames[f] = smooth
        s_ms, _, _, _ = timed(ns.interp, max(3, ns.steps // 4), 3, False)
        variants["content_smooth"] = {"value": out_pix_step * world / (s_ms * 1e-3) / 1e6, "unit": "Mpix/s",
                                      "ms_per_step": s_ms, "interp": ns.interp,
                                      "content": "sum of 8 low-frequency sinusoids per channel, seam-continuous"}
        for f in range(ns.frames):              # back to the seeded noise frames for the end-to-end leg
            g = torch.Generator(device=dev)
            g.manual_seed(1234 + rank * ns.frames + f)
            frames[f] = torch.randint(0, 256, (ERP_H, ERP_W, CHANNELS), dtype=torch.uint8, device=dev, generator=g)
        assert torch.equal(frames[0], keep)
        del keep, smooth

    # ------------------------------------------------------------------ end to end (host buffers)
    e2e = None
    if not ns.no_e2e:
        remapper = StreamingRemapper(pviews, (ns.size, ns.size), (ERP_H, ERP_W, CHANNELS), torch.uint8,
                                     interp=ns.interp, device=dev, path=ns.path)
        host_frames = [torch.empty((ERP_H, ERP_W, CHANNELS), dtype=torch.uint8).pin_memory() for _ in range(min(ns.frames, 4))]
        for k, hf in enumerate(host_frames):
            hf.copy_(frames[k].cpu())
        e2e_steps = max(2, min(ns.steps, 5))

        def e2e_step():
            sink = 0
            for res in remapper.run(host_frames[f % len(host_frames)] for f in range(ns.frames)):
                sink += int(res[0, 0, 0, 0])      # touch the host copy of every result
            return sink

        e2e_step()
        barrier()
        t0 = time.perf_counter()
        for _ in range(e2e_steps):
            e2e_step()
        barrier()
        e2e_s = (time.perf_counter() - t0) / e2e_steps
        if dist is not None:
            t = torch.tensor([e2e_s], dtype=torch.float64, device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            e2e_s = float(t.item())
        e2e = {"value": out_pix_step * world / e2e_s / 1e6, "unit": "Mpix/s",
               "h2d_bytes_per_step": ns.frames * ERP_H * ERP_W * CHANNELS,
               "d2h_bytes_per_step": out_pix_step * CHANNELS, "ms_per_step": e2e_s * 1e3, "steps": e2e_steps,
               "api": "remap360.stream.StreamingRemapper (pinned ring, H2D / kernel / D2H on three streams)"}

    cpu_baseline = None
    if rank == 0 and world == 1 and not ns.no_cpu_baseline:
        v, info = cpu_reference_arm(views, ns.size, ns.interp, ns.cpu_seconds)
        cpu_baseline = {"value": v, "unit": "Mpix/s", "cores": info["cores"], "kind": "reference",
                        "sample": "%d frames x %d views, cv2.remap %s (%s) on prebuilt float32 maps, %d threads of %d host CPUs"
                                  % (info["frames_timed"], n_views, ns.interp, info["cv2"], info["cv2_threads"], info["host_cpus"]),
                        "detail": info}

    if rank == 0:
        line = {"metric": "output Mpix/s", "value": value, "unit": "Mpix/s", "n_gpus": world, "steps": ns.steps,
                "warmup": ns.warmup, "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak",
                "vs_baseline": None, "dtype": "u8", "data": "synthetic", "config": config,
                "views_per_s": value * 1e6 / (ns.size * ns.size), "frames_per_s": value * 1e6 / (n_views * ns.size * ns.size),
                "clocks": clocks.summary(), "e2e": e2e, "gpu_launches": int(launches),
                "roofline": roofline, "cpu_baseline": cpu_baseline, "variants": variants}
        if saved_stdout_fd is not None:
            sys.stdout.flush()
            os.dup2(saved_stdout_fd, 1)
        print(json.dumps(line), flush=True)
    if dist is not None:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
