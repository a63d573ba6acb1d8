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
HARD_VIEWS = [(180.0, 0.0), (179.9, 0.0), (-179.9, 0.0), (0.0, 90.0), (0.0, -90.0), (40.0, 60.0), (-40.0, -60.0)]

# Distinct source pixels touched by one frame's (pair's) view set (U) -- computed by `python -m oracle.footprint` with
# the float64 oracle maps; O = views * size^2.  Algorithmic bytes per frame = (U + O) * channels * sizeof(sample)
# (DESIGN.md section 5).  Keys: (view set, source width, interpolation).
FOOTPRINT_PX = {
    ("full360coverage", 7680, "linear"): 24_335_997, ("full360coverage", 7680, "cubic"): 26_463_848,
    ("full360coverage+hard", 7680, "linear"): 26_447_599, ("full360coverage+hard", 7680, "cubic"): 28_174_810,
    ("fisheyelike", 7680, "linear"): 19_305_992, ("fisheyelike", 7680, "cubic"): 19_998_440,
    ("default", 7680, "linear"): 18_073_800, ("default", 7680, "cubic"): 18_250_224,
    ("default", 3840, "linear"): 4_562_272, ("default", 3840, "cubic"): 4_570_656,
    ("dualfisheye_sfm10", 3840, "linear"): 20_721_452, ("dualfisheye_sfm10", 3840, "cubic"): 20_750_796,
}

# DRAM traffic of ONE launch of the tiled kernel (dram__bytes_read.sum + dram__bytes_write.sum, ncu, per workload of
# workload_table at the default batch), valid only for the kernel sources it was measured with:
# tools/traffic_capture.py writes profiles/r02_dram_traffic.json with the digest of csrc/ at capture time, and an
# entry is reported as null (with the reason) when the sources have changed since.
TRAFFIC_FILE = ROOT / "profiles" / "r02_dram_traffic.json"


def ncu_traffic():
    """{(workload, interp, frames): (bytes, csrc digest, profile)} from TRAFFIC_FILE; {} when it is absent."""
    try:
        rec = json.loads(TRAFFIC_FILE.read_text())
    except (OSError, ValueError):
        return {}
    return {(name, e["interp"], int(e["frames"])): (int(e["dram_bytes_read"]) + int(e["dram_bytes_write"]), rec["csrc_digest"],
                                                     "profiles/" + TRAFFIC_FILE.name)
            for name, e in rec.get("workloads", {}).items() if "error" not in e}


def csrc_digest():
    import hashlib
    h = hashlib.sha256()
    for p in sorted((ROOT / "360cam-pgm-3dgs-tools_b200" / "csrc").glob("*")):
        h.update(p.name.encode())
        h.update(p.read_bytes())
    return h.hexdigest()[:16]


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
# workloads: BASELINE.json configs 1..5 (configs[1] is the headline; the others ride along as `variants`)
# --------------------------------------------------------------------------------------------

def dualfisheye_setup():
    """Config 5: the template calibration, the SFM10 ten-view layout at 1750 px and the recorded lens choice
    (tests/golden/dualfisheye.json, produced by importing the reference: tests/golden/make_golden.py)."""
    meta = json.loads((ROOT / "tests" / "golden" / "dualfisheye.json").read_text())
    cal = meta["sensors"]["0"]
    lens_of = {vid: (0 if i["lens_key"] == "X" else 1) for vid, i in meta["maps_1750"]["views"].items()}
    views = []
    for sp in meta["sfm10_default"]:
        slot = lens_of[sp["view_id"]]
        yaw_rel = ((sp["yaw_deg"] - (0.0, 180.0)[slot] + 180.0) % 360.0) - 180.0
        views.append((sp["view_id"], yaw_rel, sp["pitch_deg"], sp["hfov_deg"], sp["vfov_deg"], slot))
    return cal, views, meta


def workload_table(ns):
    """name -> dict(kind, src (W, H), dtype, out_dtype, viewset, hard, size, interp, frames)."""
    F = ns.frames
    return {
        "cfg2": dict(cfg=2, kind="erp", src=(7680, 3840), dtype="u8", out="u8", viewset=ns.preset, hard=False, size=ns.size,
                     interp=ns.interp, frames=F),
        "cfg2_other_interp": dict(cfg=2, kind="erp", src=(7680, 3840), dtype="u8", out="u8", viewset=ns.preset, hard=False,
                                  size=ns.size, interp="linear" if ns.interp == "cubic" else "cubic", frames=F),
        "cfg2_seam_pole_views": dict(cfg=2, kind="erp", src=(7680, 3840), dtype="u8", out="u8", viewset="full360coverage",
                                     hard=True, size=1600, interp="cubic", frames=max(2, F // 2)),
        "cfg1": dict(cfg=1, kind="erp", src=(3840, 1920), dtype="u8", out="u8", viewset="default", hard=False, size=1600,
                     interp="cubic", frames=F),
        "cfg3": dict(cfg=3, kind="erp", src=(7680, 3840), dtype="u8", out="u8", viewset="fisheyelike", hard=False, size=1600,
                     interp="cubic", frames=F),
        "cfg4_u16_to_u16": dict(cfg=4, kind="erp", src=(7680, 3840), dtype="u16", out="u16", viewset="full360coverage",
                                hard=True, size=1600, interp="cubic", frames=max(2, F // 4)),
        "cfg4_u16_to_f16": dict(cfg=4, kind="erp", src=(7680, 3840), dtype="u16", out="f16", viewset="full360coverage",
                                hard=True, size=1600, interp="cubic", frames=max(2, F // 4)),
        "cfg5": dict(cfg=5, kind="dualfisheye", src=(3840, 3840), dtype="u8", out="u8", viewset="dualfisheye_sfm10",
                     hard=False, size=1750, interp="cubic", frames=max(2, F // 2)),
    }


def workload_views(w):
    if w["kind"] == "dualfisheye":
        _, views, _ = dualfisheye_setup()
        return views
    views = [(vid, y, p, hf, vf, 0) for vid, y, p, hf, vf in preset_views(w["viewset"], w["size"])]
    if w["hard"]:
        views += [("x%g_%g" % (y, p), y, p, views[0][3], views[0][4], 0) for y, p in HARD_VIEWS]
    return views


def workload_text(w, n_views):
    return "%dx%d %sx3 %s x%d %s/GPU -> %s%s %d views %dx%d %s, %s" % (
        w["src"][0], w["src"][1], w["dtype"], "fisheye pairs" if w["kind"] == "dualfisheye" else "ERP", w["frames"],
        "pairs" if w["kind"] == "dualfisheye" else "frames", w["viewset"], " + seam/pole views" if w["hard"] else "",
        n_views, w["size"], w["size"], w["out"], w["interp"])


def algorithmic_bytes(w, n_views):
    """(U + O) * C * sizeof per frame / pair; None when U is not tabulated for the view set."""
    key = (w["viewset"] + ("+hard" if w["hard"] else ""), w["src"][0], w["interp"])
    u = FOOTPRINT_PX.get(key)
    expect = {"full360coverage": 12, "fisheyelike": 10, "default": 8, "dualfisheye_sfm10": 10}.get(w["viewset"])
    if u is None or expect is None or n_views != expect + (7 if w["hard"] else 0) or w["size"] not in (1600, 1750):
        return None, None
    es_in, es_out = (2 if w["dtype"] == "u16" else 1), (1 if w["out"] == "u8" else 2)
    return u * CHANNELS * es_in + n_views * w["size"] * w["size"] * CHANNELS * es_out, u


# --------------------------------------------------------------------------------------------
# CPU arm: the reference's remap (cv2.remap on prebuilt float32 maps, DF:2001-2014)
# --------------------------------------------------------------------------------------------

def cpu_model_name():
    try:
        for line in open("/proc/cpuinfo"):
            if line.startswith("model name"):
                return line.split(":", 1)[1].strip()
    except OSError:
        pass
    return "unknown"


def cpu_maps_for(w, views):
    """Float32 maps of the workload, built once outside the timed region as the reference does (DF:1857-1907 builds
    them once, DF:1996-2014 applies them to every pair).  ERP: the oracle's restatement of the geometry (a PORT: the
    reference's ERP path is an ffmpeg process).  Dual fisheye: the reference's own build_perspective_spec_maps when
    /root/reference is importable (kind "reference"), else the oracle's restatement of it."""
    import numpy as np
    from oracle import geometry as geo
    W, H = w["src"]
    size = w["size"]
    if w["kind"] == "dualfisheye":
        cal, _, meta = dualfisheye_setup()
        ref_dir = pathlib.Path("/root/reference/cli_tools")
        if ref_dir.exists():
            try:
                sys.path.insert(0, str(ref_dir))
                import gs360_DualFisheyeDistortionCalibration as DF
                calib = DF.SensorCalibration(sensor_id="0", model_type="equisolid_fisheye", **{k: cal[k] for k in (
                    "f", "cx", "cy", "k1", "k2", "k3", "k4", "p1", "p2", "b1", "b2")}, width=int(cal["width"]), height=int(cal["height"]))
                ref_specs = DF.build_sfm10_specs(size, 14.0, "36 36", 40.0, 40.0)      # the CLI defaults (DF:360-388)
                want = {sp["view_id"]: sp for sp in meta["sfm10_default"]}
                assert all(abs(float(sp["yaw_deg"]) - want[sp["view_id"]]["yaw_deg"]) < 1e-9 for sp in ref_specs)
                maps = DF.build_perspective_spec_maps({"0": calib}, "0", "0", ref_specs, 0.0, 180.0, 190.0)
                out = []
                for sp in ref_specs:
                    m = maps[str(sp["view_id"])]
                    out.append((np.ascontiguousarray(m["map_x"], dtype=np.float32), np.ascontiguousarray(m["map_y"], dtype=np.float32),
                                0 if m["lens_key"] == "X" else 1, np.ascontiguousarray(m["valid"])))
                return out, "reference", 0
            except Exception as exc:                                   # the reference's signature differs: use the port
                sys.stderr.write("bench: reference map builder not usable (%s: %s); oracle maps instead\n" % (type(exc).__name__, exc))
        specs = [dict(sp, width=size, height=size) for sp in meta["sfm10_default"]]
        vm = geo.dualfisheye_view_maps(cal, cal, specs)
        return [(v["map_x"].astype(np.float32), v["map_y"].astype(np.float32), 0 if v["lens_key"] == "X" else 1, v["valid"])
                for v in (vm[sp["view_id"]] for sp in specs)], "port", 0
    pad = 4
    maps = []
    for _, yaw, pitch, hfov, vfov, _slot in views:
        mx, my = geo.erp_map64(W, H, size, size, yaw, pitch, hfov, vfov)
        maps.append(((mx + pad).astype(np.float32), (my + pad).astype(np.float32), 0, None))
    return maps, "port", pad


def cpu_reference_arm(w, views, seconds_budget, steps=None, warmup=1, rows=("all_threads",)):
    """Returns {row: (Mpix/s, info)}.  One CPU 'step' = ONE frame (pair) cut into all views -- a bounded sample of the
    GPU step, which is B such frames.  Rows: cv2 with all host threads (frames one after the other), cv2 with one
    thread, and a pool of host threads over frames with single-threaded cv2 calls (the shape of the reference's own
    pools, PC:1049-1051 / DF:2761-2810)."""
    import cv2
    import numpy as np
    from concurrent.futures import ThreadPoolExecutor
    cores = os.cpu_count() or 1
    rng = np.random.default_rng(1234)
    W, H = w["src"]
    np_dtype = np.uint16 if w["dtype"] == "u16" else np.uint8
    n_src = 2 if w["kind"] == "dualfisheye" else 1
    frames = [rng.integers(0, np.iinfo(np_dtype).max + 1, (H, W, CHANNELS), dtype=np_dtype) for _ in range(n_src)]
    t0 = time.perf_counter()
    maps, kind, pad = cpu_maps_for(w, views)
    map_build_s = time.perf_counter() - t0
    if pad:
        frames = [np.ascontiguousarray(np.concatenate(
            [np.concatenate([f[:, -pad:], f, f[:, :pad]], axis=1)[:1].repeat(pad, 0), np.concatenate([f[:, -pad:], f, f[:, :pad]], axis=1),
             np.concatenate([f[:, -pad:], f, f[:, :pad]], axis=1)[-1:].repeat(pad, 0)], 0)) for f in frames]
    flag = {"linear": cv2.INTER_LINEAR, "cubic": cv2.INTER_CUBIC, "nearest": cv2.INTER_NEAREST}[w["interp"]]

    def one_frame():
        for mx, my, slot, valid in maps:
            out = cv2.remap(frames[slot], mx, my, flag, borderMode=cv2.BORDER_CONSTANT, borderValue=0)
            if valid is not None:
                out[~valid] = 0                                  # rendered[~valid] = mask_value (DF:2009-2014)
            if w["out"] == "f16":
                out = (out.astype(np.float32) * np.float32(1.0 / 65535.0)).astype(np.float16)

    pix = len(maps) * w["size"] * w["size"]
    results = {}
    for row in rows:
        threads = 1 if row in ("one_thread", "frame_pool") else cores
        cv2.setNumThreads(threads)
        pool_n = cores if row == "frame_pool" else 1
        budget = seconds_budget if row == "all_threads" else seconds_budget / 3.0
        for _ in range(max(1, warmup)):
            one_frame()
        times, t_start, frames_done = [], time.perf_counter(), 0
        while True:
            t = time.perf_counter()
            if pool_n > 1:
                with ThreadPoolExecutor(max_workers=pool_n) as ex:
                    list(ex.map(lambda _i: one_frame(), range(pool_n)))
                frames_done += pool_n
            else:
                one_frame()
                frames_done += 1
            times.append(time.perf_counter() - t)
            if steps is not None and row == "all_threads" and len(times) >= steps:
                break
            if (steps is None or row != "all_threads") and (time.perf_counter() - t_start) > budget and len(times) >= (3 if pool_n == 1 else 1):
                break
        total = sum(times)
        info = {"cores": pool_n if pool_n > 1 else threads, "cv2_threads": threads, "pool_workers": pool_n, "host_cpus": cores,
                "frames_timed": frames_done, "s_per_frame": total / frames_done, "map_build_s_once": map_build_s,
                "maps": kind, "cv2": cv2.__version__, "cpu_model": cpu_model_name(),
                "cv2_parallel_framework": next((l.split(":", 1)[1].strip() for l in cv2.getBuildInformation().splitlines()
                                                if "Parallel framework" in l), "unknown")}
        results[row] = (pix * frames_done / total / 1e6, info)
    cv2.setNumThreads(cores)
    return results, kind


def cpu_baseline_block(w, views, seconds, rows=("all_threads", "one_thread", "frame_pool"), steps=None, warmup=1):
    res, kind = cpu_reference_arm(w, views, seconds, steps=steps, warmup=warmup, rows=rows)
    v, info = res["all_threads"]
    what = "cv2.remap %s (%s)%s on float32 maps built once by %s" % (
        w["interp"], info["cv2"], " + invalid-mask fill" if w["kind"] == "dualfisheye" else "",
        "the reference's build_perspective_spec_maps" if kind == "reference" else "the oracle's restatement of the geometry")
    block = {"value": v, "unit": "Mpix/s", "cores": info["cores"], "kind": kind,
             "sample": "%d %s x %d views per timed row, %s, %d cv2 threads of %d host CPUs (%s)" % (
                 info["frames_timed"], "pairs" if w["kind"] == "dualfisheye" else "frames", len(views), what, info["cv2_threads"],
                 info["host_cpus"], info["cpu_model"]),
             "detail": info,
             "rows": {name: {"value": val, "unit": "Mpix/s", "cores": inf["cores"], "cv2_threads": inf["cv2_threads"],
                             "pool_workers": inf["pool_workers"], "frames_timed": inf["frames_timed"]}
                      for name, (val, inf) in res.items()},
             "note": "ffmpeg's v360 (the reference's ERP back end) is not in this image and cannot be timed; the ERP rows time "
                     "the per-frame work of the reference's NumPy/OpenCV path (DF:2001-2014) on ERP maps"}
    return block


# --------------------------------------------------------------------------------------------

def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=30)
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
    ap.add_argument("--no-compressed", action="store_true", help="skip the JPEG-in / JPEG-out end-to-end leg")
    ap.add_argument("--no-variants", action="store_true", help="time the headline workload only")
    ap.add_argument("--workload", default="cfg2", help="headline workload (see workload_table)")
    ns = ap.parse_args()

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if ns.warmup < 3 and ns.impl == "ours":
        ns.warmup = 3
    table = workload_table(ns)
    head = table[ns.workload]
    views = workload_views(head)
    n_views = len(views)
    config = {"workload": workload_text(head, n_views), "baseline_config": head["cfg"], "preset": head["viewset"], "views": n_views,
              "out_size": head["size"], "interp": head["interp"], "frames_per_gpu_per_step": head["frames"],
              "convention": "halfpixel", "content": "uniform noise (seeded)",
              "l2_policy": "inputs (%.2f GB/step) larger than L2" % (head["frames"] * head["src"][0] * head["src"][1] * 3 * (2 if head["dtype"] == "u16" else 1) / 1e9)}

    # ------------------------------------------------------------------ reference arm
    if ns.impl == "reference":
        if rank != 0:
            return
        res, kind = cpu_reference_arm(head, views, ns.cpu_seconds, steps=ns.steps, warmup=ns.warmup, rows=("all_threads",))
        value, info = res["all_threads"]
        sample = "1 %s (all %d views) per step, cv2.remap %s on prebuilt float32 maps, %d cv2 threads (%s)" % (
            "pair" if head["kind"] == "dualfisheye" else "frame", n_views, head["interp"], info["cv2_threads"], info["cpu_model"])
        line = {"impl": "reference", "metric": "output Mpix/s", "value": value, "unit": "Mpix/s",
                "n_gpus": ns.gpus, "steps": ns.steps, "warmup": ns.warmup, "ms_per_step": info["s_per_frame"] * 1e3,
                "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": head["dtype"],
                "data": "synthetic", "config": config,
                "views_per_s": value * 1e6 / (head["size"] * head["size"]),
                "cpu_baseline": {"value": value, "unit": "Mpix/s", "cores": info["cores"], "kind": kind,
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

    peaks_file = ROOT / "MEASURED_PEAKS.json"
    if peaks_file.exists():
        peak, peak_src = float(json.loads(peaks_file.read_text())["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    else:
        peak, peak_src = 6650.0, "fallback (B200_PROFILING.md)"
    digest = csrc_digest()
    T = {"u8": torch.uint8, "u16": torch.uint16, "f16": torch.float16}

    def barrier():
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    def seeded_frames(w, n):
        """Frames are sharded by index: rank r owns global frames [r*B, (r+1)*B); every frame is seeded by its index."""
        W, H = w["src"]
        per = 2 if w["kind"] == "dualfisheye" else 1
        shape = (n, per, H, W, CHANNELS) if per == 2 else (n, H, W, CHANNELS)
        t = torch.empty(shape, dtype=T[w["dtype"]], device=dev)
        for f in range(n):
            g = torch.Generator(device=dev)
            g.manual_seed(1234 + rank * n + f)
            if w["dtype"] == "u16":
                t[f] = torch.randint(0, 65536, shape[1:], dtype=torch.int32, device=dev, generator=g).to(torch.uint16)
            else:
                t[f] = torch.randint(0, 256, shape[1:], dtype=torch.uint8, device=dev, generator=g)
        return t

    class Runner:
        """One workload resident in HBM: step() is one pass of the hot path over its batch."""

        def __init__(self, w):
            self.w, self.vs = w, workload_views(w)
            self.src = seeded_frames(w, w["frames"])
            nv, sz = len(self.vs), w["size"]
            self.out = remap360.alloc_views(w["frames"], nv, sz, sz, CHANNELS, T[w["out"]], dev)
            self.out_dtype = None if w["out"] == w["dtype"] else T[w["out"]]
            if w["kind"] == "dualfisheye":
                cal, _, _ = dualfisheye_setup()
                c = remap360.FisheyeCalibration(**{k: cal[k] for k in ("width", "height", "f", "cx", "cy", "k1", "k2", "k3",
                                                                       "k4", "p1", "p2", "b1", "b2")}, lens_fov_deg=190.0)
                self.calibs = [c, c]
                self.pviews = [remap360.PerspectiveView(y, p, hf, vf, src_slot=slot, view_id=vid) for vid, y, p, hf, vf, slot in self.vs]
            else:
                self.pviews = [remap360.PerspectiveView(y, p, hf, vf, view_id=vid) for vid, y, p, hf, vf, _ in self.vs]
            self.pix_per_step = w["frames"] * nv * sz * sz

        def step(self):
            w = self.w
            if w["kind"] == "dualfisheye":
                remap360.remap_fisheye(self.src, self.calibs, self.pviews, (w["size"], w["size"]), interp=w["interp"], out=self.out,
                                       path=ns.path)
            else:
                remap360.remap_erp(self.src, self.pviews, (w["size"], w["size"]), interp=w["interp"], out=self.out,
                                   out_dtype=self.out_dtype, path=ns.path)

        def timed(self, steps, warmup, sample_clocks):
            """K steps bracketed by barrier + synchronize, CUDA events on the launch stream, max over ranks."""
            sampler = ClockSampler(local_rank) if sample_clocks else None
            if sampler:
                sampler.__enter__()
            for _ in range(warmup):
                self.step()
            if sampler:
                sampler.wait_ready()
            barrier()
            l0 = remap360.launch_count()
            ev = [torch.cuda.Event(enable_timing=True) for _ in range(steps + 1)]
            if sampler:
                sampler.mark_start()
            ev[0].record()
            for k in range(steps):
                self.step()
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

        def roofline(self, per_step_ms, name):
            w = self.w
            per_frame, u_px = algorithmic_bytes(w, len(self.vs))
            if per_frame is None:
                return None
            bytes_per_launch = per_frame * w["frames"]
            kernel_ms = statistics.median(per_step_ms)
            achieved = bytes_per_launch / (kernel_ms * 1e-3) / 1e9
            entry = ncu_traffic().get((name, w["interp"], w["frames"]))
            traffic, traffic_note = None, "no ncu capture recorded for this workload"
            if entry is not None:
                traffic, traffic_note = (entry[0], entry[2]) if entry[1] == digest else (None, "stale: %s was captured with csrc digest %s, the sources are now %s" % (entry[2], entry[1], digest))
            return {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                    "traffic": traffic, "traffic_source": traffic_note, "peak_source": peak_src,
                    "algorithmic_bytes_per_launch": bytes_per_launch, "unique_source_px_per_frame": u_px, "kernel_ms": kernel_ms,
                    "kernel": "remap_tiled_kernel<%s, %s, %s> (one launch per step covers the batch; + remap_fallback_kernel "
                              "for the tiles the plan routes to the direct path)" % (w["interp"], w["dtype"], w["out"])}

    runner = Runner(head)
    ms_per_step, step_ms, launches, clocks = runner.timed(ns.steps, ns.warmup, True)
    out_pix_step = runner.pix_per_step
    value = out_pix_step * world / (ms_per_step * 1e-3) / 1e6          # Mpix/s, whole job
    roofline = runner.roofline(step_ms, ns.workload)

    # ------------------------------------------------------------------ the other BASELINE configurations (kernel-only)
    variants = {}
    if not ns.no_variants:
        # content class (S) of SURVEY.md section 8(d): smooth band-limited frames (eight low-frequency sinusoids in
        # lon / lat per channel, continuous across the seam) instead of noise; the kernels have no data-dependent
        # control flow, so the number should not move
        if head["kind"] == "erp" and head["dtype"] == "u8":
            W, H = head["src"]
            lon = torch.linspace(0.0, 2.0 * torch.pi, W + 1, device=dev)[:-1]
            lat = torch.linspace(-0.5 * torch.pi, 0.5 * torch.pi, H, device=dev)
            gen = torch.Generator(device=dev)
            gen.manual_seed(4321 + rank)
            smooth = torch.zeros((H, W, CHANNELS), dtype=torch.float32, device=dev)
            for _ in range(8):
                kx = torch.randint(1, 9, (CHANNELS,), device=dev, generator=gen).float()
                ky = torch.randint(1, 9, (CHANNELS,), device=dev, generator=gen).float()
                ph = torch.rand((2, CHANNELS), device=dev, generator=gen) * 2.0 * torch.pi
                smooth += torch.sin(lon[None, :, None] * kx + ph[0]) * torch.cos(lat[:, None, None] * ky + ph[1])
            smooth = ((smooth / 16.0 + 0.5).clamp_(0.0, 1.0) * 255.0).round_().to(torch.uint8)
            keep = runner.src.clone()
            runner.src[:] = smooth
            s_ms, _, _, _ = runner.timed(max(3, ns.steps // 4), 3, False)
            variants["content_smooth"] = {"value": out_pix_step * world / (s_ms * 1e-3) / 1e6, "unit": "Mpix/s", "ms_per_step": s_ms,
                                          "interp": head["interp"], "content": "sum of 8 low-frequency sinusoids per channel, seam-continuous"}
            runner.src.copy_(keep)
            del keep, smooth
        for name, w in table.items():
            if name == ns.workload:
                continue
            host_keep = runner                                   # the headline's tensors stay resident for the e2e leg
            try:
                r = Runner(w)
                v_ms, v_steps, _, _ = r.timed(max(5, ns.steps // 3), 3, False)
                variants[name] = {"value": r.pix_per_step * world / (v_ms * 1e-3) / 1e6, "unit": "Mpix/s", "ms_per_step": v_ms,
                                  "workload": workload_text(w, len(r.vs)), "baseline_config": w["cfg"],
                                  "roofline": r.roofline(v_steps, name)}
            except Exception as exc:                              # a variant must never cost the headline its line
                variants[name] = {"error": "%s: %s" % (type(exc).__name__, exc)}
            finally:
                r = None
                torch.cuda.empty_cache()
            del host_keep

    # ------------------------------------------------------------------ end to end (host buffers)
    e2e = None
    if not ns.no_e2e and head["kind"] == "erp":
        W, H = head["src"]
        remapper = StreamingRemapper(runner.pviews, (head["size"], head["size"]), (H, W, CHANNELS), T[head["dtype"]],
                                     interp=head["interp"], out_dtype=T[head["out"]], device=dev, path=ns.path)
        host_frames = [torch.empty((H, W, CHANNELS), dtype=T[head["dtype"]]).pin_memory() for _ in range(min(head["frames"], 4))]
        for k, hf in enumerate(host_frames):
            hf.copy_(runner.src[k].cpu())
        e2e_steps = max(2, min(ns.steps, 5))

        def e2e_step():
            sink = 0
            for res in remapper.run(host_frames[f % len(host_frames)] for f in range(head["frames"])):
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
        es_in, es_out = runner.src.element_size(), runner.out.element_size()
        h2d = head["frames"] * H * W * CHANNELS * es_in
        d2h = out_pix_step * CHANNELS * es_out
        # what the host link gives this process when both directions run flat out (same buffers sizes, pinned, two
        # streams): the denominator of link_frac
        probe = link_probe(torch, dev, barrier)
        e2e = {"value": out_pix_step * world / e2e_s / 1e6, "unit": "Mpix/s", "h2d_bytes_per_step": h2d,
               "d2h_bytes_per_step": d2h, "ms_per_step": e2e_s * 1e3, "steps": e2e_steps,
               "moved_GBps": (h2d + d2h) / e2e_s / 1e9, "link_probe": probe,
               "link_frac": ((h2d + d2h) / e2e_s / 1e9) / probe["bidirectional_GBps"] if probe["bidirectional_GBps"] else None,
               "api": "remap360.stream.StreamingRemapper (pinned ring of %d slots x %d frames, H2D / kernel / D2H on three streams)" % (
                   remapper.depth, remapper.batch)}

    # ------------------------------------------------------------------ end to end with compressed frames
    # JPEG bytes in host memory -> nvJPEG decode on the device -> the same remap -> nvJPEG encode of every view ->
    # JPEG bytes in host memory (what the still-image runner does per panorama, minus the file system).  Fewer bytes
    # cross the host link, but the codec (library code, lossy) now bounds the rate; reported beside the raw leg.
    e2e_compressed = None
    if e2e is not None and head["dtype"] == "u8" and not ns.no_compressed:
        try:
            e2e_compressed = compressed_leg(torch, remap360, dev, runner, head, world, barrier, dist)
        except Exception as exc:
            e2e_compressed = {"error": "%s: %s" % (type(exc).__name__, exc)}

    cpu_baseline = None
    if rank == 0 and world == 1 and not ns.no_cpu_baseline:
        cpu_baseline = cpu_baseline_block(head, views, ns.cpu_seconds)
        if not ns.no_variants:
            for name in ("cfg1", "cfg5"):                         # the reference's own CPU-runnable cases, a few seconds each
                if name in variants and "error" not in variants[name]:
                    try:
                        variants[name]["cpu_baseline"] = cpu_baseline_block(table[name], workload_views(table[name]), min(6.0, ns.cpu_seconds),
                                                                            rows=("all_threads",))
                    except Exception as exc:
                        variants[name]["cpu_baseline"] = {"error": "%s: %s" % (type(exc).__name__, exc)}

    if rank == 0:
        line = {"metric": "output Mpix/s", "value": value, "unit": "Mpix/s", "n_gpus": world, "steps": ns.steps,
                "warmup": ns.warmup, "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak",
                "vs_baseline": None, "dtype": head["dtype"], "data": "synthetic", "config": config,
                "views_per_s": value * 1e6 / (head["size"] * head["size"]), "frames_per_s": value * 1e6 / (n_views * head["size"] * head["size"]),
                "clocks": clocks.summary(), "e2e": e2e, "e2e_compressed": e2e_compressed, "gpu_launches": int(launches),
                "roofline": roofline, "cpu_baseline": cpu_baseline, "variants": variants, "csrc_digest": digest}
        if saved_stdout_fd is not None:
            sys.stdout.flush()
            os.dup2(saved_stdout_fd, 1)
        print(json.dumps(line), flush=True)
    if dist is not None:
        dist.destroy_process_group()


def compressed_leg(torch, remap360, dev, runner, head, world, barrier, dist, panoramas=16, quality=95):
    """JPEG panoramas -> JPEG views through remap360.codec (nvJPEG) and the remap, one host thread per codec."""
    import threading
    from concurrent.futures import ThreadPoolExecutor
    from remap360 import codec
    W, H = head["src"]
    size = head["size"]
    # band-limited content plus a little noise: what a camera frame compresses like (noise alone does not compress)
    lon = torch.linspace(0.0, 2.0 * torch.pi, W + 1, device=dev)[:-1]
    lat = torch.linspace(-0.5 * torch.pi, 0.5 * torch.pi, H, device=dev)
    gen = torch.Generator(device=dev)
    gen.manual_seed(977)
    enc0 = codec.JpegCodec(dev)
    blobs = []
    for k in range(4):
        img = torch.zeros((H, W, CHANNELS), dtype=torch.float32, device=dev)
        for _ in range(6):
            kx = torch.randint(1, 24, (CHANNELS,), device=dev, generator=gen).float()
            ky = torch.randint(1, 24, (CHANNELS,), device=dev, generator=gen).float()
            ph = torch.rand((2, CHANNELS), device=dev, generator=gen) * 2.0 * torch.pi
            img += torch.sin(lon[None, :, None] * kx + ph[0]) * torch.cos(lat[:, None, None] * ky + ph[1])
        img = (img / 12.0 + 0.5) * 255.0 + torch.randn((H, W, CHANNELS), device=dev, generator=gen) * 3.0
        blobs.append(enc0.encode(img.clamp_(0.0, 255.0).round_().to(torch.uint8), 92))
        del img
    del enc0
    threads = max(1, min(16, (os.cpu_count() or 1) // max(1, world)))
    tls = threading.local()

    def one(n):
        if getattr(tls, "codec", None) is None:
            tls.codec, tls.stream = codec.JpegCodec(dev), torch.cuda.Stream(dev)
        with torch.cuda.stream(tls.stream):
            frame = tls.codec.decode(blobs[n % len(blobs)], stream=tls.stream)
            views = remap360.remap_erp(frame[None], runner.pviews, (size, size), interp=head["interp"], stream=tls.stream)[0]
            out = [tls.codec.encode(views[v], quality, stream=tls.stream) for v in range(views.shape[0])]
        return sum(len(b) for b in out)

    with ThreadPoolExecutor(threads) as pool:
        list(pool.map(one, range(2 * threads)))                    # codecs, streams, plans
        barrier()
        t0 = time.perf_counter()
        down = sum(pool.map(one, range(panoramas)))
        torch.cuda.synchronize(dev)
        barrier()
        dt = time.perf_counter() - t0
    if dist is not None:
        t = torch.tensor([dt], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dt = float(t.item())
    up = sum(len(blobs[n % len(blobs)]) for n in range(panoramas))
    pix = panoramas * len(runner.pviews) * size * size
    return {"value": pix * world / dt / 1e6, "unit": "Mpix/s", "panoramas_per_s": panoramas * world / dt, "panoramas": panoramas,
            "host_threads": threads, "h2d_bytes_per_step": up, "d2h_bytes_per_step": down, "ms_per_step": dt * 1e3,
            "jpeg": "nvJPEG (library): decode of quality-92 4:4:4 frames, encode of every view at quality %d 4:4:4" % quality,
            "note": "lossy leg: results are JPEG views, not the bit-checked product; the codec, not the remap or the link, bounds it"}


def link_probe(torch, dev, barrier, mb=256, reps=4):
    """Pinned host <-> device copy rates of this process: each direction alone, then both at once on two streams."""
    n = mb << 20
    h_in, h_out = torch.empty(n, dtype=torch.uint8).pin_memory(), torch.empty(n, dtype=torch.uint8).pin_memory()
    d_in, d_out = torch.empty(n, dtype=torch.uint8, device=dev), torch.empty(n, dtype=torch.uint8, device=dev)
    s1, s2 = torch.cuda.Stream(dev), torch.cuda.Stream(dev)

    def run(up, down):
        barrier()
        t0 = time.perf_counter()
        for _ in range(reps):
            if up:
                with torch.cuda.stream(s1):
                    d_in.copy_(h_in, non_blocking=True)
            if down:
                with torch.cuda.stream(s2):
                    h_out.copy_(d_out, non_blocking=True)
        s1.synchronize(); s2.synchronize()
        return time.perf_counter() - t0

    run(True, True)
    t_up, t_down, t_both = run(True, False), run(False, True), run(True, True)
    return {"h2d_GBps": reps * n / t_up / 1e9, "d2h_GBps": reps * n / t_down / 1e9,
            "bidirectional_GBps": 2 * reps * n / t_both / 1e9, "buffer_MiB": mb}


if __name__ == "__main__":
    main()
