"""BASELINE.json config 3: a 600-frame synthetic 8K ERP video, fisheyelike preset (10 views, 17 mm, 1600^2),
frames sharded contiguously over the GPUs of one box (rank r owns frames [r * ceil(600 / G), ...)), no collective
on the data path.

    python tools/config3_sweep.py                                   # G = 1
    python -m torch.distributed.run --nnodes=1 --nproc-per-node G --master-addr 127.0.0.1 tools/config3_sweep.py

Two measurements, both as the slowest rank's time (CUDA events, barrier + synchronize on both sides):
  resident  the rank's whole shard lives in HBM (53.1 GB at G = 1), cut chunk by chunk into a reused output buffer;
  streamed  frames come from pinned host memory through StreamingRemapper (H2D / kernel / D2H overlapped), on a
            bounded number of frames per rank.
Rank 0 prints one JSON line."""
import argparse
import json
import os
import pathlib
import sys

import torch

ROOT = pathlib.Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT / "360cam-pgm-3dgs-tools_b200"))
sys.path.insert(0, str(ROOT))
import remap360  # noqa: E402
from bench import preset_views  # noqa: E402
from remap360.sharding import env_rank, shard_range  # noqa: E402
from remap360.stream import StreamingRemapper  # noqa: E402

W, H, SIZE = 7680, 3840, 1600


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--frames", type=int, default=600)
    ap.add_argument("--chunk", type=int, default=24)
    ap.add_argument("--stream-frames", type=int, default=60, help="frames per rank in the host-streamed pass")
    ap.add_argument("--interp", default="cubic")
    ns = ap.parse_args()
    rank, world, local_rank = env_rank()
    bound = os.environ.get("R360_NUMA_BIND", "0") == "1" and __import__("remap360.sharding", fromlist=["x"]).bind_to_gpu_cpus(local_rank)
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    dist = None
    saved_fd = None
    if world > 1:
        import torch.distributed as dist
        sys.stdout.flush()
        saved_fd = os.dup(1)
        os.dup2(2, 1)                                   # NCCL's banner must not share stdout with the JSON line
        dist.init_process_group("nccl", device_id=dev)

    def barrier():
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    def max_ranks(x):
        if dist is None:
            return x
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    views = [remap360.PerspectiveView(y, p, hf, vf, view_id=vid) for vid, y, p, hf, vf in preset_views("fisheyelike", SIZE)]
    lo, hi = shard_range(ns.frames, world, rank)
    n_own = hi - lo
    frames = torch.empty((n_own, H, W, 3), dtype=torch.uint8, device=dev)
    for k in range(n_own):
        g = torch.Generator(device=dev)
        g.manual_seed(1234 + lo + k)                    # seeded per global frame index: any sharding sees the same video
        frames[k] = torch.randint(0, 256, (H, W, 3), dtype=torch.uint8, device=dev, generator=g)
    out = torch.empty((ns.chunk, len(views), SIZE, SIZE, 3), dtype=torch.uint8, device=dev)

    def sweep():
        for c0 in range(0, n_own, ns.chunk):
            c1 = min(n_own, c0 + ns.chunk)
            remap360.remap_erp(frames[c0:c1], views, (SIZE, SIZE), interp=ns.interp, out=out[:c1 - c0])

    remap360.remap_erp(frames[:min(n_own, ns.chunk)], views, (SIZE, SIZE), interp=ns.interp,
                       out=out[:min(n_own, ns.chunk)])  # plan build + warm-up
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    sweep()
    e1.record()
    barrier()
    resident_ms = max_ranks(e0.elapsed_time(e1))
    checksum = int(out[:1].to(torch.int64).sum().item())

    # host-streamed pass: the rank's first frames, from pinned host memory
    n_stream = min(n_own, ns.stream_frames)
    host = [frames[k].cpu().pin_memory() for k in range(min(n_stream, 6))]
    rem = StreamingRemapper(views, (SIZE, SIZE), (H, W, 3), torch.uint8, interp=ns.interp, device=dev)
    for _ in rem.run(host[k % len(host)] for k in range(6)):
        pass
    barrier()
    import time
    t0 = time.perf_counter()
    for _ in rem.run(host[k % len(host)] for k in range(n_stream)):
        pass
    torch.cuda.synchronize()
    streamed_s = max_ranks(time.perf_counter() - t0)
    n_stream_total = n_stream
    if dist is not None:
        t = torch.tensor([n_stream], dtype=torch.float64, device=dev)
        dist.all_reduce(t)
        n_stream_total = int(t.item())

    if rank == 0:
        pix_per_frame = len(views) * SIZE * SIZE
        line = {"config": "BASELINE configs[2]: 600-frame 8K u8 ERP sweep, fisheyelike (10 views 1600^2), %s" % ns.interp,
                "n_gpus": world, "numa_bound": bool(bound), "cpus_allowed": len(os.sched_getaffinity(0)), "frames": ns.frames, "frames_per_rank": -(-ns.frames // world), "chunk": ns.chunk,
                "resident": {"ms": resident_ms, "frames_per_s": ns.frames / (resident_ms * 1e-3),
                             "Mpix_per_s": ns.frames * pix_per_frame / (resident_ms * 1e-3) / 1e6,
                             "hbm_bytes_frames_per_rank": int(frames.numel())},
                "streamed": {"frames": n_stream_total, "s": streamed_s, "frames_per_s": n_stream_total / streamed_s,
                             "Mpix_per_s": n_stream_total * pix_per_frame / streamed_s / 1e6,
                             "h2d_bytes_per_frame": H * W * 3, "d2h_bytes_per_frame": pix_per_frame * 3},
                "checksum_last_chunk_rank0": checksum}
        if saved_fd is not None:
            sys.stdout.flush()
            os.dup2(saved_fd, 1)
        print(json.dumps(line), flush=True)
    if dist is not None:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
