"""Frame sharding across GPUs: one process per GPU, contiguous frame ranges, no data-path
collective (every output tile depends on one source frame only).

The reference's analogue is its per-(file, view) process pool (gs360_360PerspCut.py:1049-1051) and
its per-pair thread pool (gs360_DualFisheyeDistortionCalibration.py:2761-2810).  Contiguous ranges
keep the ``%07d`` output numbering of video frames deterministic (gs360_360PerspCut.py:746-749) and
all views of a frame on one GPU, so a frame is uploaded once."""

from __future__ import annotations

import os
from typing import Tuple


def shard_range(n_items: int, world_size: int, rank: int) -> Tuple[int, int]:
    """[start, stop) of the items owned by `rank`: item i belongs to rank i // ceil(n / world)."""
    if world_size <= 0 or not (0 <= rank < world_size):
        raise ValueError("bad rank %d / world size %d" % (rank, world_size))
    if n_items <= 0:
        return 0, 0
    per = -(-n_items // world_size)
    start = min(n_items, rank * per)
    return start, min(n_items, start + per)


def env_rank() -> Tuple[int, int, int]:
    """(rank, world_size, local_rank) from the torchrun environment (defaults: single process)."""
    return (int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1")),
            int(os.environ.get("LOCAL_RANK", "0")))


def max_over_ranks(value: float, device=None) -> float:
    """Max of a host scalar over all ranks (timing is reported as the slowest rank's)."""
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return float(value)
    t = torch.tensor([float(value)], dtype=torch.float64, device=device or "cpu")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def sum_over_ranks(value: float, device=None) -> float:
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return float(value)
    t = torch.tensor([float(value)], dtype=torch.float64, device=device or "cpu")
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return float(t.item())


def bind_to_gpu_cpus(device_index: int) -> bool:
    """Pin the calling thread (and the pinned host memory it allocates afterwards, by first touch) to the CPUs
    NVML reports as closest to the GPU.  One process per GPU shares the box's host memory and PCIe root complexes;
    without this every rank's staging buffers can land on one socket.  Returns False when NVML is unavailable."""
    try:
        import pynvml
        pynvml.nvmlInit()
        handle = pynvml.nvmlDeviceGetHandleByIndex(int(device_index))
        pynvml.nvmlDeviceSetCpuAffinity(handle)
        return True
    except Exception:
        return False
