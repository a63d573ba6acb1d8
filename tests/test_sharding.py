"""Frame sharding host logic, including a real 2-process gloo run (no GPU needed)."""

import os
import socket
import subprocess
import sys
import textwrap

import pytest

from conftest import PKG_DIR
from remap360 import sharding


def test_shard_ranges_partition_the_frames():
    for n in (0, 1, 7, 600, 601, 1024):
        for world in (1, 2, 4, 8):
            seen = []
            for rank in range(world):
                a, b = sharding.shard_range(n, world, rank)
                assert 0 <= a <= b <= n
                seen.extend(range(a, b))
            assert seen == list(range(n))            # contiguous, ordered, no overlap, nothing lost
    # the survey's rule: frame i goes to rank i // ceil(600 / G)
    for world in (2, 4, 8):
        per = -(-600 // world)
        for i in (0, 74, 75, 299, 300, 599):
            owner = [r for r in range(world) if sharding.shard_range(600, world, r)[0] <= i < sharding.shard_range(600, world, r)[1]]
            assert owner == [i // per]
    with pytest.raises(ValueError):
        sharding.shard_range(10, 2, 2)


WORKER = textwrap.dedent("""
    import os, sys
    sys.path.insert(0, {pkg!r})
    import torch.distributed as dist
    from remap360 import sharding
    dist.init_process_group("gloo")
    rank, world, _ = sharding.env_rank()
    a, b = sharding.shard_range(600, world, rank)
    frames = sharding.sum_over_ranks(b - a)                 # every frame is owned exactly once
    slowest = sharding.max_over_ranks(10.0 + rank)          # timing = max over ranks
    checksum = sharding.sum_over_ranks(sum(range(a, b)))
    dist.barrier()
    if rank == 0:
        print("RESULT", int(frames), slowest, int(checksum))
    dist.destroy_process_group()
""")


def test_two_process_gloo_sharding(tmp_path):
    pytest.importorskip("torch")
    script = tmp_path / "worker.py"
    script.write_text(WORKER.format(pkg=str(PKG_DIR)))
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    env = dict(os.environ, MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    out = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
                          "--master-addr", "127.0.0.1", "--master-port", str(port), str(script)],
                         capture_output=True, text=True, timeout=240, env=env)
    assert out.returncode == 0, out.stderr[-2000:]
    line = [ln for ln in out.stdout.splitlines() if ln.startswith("RESULT")][0].split()
    assert int(line[1]) == 600 and float(line[2]) == 11.0 and int(line[3]) == sum(range(600))
