"""The C ABI used the way a C / C++ host would use it: tests/native/abi_host.c is compiled with gcc against
include/remap360.h, linked with libremap360.so and the CUDA runtime, and run as its own process -- no Python,
no torch in the calling process.  CPU part: the header is valid C and the validation paths answer without a device.
GPU part: the program's output equals the Python host layer's and the oracle's."""

import pathlib
import shutil
import subprocess

import numpy as np
import pytest

from conftest import PKG_DIR, ROOT

CUDA = pathlib.Path("/usr/local/cuda")


@pytest.fixture(scope="module")
def abi_host(tmp_path_factory):
    if not shutil.which("gcc") or not (CUDA / "include" / "cuda_runtime_api.h").exists():
        pytest.skip("gcc / CUDA headers not available")
    from remap360 import _lib
    _lib.load()                                         # makes sure the library is built
    exe = tmp_path_factory.mktemp("native") / "abi_host"
    libdir = PKG_DIR / "remap360"
    cmd = ["gcc", "-std=c99", "-Wall", "-Werror", "-O1", str(ROOT / "tests" / "native" / "abi_host.c"),
           "-I", str(ROOT / "include"), "-I", str(CUDA / "include"), "-L", str(libdir), "-L", str(CUDA / "lib64"),
           "-Wl,-rpath," + str(libdir), "-Wl,-rpath," + str(CUDA / "lib64"), "-l:libremap360.so", "-lcudart", "-o", str(exe)]
    res = subprocess.run(cmd, capture_output=True, text=True)
    assert res.returncode == 0, res.stderr
    return exe


def test_header_is_valid_c_and_validation_needs_no_device(abi_host):
    res = subprocess.run([str(abi_host), "check"], capture_output=True, text=True)
    assert res.returncode == 0, res.stdout + res.stderr
    assert "abi_host check ok" in res.stdout


@pytest.mark.gpu
@pytest.mark.parametrize("mode", ["erp", "plan"])
def test_c_caller_equals_python_layer_and_oracle(abi_host, tmp_path, mode):
    torch = pytest.importorskip("torch")
    import remap360
    from oracle import geometry as geo, sampler
    rng = np.random.default_rng(11)
    W, H, size = 2048, 1024, 192
    frame = rng.integers(0, 256, (H, W, 3), dtype=np.uint8)
    (tmp_path / "in.raw").write_bytes(frame.tobytes())
    res = subprocess.run([str(abi_host), mode, str(tmp_path / "in.raw"), str(W), str(H), str(tmp_path / "out.raw"), str(size)],
                         capture_output=True, text=True)
    assert res.returncode == 0, res.stdout + res.stderr
    assert "launches:" in res.stdout and int(res.stdout.split("launches:")[1].split()[0]) >= 1
    got = np.frombuffer((tmp_path / "out.raw").read_bytes(), dtype=np.uint8).reshape(3, size, size, 3)
    fov = 104.2500326978036
    specs = [(0.0, 0.0), (90.0, 0.0), (180.0, 30.0)]
    views = [remap360.PerspectiveView(y, p, fov, fov) for y, p in specs]
    ours = remap360.remap_erp(torch.from_numpy(frame).cuda()[None], views, (size, size), interp="cubic",
                              path="direct" if mode == "erp" else "tiled")[0].cpu().numpy()
    assert np.array_equal(got, ours)                    # same library, same entry points: identical bytes
    for k, (yaw, pitch) in enumerate(specs):
        mx, my = geo.erp_map64(W, H, size, size, yaw, pitch, fov, fov)
        want = sampler.sample(frame, mx, my, "cubic", "erp")
        d = np.abs(got[k].astype(np.int64) - want.astype(np.int64))
        assert (d <= 1).mean() >= 0.999 and (d == 0).mean() >= 0.999
