"""The C-ABI library builds, loads and exports exactly what include/remap360.h declares.
No GPU is needed: nothing here launches a kernel."""

import ctypes
import pathlib
import re
import shutil
import subprocess

import numpy as np
import pytest

from conftest import PKG_DIR, ROOT


@pytest.fixture(scope="session")
def lib():
    import importlib.util
    spec = importlib.util.spec_from_file_location("r360_build", PKG_DIR / "build.py")
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    if shutil.which("nvcc") or pathlib.Path("/usr/local/cuda/bin/nvcc").exists():
        mod.build_library()
    from remap360 import _lib
    return _lib.load()


def _declared_functions():
    text = (ROOT / "include" / "remap360.h").read_text()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(r360_[a-z0-9_]+)\s*\(", text)))


def test_header_and_binding_list_the_same_symbols():
    from remap360 import _lib
    assert _declared_functions() == sorted(_lib.EXPORTS)


def test_library_exports_every_declared_symbol(lib):
    for name in _declared_functions():
        assert hasattr(lib, name), name
    out = subprocess.run(["nm", "-D", "--defined-only", str(PKG_DIR / "remap360" / "libremap360.so")],
                         capture_output=True, text=True).stdout
    for name in _declared_functions():
        assert re.search(r"\bT %s\b" % name, out), name


def test_library_contains_sm100a_code_only(lib):
    cuobjdump = shutil.which("cuobjdump") or "/usr/local/cuda/bin/cuobjdump"
    if not pathlib.Path(cuobjdump).exists():
        pytest.skip("cuobjdump not available")
    out = subprocess.run([cuobjdump, "-lelf", str(PKG_DIR / "remap360" / "libremap360.so")],
                         capture_output=True, text=True).stdout
    archs = set(re.findall(r"sm_(\d+a?)", out))
    assert archs == {"100a"}, archs


def test_error_strings_and_defaults(lib):
    from remap360 import _lib
    assert lib.r360_abi_version() == 2
    assert lib.r360_error_string(0) == b"ok"
    assert b"invalid" in lib.r360_error_string(-1)
    opt = _lib.default_options()
    # reference defaults: cubic (PC:730, DF:232), mask on (DF:258), mask value 0 (DF:262)
    assert (opt.interp, opt.fill_invalid, opt.border_value, opt.out_dtype) == (2, 1, 0.0, -1)


def test_argument_validation_needs_no_device(lib):
    from remap360 import _lib
    v = (_lib.View * 1)(_lib.View(0, 0, 0, 90, 90, 0, 0))
    img = _lib.Images(data=None, width=8, height=8, channels=3, dtype=0, pitch_bytes=24,
                      image_stride_bytes=192, count=1, reserved=0)
    assert lib.r360_remap_erp(ctypes.byref(img), ctypes.byref(img), v, 1, None, None) == -1   # null data
    img.data = 4096
    img.channels = 5
    assert lib.r360_remap_erp(ctypes.byref(img), ctypes.byref(img), v, 1, None, None) == -2
    img.channels = 3
    img.pitch_bytes = 8
    assert lib.r360_remap_erp(ctypes.byref(img), ctypes.byref(img), v, 1, None, None) == -1
    img.pitch_bytes = 24
    assert lib.r360_remap_erp(ctypes.byref(img), ctypes.byref(img), v, 0, None, None) == -1


def test_host_built_weight_tables_equal_the_oracles(lib):
    """The tables the kernels use are built on the host (csrc/weights.cpp); they must equal the
    oracle's, which are pinned bit-exactly to cv2.remap (tests/test_oracle_sampler.py)."""
    from oracle import sampler
    fixed = np.zeros(32 * 32 * 16, np.int16)
    one_d = np.zeros(32 * 4, np.float32)
    assert lib.r360_debug_weight_tables(fixed.ctypes.data, one_d.ctypes.data) == 0
    t1, _, itab = sampler.tables("cubic")
    assert np.array_equal(one_d.reshape(32, 4), t1)
    assert np.array_equal(fixed.reshape(32, 32, 4, 4), itab)


def test_host_built_lanczos4_tables_equal_the_oracles(lib):
    from oracle import sampler
    fixed = np.zeros(32 * 32 * 64, np.int16)
    one_d = np.zeros(32 * 8, np.float32)
    assert lib.r360_debug_weight_tables_lanczos4(fixed.ctypes.data, one_d.ctypes.data) == 0
    t1, _, itab = sampler.tables("lanczos4")
    assert np.array_equal(one_d.reshape(32, 8), t1)
    assert np.array_equal(fixed.reshape(32, 32, 8, 8), itab)
    assert (fixed.reshape(1024, 64).astype(np.int64).sum(axis=1) == 32768).all()


def test_python_api_refuses_cpu_tensors():
    torch = pytest.importorskip("torch")
    import remap360
    with pytest.raises(ValueError, match="no CPU path"):
        remap360.remap_erp(torch.zeros((1, 8, 16, 3), dtype=torch.uint8),
                           [remap360.PerspectiveView(0, 0, 90, 90)], (4, 4))


def test_shared_memory_ring_allocator_never_overlaps(lib):
    """The producer warp's patch allocator (PatchRing in csrc/r360_tiled.cuh), driven on the host:
    random patch sizes including empty requests (fallback tiles) and patches close to the ring
    capacity -- the mix that occurs around the poles."""
    import ctypes
    import random
    lib.r360_debug_ring_check.argtypes = [ctypes.c_void_p, ctypes.c_int32, ctypes.c_int32, ctypes.c_int32]
    cap = 100 * 1024
    for seed in range(300):
        rnd = random.Random(seed)
        sizes = []
        for _ in range(400):
            r = rnd.random()
            if r < 0.15:
                sizes.append(0)
            elif r < 0.5:
                sizes.append(rnd.randrange(8, 16) * 1024)
            elif r < 0.8:
                sizes.append(rnd.randrange(30, 80) * 1024 // 128 * 128)
            else:
                sizes.append(rnd.randrange(80, 101) * 1024 // 128 * 128)
        arr = np.asarray(sizes, dtype=np.int32)
        assert lib.r360_debug_ring_check(arr.ctypes.data, len(sizes), cap, 8) == -1, seed
    one = np.asarray([cap + 128], dtype=np.int32)
    assert lib.r360_debug_ring_check(one.ctypes.data, 1, cap, 8) == 0
