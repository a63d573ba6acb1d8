"""Shared test plumbing: marker registration, import paths, fixture loaders."""

import json
import pathlib
import sys

import numpy as np
import pytest

ROOT = pathlib.Path(__file__).resolve().parent.parent
PKG_DIR = ROOT / "360cam-pgm-3dgs-tools_b200"
GOLDEN = ROOT / "tests" / "golden"

for p in (str(ROOT), str(PKG_DIR)):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def golden_dir():
    return GOLDEN


@pytest.fixture(scope="session")
def golden_views():
    return json.loads((GOLDEN / "perspcut_views.json").read_text())


@pytest.fixture(scope="session")
def golden_helpers():
    return json.loads((GOLDEN / "perspcut_helpers.json").read_text())


@pytest.fixture(scope="session")
def golden_df():
    return json.loads((GOLDEN / "dualfisheye.json").read_text())


@pytest.fixture(scope="session")
def golden_df_maps():
    return np.load(GOLDEN / "dualfisheye_maps.npz")


@pytest.fixture(scope="session")
def golden_undistort():
    return json.loads((GOLDEN / "undistort.json").read_text()), np.load(GOLDEN / "undistort_maps.npz")


@pytest.fixture(scope="session")
def golden_cv2():
    return np.load(GOLDEN / "cv2_remap.npz")
