#!/usr/bin/env python3
"""Build libremap360.so (sm_100a only) in-tree.

    python 360cam-pgm-3dgs-tools_b200/build.py [--force] [--verbose]

nvcc cross-compiles without a GPU.  The library lands next to the Python host layer
(remap360/libremap360.so) so that it travels with the source tree."""

import argparse
import hashlib
import os
import pathlib
import shutil
import subprocess
import sys

HERE = pathlib.Path(__file__).resolve().parent
CSRC = HERE / "csrc"
OUT = HERE / "remap360" / "libremap360.so"
STAMP = HERE / "remap360" / ".libremap360.stamp"
CODEC_OUT = HERE / "remap360" / "libr360codec.so"

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-lineinfo", "-O3", "-std=c++17", "-fmad=false",
    "-Xcompiler", "-fPIC,-ffp-contract=off",
    "-Xptxas", "-v", "-split-compile", "0",
]


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and pathlib.Path(cand).exists():
            return cand
    raise RuntimeError("nvcc not found")


def _sources():
    return sorted(list(CSRC.glob("*.cu")) + list(CSRC.glob("*.cuh")) + list(CSRC.glob("*.cpp")) +
                  list(CSRC.glob("*.h")) + list((HERE / "codec").glob("*.cpp")) +
                  list((HERE.parent / "include").glob("*.h")) + [pathlib.Path(__file__)])


def _digest() -> str:
    h = hashlib.sha256()
    for p in _sources():
        h.update(p.name.encode())
        h.update(p.read_bytes())
    return h.hexdigest()


def build_library(force: bool = False, verbose: bool = False) -> pathlib.Path:
    digest = _digest()
    if not force and OUT.exists() and CODEC_OUT.exists() and STAMP.exists() and STAMP.read_text().strip() == digest:
        return OUT
    build_dir = HERE / "build"
    build_dir.mkdir(exist_ok=True)
    wobj = build_dir / "weights.o"
    cmds = [
        ["g++", "-O2", "-std=c++17", "-fPIC", "-ffp-contract=off", "-c", str(CSRC / "weights.cpp"), "-o", str(wobj)],
        [_nvcc(), *NVCC_FLAGS, "-shared", str(CSRC / "remap360.cu"), str(wobj), "-o", str(OUT)],
        # JPEG codec glue over nvJPEG (library code, linked statically so that nothing but the CUDA driver is
        # needed at run time); no device code of ours in it
        [_nvcc(), "-O2", "-std=c++17", "-Xcompiler", "-fPIC", "-shared", "-I", str(HERE.parent / "include"),
         str(HERE / "codec" / "jpeg_codec.cpp"), "-lnvjpeg_static", "-lculibos", "-o", str(CODEC_OUT)],
    ]
    for cmd in cmds:
        res = subprocess.run(cmd, capture_output=True, text=True)
        if verbose or res.returncode:
            sys.stderr.write(" ".join(cmd) + "\n" + res.stdout + res.stderr)
        if res.returncode:
            raise RuntimeError("build failed: " + " ".join(cmd))
        (build_dir / (pathlib.Path(cmd[-1]).name + ".log")).write_text(res.stdout + res.stderr)
    STAMP.write_text(digest + "\n")
    return OUT


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--force", action="store_true")
    ap.add_argument("--verbose", action="store_true")
    ns = ap.parse_args()
    print(build_library(ns.force, ns.verbose))
