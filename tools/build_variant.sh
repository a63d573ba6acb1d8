#!/bin/bash
# Kernel experiments: build libremap360 with extra -D flags into tools/variants/lib_<name>.so (git-ignored, travels
# to the GPU box).   usage: tools/build_variant.sh <name> [-DR360_TILED_TEAMS=2 ...]
set -e
name=$1; shift
root=$(cd "$(dirname "$0")/.." && pwd)
pkg="$root/360cam-pgm-3dgs-tools_b200"
mkdir -p "$root/tools/variants" "$pkg/build"
[ -f "$pkg/build/weights.o" ] || g++ -O2 -std=c++17 -fPIC -ffp-contract=off -c "$pkg/csrc/weights.cpp" -o "$pkg/build/weights.o"
nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -fmad=false -Xcompiler -fPIC,-ffp-contract=off -Xptxas -v \
     -split-compile 0 "$@" -shared "$pkg/csrc/remap360.cu" "$pkg/build/weights.o" -o "$root/tools/variants/lib_$name.so" \
     > "$root/tools/variants/lib_$name.log" 2>&1
grep -A3 "remap_tiled_kernelILi[12]EhhLi" "$root/tools/variants/lib_$name.log" | grep -E "Compiling|registers" | sed 's/.*remap_tiled_kernel/  /; s/EEvNS.*//' | paste - - | awk '{print $1, $6, "regs"}'
echo "built tools/variants/lib_$name.so"
