#!/bin/bash
# usage: tools/run_variants.sh name...   (libraries tools/variants/lib_<name>.so; "shipped" = the in-tree one)
mkdir -p gpurun_out
for n in "$@"; do
  if [ "$n" = shipped ]; then python tools/variant_check.py; else R360_LIBRARY=$PWD/tools/variants/lib_$n.so python tools/variant_check.py; fi
done 2>&1 | grep -v Warning | tee -a gpurun_out/variants.jsonl
