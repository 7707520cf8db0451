#!/bin/bash
# A/B timing of traversal kernel variants on one box: tools/gpu_ab.sh <variant> <variant> ...   ("default" = libimba_b200.so)
mkdir -p gpurun_out
for v in "$@"; do
  if [ "$v" = "default" ]; then unset IMBA_LIB_VARIANT; else export IMBA_LIB_VARIANT=libimba_var_$v.so; fi
  echo "== $v"; timeout 240 python tools/bench_quick.py 2>&1 | tail -1
done | tee gpurun_out/ab.log
