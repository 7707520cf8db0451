#!/bin/bash
# A/B with environment knobs: tools/gpu_ab_env.sh "<VAR=val ...>" "<VAR=val ...>" ...   (each argument = one configuration; "" = defaults)
mkdir -p gpurun_out
for rep in 1 2; do for cfg in "$@"; do
  echo "== [$cfg]"; env $cfg timeout 240 python tools/bench_quick.py 2>&1 | tail -1
done; done | tee gpurun_out/ab_env.log
