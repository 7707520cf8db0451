#!/bin/bash
# Effect of the tree / refill defaults on the integrators: each argument = one environment configuration
for rep in 1 2; do
for cfg in "$@"; do
  echo "== [$cfg]"
  env $cfg python tools/pt_iter.py pt 30 2>&1 | tail -1
  env $cfg python tools/algo_iter.py bpt 5 2>&1 | tail -1
done; done > gpurun_out/cmp.log 2>&1
cat gpurun_out/cmp.log
