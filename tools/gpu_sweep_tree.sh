#!/bin/bash
# Sweep of the tree-shape knobs (IMBA_MAX_LEAF x IMBA_COST_PRIM) on the bench batches: tools/gpu_sweep_tree.sh "<leafs>" "<costs>"
mkdir -p gpurun_out
for ml in $1; do for cp in $2; do
  echo "== max_leaf $ml cost_prim $cp"; IMBA_MAX_LEAF=$ml IMBA_COST_PRIM=$cp timeout 240 python tools/bench_quick.py 2>&1 | tail -1
done; done | tee gpurun_out/tree_sweep.log
