#!/bin/bash
# Round-2 profiling session on one box: ncu --set full captures of the kernels that had no counters in round 1
# (VERDICT.md "what's weak" 6, 7), each after the same command has exited 0 without ncu.  Reports land in gpurun_out/.
mkdir -p gpurun_out
NCU="ncu --set full --clock-control none --import-source on"
run() { # name, kernel regex, skip, count, command...
  local name=$1 regex=$2 skip=$3 count=$4; shift 4
  "$@" > gpurun_out/plain_$name.log 2>&1 && $NCU -k regex:"$regex" -s $skip -c $count -o /tmp/r2_$name "$@" > gpurun_out/ncu_$name.log 2>&1
  echo "$name rc=$? $(tail -n 1 gpurun_out/plain_$name.log | cut -c1-200)"
  # only the CSV pages travel back (gpurun_out is limited to 64 MiB): raw metrics per launch, and the per-instruction source page
  if [ -f /tmp/r2_$name.ncu-rep ]; then
    ncu -i /tmp/r2_$name.ncu-rep --page raw --csv > gpurun_out/r2_${name}_raw.csv 2>/dev/null
    ncu -i /tmp/r2_$name.ncu-rep --page source --csv --print-source sass > gpurun_out/r2_${name}_sass.csv 2>/dev/null
    gzip -f gpurun_out/r2_${name}_sass.csv
  fi
}
run terrain 'trace_' 2 2 python tools/prof_trace.py terrain
run instanced 'trace_' 2 2 python tools/prof_trace.py instanced
run pt_shade_unsorted 'pt_shade' 10 3 python tools/pt_iter.py pt 1
IMBA_PT_SORT=1 run pt_shade_sorted 'pt_shade' 10 3 python tools/pt_iter.py pt 1
run vcm_light_shade 'vcm_light_shade' 0 2 python tools/algo_iter.py vcm 1
run vcm_camera_shade 'vcm_camera_shade' 0 2 python tools/algo_iter.py vcm 1
run vcm_grid 'grid_count|grid_scatter|vcm_merge_query' 0 3 python tools/algo_iter.py vcm 1
run lt_splat 'vcm_light_shade' 0 2 python tools/algo_iter.py lt 1
ls -la gpurun_out/ | head -40; du -sh gpurun_out
