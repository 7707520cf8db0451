#!/bin/bash
# One GPU-box session: parity tests, smoke, bench, then ncu launch list + one full capture of the traversal kernel.
set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi.txt 2>&1
python -m pytest tests -x -q -m gpu > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -15 gpurun_out/pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/smoke.log
tail -3 gpurun_out/smoke.log
python bench.py --steps 10 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"
cat gpurun_out/bench.json; tail -5 gpurun_out/bench.err
# at most ONE profiler pass per box session, and only after the same command exited 0 without it
if [ "$1" = "list" ]; then
python bench.py --steps 2 --warmup 3 > gpurun_out/plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 3 > gpurun_out/ncu_list.log 2>&1
tail -3 gpurun_out/ncu_list.log
fi
if [ "$1" = "full" ]; then
python bench.py --steps 2 --warmup 3 > gpurun_out/plain2.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:trace_.*kernel -s 29 -c 4 -o gpurun_out/prof python bench.py --steps 2 --warmup 3 > gpurun_out/ncu_full.log 2>&1
tail -3 gpurun_out/ncu_full.log
fi
