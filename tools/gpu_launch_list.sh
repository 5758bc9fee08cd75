#!/bin/bash
# ncu launch list of the short bench (same workload) after the CLAM rewrite
set -u
mkdir -p gpurun_out
SHORT="python bench.py --steps 1 --warmup 3 --regions 64 --no-cpu-baseline --no-sections"
timeout 600 $SHORT > gpurun_out/r02y_short.json 2> gpurun_out/r02y_short.err; echo "short rc=$?"; cut -c1-300 gpurun_out/r02y_short.json
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/r02y_launches.csv $SHORT > gpurun_out/r02y_ncu_ll.log 2>&1
echo "launch list rc=$?"; wc -l gpurun_out/r02y_launches.csv
