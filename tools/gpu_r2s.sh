#!/bin/bash
# per-kernel durations of the config-4 CLAM forward (work table, score kernel, combine)
set -u
mkdir -p gpurun_out
for f in 1 5; do
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:clam -s 9 -c 9 --csv --log-file gpurun_out/r02s_clam_launches_$f.csv python tools/bench_clam.py --size hipt_smaller --folds $f > /dev/null 2>&1
grep -o '"hb::[a-z_0-9]*[^"]*".*' gpurun_out/r02s_clam_launches_$f.csv | awk -F'","' '{print $1, $NF}' | tail -9
done
