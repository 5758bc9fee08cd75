#!/bin/bash
# round-2 evidence call: GPU tests, smoke, default bench, reference arm, ncu launch list of a short bench
set -u
mkdir -p gpurun_out
T=${1:-r02a}
python -m pytest tests -m gpu -x -q > gpurun_out/${T}_tests.log 2>&1; echo "tests rc=$?"; tail -5 gpurun_out/${T}_tests.log
python __graft_entry__.py smoke > gpurun_out/${T}_smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/${T}_smoke.log
python bench.py > gpurun_out/${T}_bench.json 2> gpurun_out/${T}_bench.err; echo "bench rc=$?"; tail -5 gpurun_out/${T}_bench.err
python tools/show_bench.py gpurun_out/${T}_bench.json
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/${T}_bench_reference.json 2> gpurun_out/${T}_bench_reference.err; echo "ref rc=$?"; tail -c 600 gpurun_out/${T}_bench_reference.json
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/${T}_launches.csv python bench.py --regions 64 --steps 1 --warmup 3 --no-cpu-baseline --no-sections > gpurun_out/${T}_ncu.log 2>&1; echo "ncu rc=$?"
