#!/bin/bash
set -u
mkdir -p gpurun_out
T=${1:-r02b}
python -m pytest tests -m gpu -x -q > gpurun_out/${T}_tests.log 2>&1; echo "tests rc=$?"; tail -3 gpurun_out/${T}_tests.log
python bench.py > gpurun_out/${T}_bench.json 2> gpurun_out/${T}_bench.err; echo "bench rc=$?"; tail -5 gpurun_out/${T}_bench.err
python tools/show_bench.py gpurun_out/${T}_bench.json | cut -c1-400
