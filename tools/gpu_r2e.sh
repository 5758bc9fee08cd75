#!/bin/bash
set -u
mkdir -p gpurun_out
python tools/diag_attention_race.py 4000 256 > gpurun_out/r2e_race.txt 2>&1; tail -8 gpurun_out/r2e_race.txt
python -m pytest tests -m gpu -x -q > gpurun_out/r2e_tests.log 2>&1; echo "tests rc=$?"; tail -5 gpurun_out/r2e_tests.log
