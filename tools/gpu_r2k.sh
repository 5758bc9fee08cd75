#!/bin/bash
set -u
mkdir -p gpurun_out
python -m pytest tests/test_gpu_clam_train.py tests/test_gpu_ingest.py -m gpu -x -q > gpurun_out/r2k_tests.log 2>&1; echo "tests rc=$?"; tail -15 gpurun_out/r2k_tests.log
