#!/bin/bash
set -u
mkdir -p gpurun_out
python -m pytest tests/test_gpu_models.py tests/test_gpu_kernels.py -m gpu -x -q > gpurun_out/r2n_tests.log 2>&1; echo "tests rc=$?"; tail -3 gpurun_out/r2n_tests.log
HB_ATTENTION_LEGACY=0 python tools/exp_group_size.py --child 16 | tail -1
HB_ATTENTION_LEGACY=0 python tools/exp_group_size.py --child 16 | tail -1
