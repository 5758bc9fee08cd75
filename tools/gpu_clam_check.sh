#!/bin/bash
# CLAM tensor-core kernel with the gate on the tensor core: accuracy, tests, config-4 timings (every command under its own timeout)
set -u
mkdir -p gpurun_out
for cfg in "hipt_smaller 1" "hipt_smaller 5" "hipt_small 2" "hipt_medium 1"; do timeout 120 python tools/check_clam_tc_accuracy.py $cfg 2>&1 | tail -1; done
timeout 600 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_models.py tests/test_gpu_clam_train.py -m gpu -x -q -k "clam or CLAM or pool or bag or smoke or pipeline" > gpurun_out/r2q_tests.log 2>&1; echo "clam tests rc=$?"; tail -4 gpurun_out/r2q_tests.log
for f in 1 2 5; do timeout 120 python tools/bench_clam.py --size hipt_smaller --folds $f 2>&1 | grep '"folds"' | cut -c1-200; done
timeout 120 python tools/bench_clam.py --size hipt_small --folds 1 2>&1 | grep '"folds"' | cut -c1-200
timeout 120 python tools/bench_clam.py --size hipt_small --folds 2 2>&1 | grep '"folds"' | cut -c1-200
timeout 120 python tools/bench_clam.py --size hipt_medium --folds 1 2>&1 | grep '"folds"' | cut -c1-200
