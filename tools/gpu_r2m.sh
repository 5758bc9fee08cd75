#!/bin/bash
set -u
mkdir -p gpurun_out
python -m pytest tests/test_gpu_kernels.py tests/test_gpu_models.py tests/test_gpu_clam_train.py -m gpu -x -q -k "clam or CLAM or pool or bag or smoke or pipeline" > gpurun_out/r2m_tests.log 2>&1; echo "clam tests rc=$?"; tail -4 gpurun_out/r2m_tests.log
python tools/check_clam_tc_accuracy.py 2>&1 | tail -2
for f in 1 2 5; do python tools/bench_clam.py --size hipt_smaller --folds $f 2>&1 | grep '"folds"' | cut -c1-300; done
python tools/bench_clam.py --size hipt_small --folds 1 2>&1 | grep '"folds"' | cut -c1-300
python tools/bench_clam.py --size hipt_small --folds 2 2>&1 | grep '"folds"' | cut -c1-300
