#!/bin/bash
# compute-sanitizer over the kernel tests; logs under gpurun_out/, summaries copied to profiles/ by hand
set -u
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 7 --log-file gpurun_out/r02_sanitizer_memcheck.log \
  python -m pytest tests/test_gpu_kernels.py -m gpu -x -q -k "attention or mlp or gemm or layernorm or im2col or embed" -p no:cacheprovider > gpurun_out/r02_sanitizer_memcheck_pytest.log 2>&1
echo "memcheck rc=$?"; tail -3 gpurun_out/r02_sanitizer_memcheck_pytest.log; tail -4 gpurun_out/r02_sanitizer_memcheck.log
timeout 600 compute-sanitizer --tool memcheck --error-exitcode 7 --log-file gpurun_out/r02_sanitizer_memcheck_clam.log \
  python -m pytest tests/test_gpu_clam_train.py tests/test_gpu_kernels.py -m gpu -x -q -k "clam" -p no:cacheprovider > gpurun_out/r02_sanitizer_memcheck_clam_pytest.log 2>&1
echo "memcheck clam rc=$?"; tail -3 gpurun_out/r02_sanitizer_memcheck_clam_pytest.log; tail -4 gpurun_out/r02_sanitizer_memcheck_clam.log
timeout 600 compute-sanitizer --tool racecheck --error-exitcode 7 --log-file gpurun_out/r02_sanitizer_racecheck.log \
  python -m pytest tests/test_gpu_kernels.py -m gpu -x -q -k "layernorm or im2col or clam" -p no:cacheprovider > gpurun_out/r02_sanitizer_racecheck_pytest.log 2>&1
echo "racecheck rc=$?"; tail -3 gpurun_out/r02_sanitizer_racecheck_pytest.log; tail -6 gpurun_out/r02_sanitizer_racecheck.log
