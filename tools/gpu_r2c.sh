#!/bin/bash
set -u
mkdir -p gpurun_out
python -m pytest tests/test_gpu_kernels.py -m gpu -x -q -k attention > gpurun_out/r2c_tests.log 2>&1; echo "attention tests rc=$?"; tail -5 gpurun_out/r2c_tests.log
python tools/run_kernel.py attention 20 512 2>&1 | tail -1
HB_ATTENTION_V1=1 python tools/run_kernel.py attention 20 512 2>&1 | tail -1
HB_LIB_PATH=hipt_abmil_atec23_b200/lib/exp_at2trace.so python tools/exp_at2_trace.py 512 > gpurun_out/at2_trace3.txt 2>&1; tail -45 gpurun_out/at2_trace3.txt
