#!/bin/bash
# attention A/B: committed kernel (exp_at2old.so) vs working tree, then the in-kernel trace of the working tree
set -u
mkdir -p gpurun_out
python -m pytest tests/test_gpu_kernels.py -m gpu -x -q -k attention 2>&1 | tail -2
python tools/diag_attention_race.py 6000 256 > gpurun_out/r2g_race.txt 2>&1; tail -2 gpurun_out/r2g_race.txt
python tools/diag_attention_race.py 1500 512 2>&1 | tail -1
for i in 1 2; do
HB_LIB_PATH=hipt_abmil_atec23_b200/lib/exp_at2old.so python tools/run_kernel.py attention 50 512 2>&1 | tail -1
python tools/run_kernel.py attention 50 512 2>&1 | tail -1
done
HB_LIB_PATH=hipt_abmil_atec23_b200/lib/exp_at2trace.so python tools/exp_at2_trace.py 512 > gpurun_out/at2_trace_new.txt 2>&1; tail -14 gpurun_out/at2_trace_new.txt
