#!/bin/bash
set -u
mkdir -p gpurun_out
HB_LIB_PATH=hipt_abmil_atec23_b200/lib/exp_at2dbg.so python tools/diag_attention_dbg.py 3000 > gpurun_out/r2h_dbg.txt 2>&1; tail -30 gpurun_out/r2h_dbg.txt | cut -c1-600
