#!/bin/bash
# Experimental build of the library with extra -D flags:  tools/build_exp.sh NAME -DHB_EXP_TRACE ...
# -> hipt_abmil_atec23_b200/lib/exp_NAME.so (select it with HB_LIB_PATH).  Not part of the product build.
set -e
NAME=$1; shift
cd "$(dirname "$0")/../hipt_abmil_atec23_b200/csrc"
mkdir -p ../lib
for f in hb_api hb_gemm hb_mlp hb_rowops hb_attention hb_attention_tc hb_clam hb_embed hb_ingest; do
  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC "$@" -diag-suppress 128 -c $f.cu -o /tmp/exp_${NAME}_$f.o &
done
wait
nvcc -shared -o ../lib/exp_${NAME}.so /tmp/exp_${NAME}_hb_*.o -lcuda -ldl
echo built ../lib/exp_${NAME}.so
