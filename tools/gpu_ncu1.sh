#!/bin/bash
# ncu --set full captures: CLAM tensor-core score kernel (1 fold, 5 folds), attention kernel
set -u
mkdir -p gpurun_out
python tools/bench_clam.py --size hipt_smaller --folds 1 > gpurun_out/ncu1_clam1_plain.log 2>&1 && \
timeout 600 ncu --set full --clock-control none --import-source on -k regex:clam_scores_tc -s 3 -c 1 -f -o gpurun_out/r02_clam_tc_1fold python tools/bench_clam.py --size hipt_smaller --folds 1 > gpurun_out/ncu1_clam1.log 2>&1; echo "clam1 rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:clam_scores_tc -s 3 -c 1 -f -o gpurun_out/r02_clam_tc_5fold python tools/bench_clam.py --size hipt_smaller --folds 5 > gpurun_out/ncu1_clam5.log 2>&1; echo "clam5 rc=$?"
python tools/run_kernel.py attention 5 512 > gpurun_out/ncu1_att_plain.log 2>&1 && \
timeout 600 ncu --set full --clock-control none --import-source on -k regex:attention_tc2 -s 3 -c 1 -f -o gpurun_out/r02_attention_tc2 python tools/run_kernel.py attention 5 512 > gpurun_out/ncu1_att.log 2>&1; echo "att rc=$?"
ls -la gpurun_out/*.ncu-rep
